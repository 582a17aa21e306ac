// loss.cu — stage 2 of MultiBoxLoss (multibox_loss.py:80-120) with its backward fused in.
//
// One thread-block cluster per image, contiguous prior slices per CTA.
//   sweep 1: tag (2 B) + conf row per prior; positives also read loc + prior, encode their target on
//            the fly (box_utils.py:114-135; never materialised), add smooth-L1 and emit d(loss_l)/d(loc);
//            every prior gets its mining key log(sum exp(x - x_max)) + x_max - x[0] with the
//            BATCH-GLOBAL x_max (box_utils.py:160-168, multibox_loss.py:91-99) as an ordered uint32 in
//            shared memory;
//   select : radix select of the min(ratio*num_pos, P-1) largest keys (select.cuh) — replaces the two
//            full sorts of multibox_loss.py:102-103;
//   sweep 2: cross-entropy (torch log_softmax, row max) over pos|neg and d(loss_c)/d(conf); every other
//            prior gets its zero gradients here (stores only); conf is re-read from L2;
//   finish : per-CTA partial sums in double, the last CTA reduces them in a fixed order and divides
//            by N (multibox_loss.py:117-119).
#include "select.cuh"

GSSD_PHASE_DECL(loss)

namespace gssd {

constexpr int LOSS_NT = 256;

struct LossArgs {
    const float4 *loc; const float *conf; const float4 *priors;
    int B, P, C;
    const float *gt; const int32_t *gt_off;
    const uint16_t *tags;
    uint32_t *stats;                        // local stats_buf (header + num_pos[B])
    const uint32_t *gstats; int n_gstats;   // headers of all ranks (or null)
    int ratio; float var0, var1;
    float *losses; float4 *grad_loc; float *grad_conf;
    uint8_t *pos_mask, *neg_mask;
    double *partials;                       // [2 * n_ctas]
    int slice;
    XDev x;                                 // peer exchange of the statistics (world == 0: off)
};

__device__ __forceinline__ float smooth_l1(float d, float &grad) {
    float ad = fabsf(d);
    if (ad < 1.f) { grad = d; return __fmul_rn(__fmul_rn(0.5f, d), d); }
    grad = d > 0.f ? 1.f : -1.f;
    return __fsub_rn(ad, 0.5f);
}

// C2: num_classes == 2 fast path (float2 rows)
template <bool C2, bool CLUSTER, bool GRADS>
__global__ void __launch_bounds__(LOSS_NT, 4) loss_kernel(LossArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SelectShared sel_s;
    __shared__ double s_red[2][LOSS_NT / 32];
    __shared__ bool s_last;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = CLUSTER ? cluster.block_rank() : 0;
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g0 = a.gt_off[b];
    const int G = a.gt_off[b + 1] - g0;
    const int C = C2 ? 2 : a.C;

    if (CLUSTER) {                       // the select's receiving buffers are cleared long before any CTA of the image adds into them
        radix_select_prepare<LOSS_NT>(&sel_s);
        cluster_arrive();
    }
    float *sgt = reinterpret_cast<float *>(smem_raw);                    // [G][5]
    uint32_t *keys = reinterpret_cast<uint32_t *>(sgt + 5 * ((G + 3) & ~3));
    for (int i = tid; i < 5 * G; i += LOSS_NT) sgt[i] = a.gt[5 * (size_t)g0 + i];

    // batch-global scalars: max over ranks of x_max, sum over ranks of N
    float x_max; int n_total;
    uint32_t x_epoch = 0;
    if (a.x.world > 0) {
        // peer exchange: wait until every rank's stage 1 has stored its statistics of THIS step into our buffer
        __shared__ uint32_t s_mo; __shared__ int s_nt;
        const XBuf *xl = a.x.peers[a.x.rank];
        x_epoch = *reinterpret_cast<const volatile uint32_t *>(&xl->epoch) + 1;
        if (warp == 0) {
            uint32_t mo = 0; int nt = 0;
            if (lane < a.x.world) {
                const unsigned long long *slots[2] = {&xl->xmax[x_epoch & 1][lane], &xl->npos[x_epoch & 1][lane]};
                unsigned long long got[2];
                for (int q = 0; q < 2; ++q) {
                    unsigned long long t0 = 0;
                    unsigned spins = 0;
                    while (true) {
                        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got[q]) : "l"(slots[q]) : "memory");
                        if ((uint32_t)(got[q] >> 32) == x_epoch) break;
                        if ((++spins & 0xff) == 0 && a.x.timeout_ns) {     // a rank that never arrives must not wedge the GPU
                            unsigned long long now;
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                            if (t0 == 0) t0 = now;
                            else if (now - t0 > a.x.timeout_ns) __trap();
                        }
                    }
                }
                mo = (uint32_t)got[0]; nt = (int)(uint32_t)got[1];
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) { mo = max(mo, __shfl_xor_sync(FULL, mo, o)); nt += __shfl_xor_sync(FULL, nt, o); }
            if (lane == 0) { s_mo = mo; s_nt = nt; }
        }
        __syncthreads();
        x_max = ord2f(s_mo); n_total = s_nt;
    } else {
        uint32_t mo = a.stats[0]; int nt = (int)a.stats[1];
        if (a.gstats) {
            mo = 0; nt = 0;
            for (int r = 0; r < a.n_gstats; ++r) { mo = max(mo, a.gstats[4 * r]); nt += (int)a.gstats[4 * r + 1]; }
        }
        x_max = ord2f(mo); n_total = nt;
    }
    const float n_f = (float)n_total;
    const int num_pos = reinterpret_cast<const int *>(a.stats + 4)[b];
    __syncthreads();

    const int p0 = rank * a.slice;
    const int p1 = min(a.P, p0 + a.slice);
    const int n_local = max(p1 - p0, 0);
    double acc_l = 0.0, acc_c = 0.0;
    const bool dbg = blockIdx.x == 0 && blockIdx.y == 0;
    GSSD_PHASE(loss, 0, dbg);

    // ---- sweep 1 ---------------------------------------------------------------------------------
    // U priors per thread per trip, loads first: tag (2 B) + conf row.  No gradient is stored here except d(loss_l)/d(loc) of
    // the few positives: the cluster barriers of the select below start with a GPU-scope memory barrier (that is what
    // barrier.cluster.arrive.release compiles to), which would wait for every store still in flight.
    uint32_t *posbits = keys + ((a.slice + 3) & ~3);             // one bit per local prior
    for (int i = tid; i < (n_local + 31) / 32; i += LOSS_NT) posbits[i] = 0;
    __syncthreads();
    constexpr int U = 4;
    for (int base = p0; base < p1; base += LOSS_NT * U) {
        uint16_t tg[U];
        float2 x2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = base + u * LOSS_NT + tid;
            const size_t o = (size_t)b * a.P + min(p, p1 - 1);
            tg[u] = __ldcs(a.tags + o);
            if (C2) x2[u] = ldg_stream(reinterpret_cast<const float2 *>(a.conf + o * 2));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = base + u * LOSS_NT + tid;
            const bool ok = p < p1;
            const size_t o = (size_t)b * a.P + min(p, p1 - 1);
            const uint16_t tag = tg[u];
            const bool pos = ok && (tag & 0x8000);
            float key;
            if (C2) {
                const float2 x = x2[u];
                const float s = __fadd_rn(expf(__fsub_rn(x.x, x_max)), expf(__fsub_rn(x.y, x_max)));
                key = __fsub_rn(__fadd_rn(logf(s), x_max), x.x);
            } else {
                const float *row = a.conf + o * C;
                float s = 0.f;
                for (int c = 0; c < C; ++c) s = __fadd_rn(s, expf(__fsub_rn(row[c], x_max)));
                key = __fsub_rn(__fadd_rn(logf(s), x_max), row[0]);
            }
            key = pos ? 0.f : __fadd_rn(key, 0.f);               // loss_c[pos] = 0 ; -0 -> +0
            const unsigned pm = __ballot_sync(FULL, pos);
            if (ok) keys[p - p0] = f2ord(key);
            if (pm && lane == 0) posbits[(p - p0) >> 5] = pm;    // p - p0 is a multiple of 32 for lane 0
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pos) {
                const float *t = sgt + 5 * (tag & 0x7fff);
                const float4 lt = encode_box(make_float4(t[0], t[1], t[2], t[3]), a.priors[p], a.var0, a.var1);
                const float4 l = ldg_stream(a.loc + o);
                const float l0 = smooth_l1(__fsub_rn(l.x, lt.x), g4.x), l1 = smooth_l1(__fsub_rn(l.y, lt.y), g4.y);
                const float l2 = smooth_l1(__fsub_rn(l.z, lt.z), g4.z), l3 = smooth_l1(__fsub_rn(l.w, lt.w), g4.w);
                acc_l += (double)l0 + (double)l1 + (double)l2 + (double)l3;
                g4.x = __fdiv_rn(g4.x, n_f); g4.y = __fdiv_rn(g4.y, n_f); g4.z = __fdiv_rn(g4.z, n_f); g4.w = __fdiv_rn(g4.w, n_f);
            }
            if (GRADS && pos) __stcs(&a.grad_loc[o], g4);        // the zeros of everybody else are written in sweep 2
        }
    }
    __syncthreads();
    GSSD_PHASE(loss, 1, dbg);

    // ---- hard-negative selection (multibox_loss.py:102-106) -----------------------------------------
    long long k = (long long)a.ratio * num_pos;
    if (k > a.P - 1) k = a.P - 1;
    const bool have_sel = k > 0;
    unsigned long long cut = ~0ull;
    if (CLUSTER) cluster_wait();         // every CTA of the image is running and has cleared its buffers
#ifdef GSSD_PHASE_TIMING
    long long *sel_clk = dbg ? g_phase_clock_loss + 8 : nullptr;
#else
    long long *sel_clk = nullptr;
#endif
    if (have_sel) cut = radix_select<LOSS_NT, CLUSTER>(keys, n_local, (uint32_t)p0, (uint32_t)k, true, &sel_s, sel_clk);

    GSSD_PHASE(loss, 2, dbg);
    // ---- sweep 2: only pos | neg priors do any work -----------------------------------------------------
    for (int p = p0 + tid; p < p1; p += LOSS_NT) {
        const size_t o = (size_t)b * a.P + p;
        const int li = p - p0;
        const bool pos = (posbits[li >> 5] >> (li & 31)) & 1u;
        const bool neg = have_sel && sel_composite(keys[li], (uint32_t)p, true) >= cut;
        if (a.pos_mask) a.pos_mask[o] = pos;
        if (a.neg_mask) a.neg_mask[o] = neg;
        if (GRADS && !pos) __stcs(&a.grad_loc[o], make_float4(0.f, 0.f, 0.f, 0.f));
        if (!(pos || neg)) {
            if (GRADS) {
                if (C2) __stcs(reinterpret_cast<float2 *>(a.grad_conf + o * 2), make_float2(0.f, 0.f));
                else for (int c = 0; c < C; ++c) a.grad_conf[o * C + c] = 0.f;
            }
            continue;
        }
        const float *row = a.conf + o * C;
        const int t = pos ? (int)__fadd_rn(sgt[5 * (a.tags[o] & 0x7fff) + 4], 1.f) : 0;
        if (C2) {
            const float2 x = *reinterpret_cast<const float2 *>(row);
            const float m = fmaxf(x.x, x.y);
            const float e0 = expf(__fsub_rn(x.x, m)), e1 = expf(__fsub_rn(x.y, m));
            const float ls = logf(__fadd_rn(e0, e1));
            const float xt = t == 0 ? x.x : x.y;
            acc_c += (double)(-(__fsub_rn(__fsub_rn(xt, m), ls)));
            if (GRADS) {
                float2 gz;
                gz.x = __fdiv_rn(expf(__fsub_rn(__fsub_rn(x.x, m), ls)) - (t == 0 ? 1.f : 0.f), n_f);
                gz.y = __fdiv_rn(expf(__fsub_rn(__fsub_rn(x.y, m), ls)) - (t == 1 ? 1.f : 0.f), n_f);
                __stcs(reinterpret_cast<float2 *>(a.grad_conf + o * 2), gz);
            }
        } else {
            float m = row[0];
            for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
            float s = 0.f;
            for (int c = 0; c < C; ++c) s = __fadd_rn(s, expf(__fsub_rn(row[c], m)));
            const float ls = logf(s);
            const float xt = (t >= 0 && t < C) ? row[t] : row[0];
            acc_c += (double)(-(__fsub_rn(__fsub_rn(xt, m), ls)));
            if (GRADS)
                for (int c = 0; c < C; ++c)
                    a.grad_conf[o * C + c] = __fdiv_rn(expf(__fsub_rn(__fsub_rn(row[c], m), ls)) - (c == t ? 1.f : 0.f), n_f);
        }
    }

    GSSD_PHASE(loss, 3, dbg);
    // ---- finish ----------------------------------------------------------------------------------
    acc_l = warp_sum(acc_l); acc_c = warp_sum(acc_c);
    if (lane == 0) { s_red[0][warp] = acc_l; s_red[1][warp] = acc_c; }
    __syncthreads();
    const unsigned n_ctas = gridDim.x * gridDim.y;
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double l = 0.0, c = 0.0;
        for (int w = 0; w < LOSS_NT / 32; ++w) { l += s_red[0][w]; c += s_red[1][w]; }
        a.partials[2 * cta] = l; a.partials[2 * cta + 1] = c;
        __threadfence();
        unsigned done = atomicAdd(&a.stats[2], 1u);
        s_last = done == n_ctas - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double l = 0.0, c = 0.0;
        for (unsigned i = tid; i < n_ctas; i += LOSS_NT) {           // fixed order -> deterministic
            l += __ldcg(&a.partials[2 * i]); c += __ldcg(&a.partials[2 * i + 1]);
        }
        l = warp_sum(l); c = warp_sum(c);
        if (lane == 0) { s_red[0][warp] = l; s_red[1][warp] = c; }
        __syncthreads();
        if (tid == 0) {
            l = 0.0; c = 0.0;
            for (int w = 0; w < LOSS_NT / 32; ++w) { l += s_red[0][w]; c += s_red[1][w]; }
            a.losses[0] = __fdiv_rn((float)l, (float)n_total);        // multibox_loss.py:117-119
            a.losses[1] = __fdiv_rn((float)c, (float)n_total);
            a.stats[2] = 0;
            if (a.x.world > 0) *reinterpret_cast<volatile uint32_t *>(&a.x.peers[a.x.rank]->epoch) = x_epoch;   // the step is complete
        }
    }
    GSSD_PHASE(loss, 4, dbg);
}

static size_t loss_smem_bytes(int g_max, int slice) {
    return (size_t)((g_max + 3) & ~3) * 5 * 4 + (size_t)((slice + 3) & ~3) * 4 + (size_t)((slice + 31) / 32) * 4 + 32;
}

template <bool C2, bool GR>
static int launch_loss(LossArgs &a, int g_max, cudaStream_t stream) {
    // the cluster-capable instantiation also runs as a 1-CTA "cluster"; pick the size from its occupancy
    auto kern_cl = loss_kernel<C2, true, GR>;
    GSSD_RETURN_IF_CUDA(allow_max_smem(reinterpret_cast<const void *>(kern_cl)));
    const int S = pick_cluster_size(GSSD_KERNEL_LOSS, a.B, a.P);
    a.slice = ceil_div(a.P, S);
    auto kern = S > 1 ? kern_cl : loss_kernel<C2, false, GR>;
    size_t smem = loss_smem_bytes(g_max, a.slice);
    GSSD_RETURN_IF_CUDA(allow_max_smem(reinterpret_cast<const void *>(kern)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(S, a.B, 1);
    cfg.blockDim = dim3(LOSS_NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GSSD_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

// ---- backward helper: in-place scaling by the upstream gradients, free when they are 1 ------------
__global__ void __launch_bounds__(256) scale_grads_kernel(float4 *gl, size_t n4_loc, float4 *gc, size_t n4_conf,
                                                          float *gc_tail, int n_tail,
                                                          const float *g_loc, const float *g_conf) {
    const float sl = g_loc ? *g_loc : 1.f, sc = g_conf ? *g_conf : 1.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sl != 1.f)
        for (size_t i = t0; i < n4_loc; i += stride) { float4 v = gl[i]; v.x *= sl; v.y *= sl; v.z *= sl; v.w *= sl; gl[i] = v; }
    if (sc != 1.f) {
        for (size_t i = t0; i < n4_conf; i += stride) { float4 v = gc[i]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; gc[i] = v; }
        if (t0 < (size_t)n_tail) gc_tail[t0] *= sc;
    }
}

}  // namespace gssd

using namespace gssd;

extern "C" size_t gssd_stats_bytes(int B) { return sizeof(gssd_loss_stats) + sizeof(int32_t) * (size_t)(B > 0 ? B : 0); }

static int mbox_loss_impl(const float *loc, const float *conf, const float *priors, int B, int P, int C,
                          const float *gt, const int32_t *gt_off, int sum_G, int g_max,
                          const uint16_t *tags, void *stats_buf,
                          const gssd_loss_stats *global_stats, int n_global_stats, const gssd_xchg *x,
                          int negpos_ratio, float var0, float var1,
                          float *losses, float *grad_loc, float *grad_conf,
                          uint8_t *pos_mask, uint8_t *neg_mask,
                          void *ws, size_t ws_bytes, void *stream) {
    if (x && (x->world < 1 || x->world > GSSD_XCHG_MAX_RANKS || x->world > 32 || x->rank < 0 || x->rank >= x->world)) return GSSD_ERR_ARG;
    if (!loc || !conf || !priors || !gt || !gt_off || !tags || !stats_buf || !losses || !ws) return GSSD_ERR_ARG;
    if (B <= 0 || P <= 0 || sum_G <= 0 || g_max <= 0 || negpos_ratio < 0) return GSSD_ERR_ARG;
    if (C < 2 || C > GSSD_MAX_CLASSES) return GSSD_ERR_ARG;
    if ((grad_loc == nullptr) != (grad_conf == nullptr)) return GSSD_ERR_ARG;
    if (g_max > GSSD_MAX_GT_PER_IMAGE || P > GSSD_MAX_PRIORS) return GSSD_ERR_LIMIT;
    if (global_stats && n_global_stats <= 0) return GSSD_ERR_ARG;
    if (ws_bytes < gssd_workspace_bytes(GSSD_WS_LOSS, B, P, C, sum_G, 0)) return GSSD_ERR_WS;
    LossArgs a = {};
    a.loc = reinterpret_cast<const float4 *>(loc); a.conf = conf; a.priors = reinterpret_cast<const float4 *>(priors);
    a.B = B; a.P = P; a.C = C; a.gt = gt; a.gt_off = gt_off; a.tags = tags;
    a.stats = reinterpret_cast<uint32_t *>(stats_buf);
    a.gstats = reinterpret_cast<const uint32_t *>(global_stats); a.n_gstats = n_global_stats;
    a.ratio = negpos_ratio; a.var0 = var0; a.var1 = var1;
    a.losses = losses; a.grad_loc = reinterpret_cast<float4 *>(grad_loc); a.grad_conf = grad_conf;
    a.pos_mask = pos_mask; a.neg_mask = neg_mask;
    a.partials = reinterpret_cast<double *>(ws);
    a.x = xdev_from(x);
    cudaStream_t st = (cudaStream_t)stream;
    const bool gr = grad_loc != nullptr;
    const bool c2 = C == 2;
    if (c2) return gr ? launch_loss<true, true>(a, g_max, st) : launch_loss<true, false>(a, g_max, st);
    return gr ? launch_loss<false, true>(a, g_max, st) : launch_loss<false, false>(a, g_max, st);
}

extern "C" int gssd_mbox_loss(const float *loc, const float *conf, const float *priors, int B, int P, int C,
                              const float *gt, const int32_t *gt_off, int sum_G, int g_max,
                              const uint16_t *tags, void *stats_buf,
                              const gssd_loss_stats *global_stats, int n_global_stats,
                              int negpos_ratio, float var0, float var1,
                              float *losses, float *grad_loc, float *grad_conf,
                              uint8_t *pos_mask, uint8_t *neg_mask,
                              void *ws, size_t ws_bytes, void *stream) {
    return mbox_loss_impl(loc, conf, priors, B, P, C, gt, gt_off, sum_G, g_max, tags, stats_buf, global_stats, n_global_stats, nullptr,
                          negpos_ratio, var0, var1, losses, grad_loc, grad_conf, pos_mask, neg_mask, ws, ws_bytes, stream);
}

extern "C" int gssd_mbox_loss_x(const float *loc, const float *conf, const float *priors, int B, int P, int C,
                                const float *gt, const int32_t *gt_off, int sum_G, int g_max,
                                const uint16_t *tags, void *stats_buf, const gssd_xchg *x,
                                int negpos_ratio, float var0, float var1,
                                float *losses, float *grad_loc, float *grad_conf,
                                uint8_t *pos_mask, uint8_t *neg_mask,
                                void *ws, size_t ws_bytes, void *stream) {
    if (!x) return GSSD_ERR_ARG;
    return mbox_loss_impl(loc, conf, priors, B, P, C, gt, gt_off, sum_G, g_max, tags, stats_buf, nullptr, 0, x,
                          negpos_ratio, var0, var1, losses, grad_loc, grad_conf, pos_mask, neg_mask, ws, ws_bytes, stream);
}

extern "C" int gssd_mbox_scale_grads(float *grad_loc, size_t n_loc, float *grad_conf, size_t n_conf,
                                     const float *g_loc, const float *g_conf, void *stream) {
    if (!grad_loc || !grad_conf) return GSSD_ERR_ARG;
    if (n_loc % 4) return GSSD_ERR_ARG;
    size_t n4c = n_conf / 4; int tail = (int)(n_conf % 4);
    size_t work = (n_loc / 4 > n4c ? n_loc / 4 : n4c);
    // one CTA per SM: in the usual case (both upstream gradients are 1: `(loss_l + loss_c).backward()`) every CTA only reads the
    // two scalars and leaves, and the fewer there are to launch and retire the sooner that is over
    int blocks = (int)((work + 255) / 256);
    if (blocks > 148) blocks = 148;
    if (blocks < 1) blocks = 1;
    scale_grads_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4 *>(grad_loc), n_loc / 4, reinterpret_cast<float4 *>(grad_conf), n4c,
        grad_conf + n4c * 4, tail, g_loc, g_conf);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}
