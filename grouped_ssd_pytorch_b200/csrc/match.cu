// match.cu — prior <-> ground-truth matching (box_utils.py:70-111, batched over the loop at
// multibox_loss.py:67-72).
//
// One thread-block cluster per image (1..8 CTAs, contiguous prior slices).  GT boxes are staged in
// shared memory; each thread sweeps its priors (one coalesced float4 each), keeps the best GT per
// prior in registers and feeds a per-GT running best prior through a warp redux + ballot into a
// shared-memory atomicMax on a packed (IoU bits, ~prior index) key, so that ties resolve to the lowest
// prior index exactly like torch.max.  The per-GT winners are combined over distributed shared
// memory, then the sequential "force match" (box_utils.py:101-105, last GT wins) is applied by one
// thread, and the result is emitted either as 2-byte tags for the fused loss or as the reference's
// materialised loc_t / conf_t.
#include <string.h>

#include "common.cuh"

GSSD_PHASE_DECL(match)

namespace gssd {

constexpr int MATCH_NT = 256;
#ifndef GSSD_MATCH_MIN_CTAS
#define GSSD_MATCH_MIN_CTAS 5                  // resident CTAs per SM the register allocation is held to
#endif

struct MatchArgs {
    const float4 *priors; int P;
    const float *conf; int C;                 // optional: batch max of conf (stage 1 of the loss)
    const float *gt; const int32_t *gt_off;
    float threshold, var0, var1;
    uint16_t *tags;                           // [B,P] or null
    uint32_t *stats;                          // stats header (conf_max_ord, num_pos_total, ...) or null
    int32_t *num_pos;                         // [B] or null
    float4 *loc_t; int64_t *conf_t; int32_t *bti_out;   // materialised outputs or null
    int slice;                                // priors per CTA
    int S;                                    // CTAs per image (cluster size)
    XDev x;                                   // peer exchange of the statistics (world == 0: off)
};

// dynamic shared memory layout
//   u64    sbest[G]   packed (IoU bits, ~prior) best prior per GT, this CTA's slice
//   u64    sin[S][G]  the same from every CTA of the image, stored here by its owner (distributed shared memory)
//   float4 sgt4[G]    GT boxes
//   float  sarea[G], slabel[G]
//   int    sbp[G]     best prior per GT over the whole image
//   int    glist[G]   GT rows that can overlap this CTA's prior slice (ascending)
//   u16    stag[slice]
static size_t match_smem_bytes(int g_max, int slice, int S) {
    return (size_t)g_max * (8 + 16 + 4 + 4 + 4 + 4) + (size_t)(g_max + 1) * 8 * (S > 1 ? S : 0) + (size_t)slice * 2 + 32;
}

constexpr int MATCH_CULL_MIN_G = 8;       // below this the bounding-box pre-pass costs more than it saves

template <bool MATERIALISE, bool CONF_MAX>
__global__ void __launch_bounds__(MATCH_NT, GSSD_MATCH_MIN_CTAS) match_kernel(MatchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned nranks = cluster.num_blocks();
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g0 = a.gt_off[b];
    const int G = a.gt_off[b + 1] - g0;

    unsigned long long *sbest = reinterpret_cast<unsigned long long *>(smem_raw);     // 16-byte aligned first
    const int Gp = (G + 1) & ~1;
    unsigned long long *sin = sbest + Gp;                                             // [nranks][Gp], clusters only
    float4 *sgt4 = reinterpret_cast<float4 *>(sin + (nranks > 1 ? nranks * Gp : 0));
    float *sarea = reinterpret_cast<float *>(sgt4 + G);
    float *slabel = sarea + G;
    int *sbp = reinterpret_cast<int *>(slabel + G);
    int *glist = sbp + G;
    uint16_t *stag = reinterpret_cast<uint16_t *>(glist + G);
    __shared__ int s_warp_cnt[MATCH_NT / 32];
    __shared__ float s_warp_max[MATCH_NT / 32];
    __shared__ float s_bbox[4][MATCH_NT / 32];
    __shared__ int s_nlist;
    __shared__ bool s_is_last;

    if (nranks > 1) cluster_arrive();    // waited for right before the first store into another CTA's shared memory
    // conf-max pass: a contiguous slice per CTA (uniform, streaming work)
    const int p0 = rank * a.slice;
    const int p1 = min(a.P, p0 + a.slice);
    // IoU sweep: the CTAs of the image own the priors in interleaved chunks of MATCH_NT (CTA r: chunks r, r + S, ...).  The
    // work per prior depends on how many GT boxes it overlaps — the large priors at the end of the list overlap most of
    // them — so contiguous slices leave the first CTAs waiting for the last ones at the cluster barrier.
    const int n_chunks = (a.P + MATCH_NT - 1) / MATCH_NT;
    const int my_chunks = (int)rank < n_chunks ? (n_chunks - (int)rank + (int)nranks - 1) / (int)nranks : 0;

    for (int g = tid; g < G; g += MATCH_NT) {
        const float *row = a.gt + 5 * (size_t)(g0 + g);
        const float4 t = make_float4(row[0], row[1], row[2], row[3]);
        sgt4[g] = t;
        sarea[g] = box_area(t);
        slabel[g] = row[4];
        // an all-zero IoU row resolves to prior 0 (torch.max returns the first maximum)
        sbest[g] = 0x00000000ffffffffull;
        glist[g] = g;
    }
    if (tid == 0) s_nlist = G;
    __syncthreads();

    // ---- optional: drop the GT boxes that cannot touch this CTA's slice of priors --------------------------
    if (G >= MATCH_CULL_MIN_G) {
        float bx1 = INFINITY, by1 = INFINITY, bx2 = -INFINITY, by2 = -INFINITY;
        for (int j = 0; j < my_chunks; ++j) {
            const int p = ((int)rank + j * (int)nranks) * MATCH_NT + tid;
            if (p >= a.P) continue;
            const float4 pb = point_form(a.priors[p]);
            bx1 = fminf(bx1, pb.x); by1 = fminf(by1, pb.y); bx2 = fmaxf(bx2, pb.z); by2 = fmaxf(by2, pb.w);
        }
        bx1 = warp_min(bx1); by1 = warp_min(by1); bx2 = warp_max(bx2); by2 = warp_max(by2);
        if (lane == 0) { s_bbox[0][warp] = bx1; s_bbox[1][warp] = by1; s_bbox[2][warp] = bx2; s_bbox[3][warp] = by2; }
        __syncthreads();
        if (warp == 0) {
            for (int w = 0; w < MATCH_NT / 32; ++w) {
                bx1 = fminf(bx1, s_bbox[0][w]); by1 = fminf(by1, s_bbox[1][w]);
                bx2 = fmaxf(bx2, s_bbox[2][w]); by2 = fmaxf(by2, s_bbox[3][w]);
            }
            int n = 0;                                           // ordered compaction, 32 GT rows per step
            for (int gb = 0; gb < G; gb += 32) {
                const int g = gb + lane;
                bool hit = false;
                if (g < G) {
                    const float4 t = sgt4[g];
                    hit = fminf(t.z, bx2) > fmaxf(t.x, bx1) && fminf(t.w, by2) > fmaxf(t.y, by1);
                }
                const unsigned m = __ballot_sync(FULL, hit);
                if (hit) glist[n + __popc(m & ((1u << lane) - 1))] = g;
                n += __popc(m);
            }
            if (lane == 0) s_nlist = n;
        }
        __syncthreads();
    }
    const int n_list = s_nlist;

    float cmax = -INFINITY;
    const bool dbg = blockIdx.x == 0 && blockIdx.y == 0;
    GSSD_PHASE(match, 0, dbg);

    // ---- batch max of conf, part 1: ask L2 for this CTA's rows now; they are reduced after the IoU sweep, which needs no
    // HBM traffic of its own (the priors stay in L2), so the DRAM latency / transfer hides behind it ----------------
    if (CONF_MAX && p1 > p0) {
        const char *c0 = reinterpret_cast<const char *>(a.conf + ((size_t)b * a.P + p0) * a.C);
        const size_t nbytes = (size_t)(p1 - p0) * a.C * sizeof(float);
        for (size_t o = (size_t)tid * 128; o < nbytes; o += (size_t)MATCH_NT * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(c0 + o));
    }

    // ---- IoU sweep, one warp per 32 consecutive priors ------------------------------------------------------
    // A disjoint (GT, prior) pair costs two min/max/sub per axis and one test.  With many GT boxes the 32
    // priors of a warp (neighbouring cells / anchors) first test the GT list against their common bounding
    // box, one GT per lane, and only the GT rows that hit it are swept.
    const bool warp_cull = n_list >= MATCH_CULL_MIN_G;
    const int chunk_stride = a.S * MATCH_NT;                     // a.S == nranks, as a kernel constant
    // a lane past the end of the list repeats the LAST prior: it produces that prior's own (IoU, index) keys a second time,
    // which the running maxima ignore, and its tag lands in the padding of the chunk
    const int p_last = a.P - 1;
    int p = (int)rank * MATCH_NT + tid;
    uint16_t *stag_w = stag + tid;
    float4 nxt = a.priors[min(p, p_last)];
    for (int j = my_chunks; j > 0; --j, p += chunk_stride, stag_w += MATCH_NT) {
        const int pc = min(p, p_last);
        const float4 pb = point_form(nxt);
        nxt = a.priors[min(p + chunk_stride, p_last)];           // prefetch (the last round re-reads a cached row)
        const float area_b = box_area(pb);
        float best = 0.f;                                        // IoU >= 0: row 0 wins an all-zero column
        int bidx = 0;
        auto sweep_one = [&](int g) {
            const float4 t = sgt4[g];
            const float iw = __fsub_rn(fminf(t.z, pb.z), fmaxf(t.x, pb.x));
            const float ih = __fsub_rn(fminf(t.w, pb.w), fmaxf(t.y, pb.y));
            if (iw > 0.f && ih > 0.f) {                          // the boxes overlap
                const float inter = __fmul_rn(iw, ih);
                const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sarea[g], area_b), inter));
                if (iou > best) { best = iou; bidx = g; }        // first max over GT (torch.max dim 0)
                // best prior of this GT: max of (IoU bits, ~prior); IoU >= +0 so the uint order of the bits is the float
                // order, ~prior breaks ties downwards.  Only the lanes that beat the running maximum (few, after the first
                // rows) go on to the warp reduction and the shared-memory atomic.
                const unsigned bits = __float_as_uint(iou);
                const unsigned long long key = ((unsigned long long)bits << 32) | (0xffffffffu - (unsigned)pc);
                if (key > sbest[g]) {
                    const unsigned act = __activemask();
                    const unsigned m = __reduce_max_sync(act, bits);
                    const unsigned who = __ballot_sync(act, bits == m);
                    if (lane == __ffs(who) - 1) atomicMax(&sbest[g], key);
                }
            }
        };
        if (G < MATCH_CULL_MIN_G) {                              // no culling at all: glist is the identity
#pragma unroll 1
            for (int g = 0; g < G; ++g) sweep_one(g);
        } else if (!warp_cull) {
            for (int q = 0; q < n_list; ++q) sweep_one(glist[q]);
        } else {
            float bx1 = pb.x, by1 = pb.y, bx2 = pb.z, by2 = pb.w;
            bx1 = warp_min(bx1); by1 = warp_min(by1); bx2 = warp_max(bx2); by2 = warp_max(by2);
            for (int gb = 0; gb < n_list; gb += 32) {
                const int q = gb + lane;
                bool hit = false;
                if (q < n_list) {
                    const float4 t = sgt4[glist[q]];
                    hit = fminf(t.z, bx2) > fmaxf(t.x, bx1) && fminf(t.w, by2) > fmaxf(t.y, by1);
                }
                unsigned m = __ballot_sync(FULL, hit);
                while (m) {                                      // ascending GT row: first-max order is kept
                    const int qq = gb + __ffs(m) - 1;
                    m &= m - 1;
                    sweep_one(glist[qq]);
                }
            }
        }
        *stag_w = (uint16_t)(bidx | (!(best < a.threshold) ? 0x8000 : 0));   // box_utils.py:108
    }
    // ---- batch max of conf, part 2: a streaming, vectorised pass over the rows prefetched above ---------
    if (CONF_MAX && p1 > p0) {
        const float *src = a.conf + ((size_t)b * a.P + p0) * a.C;
        const size_t n = (size_t)(p1 - p0) * a.C;
        size_t head = ((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15) >> 2;   // floats up to 16-byte alignment
        if (head > n) head = n;
        const size_t n4 = (n - head) >> 2;
        const float4 *src4 = reinterpret_cast<const float4 *>(src + head);
        for (size_t i = tid; i < head; i += MATCH_NT) cmax = fmaxf(cmax, src[i]);
        constexpr int U = 4;
        for (size_t base = 0; base < n4; base += MATCH_NT * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t i = base + (size_t)u * MATCH_NT + tid;
                v[u] = i < n4 ? ldg_stream(src4 + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) cmax = fmaxf(cmax, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
        }
        for (size_t i = head + 4 * n4 + tid; i < n; i += MATCH_NT) cmax = fmaxf(cmax, src[i]);
    }
    __syncthreads();
    GSSD_PHASE(match, 1, dbg);

    // ---- best prior per GT over the whole image, then the sequential force match -------------------
    if (nranks > 1) {
        // every CTA stores its per-slice maxima into its own row of every other CTA's `sin`: plain remote stores that need no
        // answer, one barrier, and nothing is read remotely afterwards — no CTA has to outlive another
        cluster_wait();                                              // all CTAs of the image are running
        for (unsigned i = tid; i < (unsigned)G * nranks; i += MATCH_NT) {
            const unsigned g = i / nranks, r = i - g * nranks;
            if (r != rank) cluster.map_shared_rank(sin, r)[rank * Gp + g] = sbest[g];
        }
        cluster.sync();
        for (int g = tid; g < G; g += MATCH_NT) {
            unsigned long long m = sbest[g];
            for (unsigned r = 0; r < nranks; ++r)
                if (r != rank) m = max(m, sin[r * Gp + g]);
            sbest[g] = m;
        }
        __syncthreads();
    }
    for (int g = tid; g < G; g += MATCH_NT) sbp[g] = (int)(0xffffffffu - (unsigned)(sbest[g] & 0xffffffffu));
    __syncthreads();
    if (tid == 0) {
        for (int g = 0; g < G; ++g) {                            // box_utils.py:101-105, in GT order
            const int bp = sbp[g];
            const int ck = bp / MATCH_NT;
            if (ck % (int)nranks == (int)rank) stag[(ck / (int)nranks) * MATCH_NT + bp % MATCH_NT] = (uint16_t)(0x8000 | g);
        }
    }
    __syncthreads();

    GSSD_PHASE(match, 2, dbg);
    // ---- emit --------------------------------------------------------------------------------------
    int npos = 0;
    for (int j = 0; j < my_chunks; ++j) {
        const int p = (int)rank * MATCH_NT + j * chunk_stride + tid;
        if (p >= a.P) continue;
        const uint16_t tag = stag[j * MATCH_NT + tid];
        const bool pos = tag & 0x8000;
        const int g = tag & 0x7fff;
        npos += pos;
        const size_t o = (size_t)b * a.P + p;
        if (a.tags) a.tags[o] = tag;
        if (MATERIALISE) {
            a.loc_t[o] = encode_box(sgt4[g], a.priors[p], a.var0, a.var1);            // box_utils.py:109-110
            const float c = pos ? __fadd_rn(slabel[g], 1.f) : 0.f;                    // 107-108
            a.conf_t[o] = (int64_t)c;                                                 // 111
            if (a.bti_out) a.bti_out[o] = g;
        }
    }
    if (a.num_pos || CONF_MAX) {
        npos = warp_sum(npos);
        cmax = warp_max(cmax);
        if (lane == 0) { s_warp_cnt[warp] = npos; s_warp_max[warp] = cmax; }
        __syncthreads();
        if (tid == 0) {
            int tot = 0; float mx = -INFINITY;
            for (int w = 0; w < MATCH_NT / 32; ++w) { tot += s_warp_cnt[w]; mx = fmaxf(mx, s_warp_max[w]); }
            if (a.num_pos) atomicAdd(&a.num_pos[b], tot);
            if (a.stats) {
                atomicAdd(reinterpret_cast<int *>(&a.stats[1]), tot);
                if (CONF_MAX && p1 > p0) atomicMax(&a.stats[0], f2ord(mx));
            }
            if (a.x.world > 0) {
                // the LAST CTA of the grid publishes this rank's statistics into every peer's exchange buffer
                XBuf *xl = a.x.peers[a.x.rank];
                __threadfence();
                s_is_last = atomicAdd(&xl->match_done, 1u) == gridDim.x * gridDim.y - 1;
            }
        }
    }
    if (a.x.world > 0) {
        __syncthreads();
        if (s_is_last && warp == 0) {
            // one lane per peer: both statistics, each tagged with the step's epoch inside its own 64-bit word (common.cuh: XBuf)
            XBuf *xl = a.x.peers[a.x.rank];
            __threadfence();
            const uint32_t mo = atomicMax(&a.stats[0], 0u);
            const int np = atomicAdd(reinterpret_cast<int *>(&a.stats[1]), 0);
            const uint32_t e = *reinterpret_cast<volatile uint32_t *>(&xl->epoch) + 1;      // advanced by stage 2's last CTA
            if (lane < a.x.world) {
                XBuf *dst = a.x.peers[lane];
                *reinterpret_cast<volatile unsigned long long *>(&dst->xmax[e & 1][a.x.rank]) = ((unsigned long long)e << 32) | mo;
                *reinterpret_cast<volatile unsigned long long *>(&dst->npos[e & 1][a.x.rank]) = ((unsigned long long)e << 32) | (unsigned)np;
                __threadfence_system();
            }
            __syncwarp();
            if (lane == 0) xl->match_done = 0;
        }
    }
    GSSD_PHASE(match, 3, dbg);
    GSSD_PHASE(match, 4, dbg);
}

template <bool MAT, bool CMAX>
static int launch_match(const MatchArgs &a_in, int B, int g_max, cudaStream_t stream) {
    MatchArgs a = a_in;
    auto kern = match_kernel<MAT, CMAX>;
    GSSD_RETURN_IF_CUDA(allow_max_smem(reinterpret_cast<const void *>(kern)));
    const int S = pick_cluster_size(GSSD_KERNEL_MATCH, B, a.P);
    a.slice = ceil_div(a.P, S);
    a.S = S;
    size_t smem = match_smem_bytes(g_max, ceil_div(ceil_div(a.P, MATCH_NT), S) * MATCH_NT, S);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(S, B, 1);
    cfg.blockDim = dim3(MATCH_NT, 1, 1);
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

static int check_match_sizes(int P, int B, int sum_G, int g_max) {
    if (P <= 0 || B <= 0 || sum_G <= 0 || g_max <= 0) return sum_G <= 0 && B > 0 ? GSSD_ERR_EMPTY : GSSD_ERR_ARG;
    if (g_max > GSSD_MAX_GT_PER_IMAGE || P > GSSD_MAX_PRIORS) return GSSD_ERR_LIMIT;
    return GSSD_OK;
}

}  // namespace gssd

using namespace gssd;

extern "C" int gssd_match(const float *priors, int P, const float *gt, const int32_t *gt_off, int B,
                          int sum_G, int g_max, float threshold, float var0, float var1,
                          float *loc_t, int64_t *conf_t, int32_t *best_truth_idx,
                          void *ws, size_t ws_bytes, void *stream) {
    (void)ws; (void)ws_bytes;
    if (!priors || !gt || !gt_off || !loc_t || !conf_t) return GSSD_ERR_ARG;
    int rc = check_match_sizes(P, B, sum_G, g_max);
    if (rc) return rc;
    MatchArgs a = {};
    a.priors = reinterpret_cast<const float4 *>(priors); a.P = P;
    a.gt = gt; a.gt_off = gt_off; a.threshold = threshold; a.var0 = var0; a.var1 = var1;
    a.loc_t = reinterpret_cast<float4 *>(loc_t); a.conf_t = conf_t; a.bti_out = best_truth_idx;
    return launch_match<true, false>(a, B, g_max, (cudaStream_t)stream);
}

static int mbox_match_impl(const float *priors, int P, const float *conf, int C,
                           const float *gt, const int32_t *gt_off, int B, int sum_G, int g_max,
                           float threshold, uint16_t *tags, void *stats_buf, const gssd_xchg *x, void *stream) {
    if (!priors || !gt || !gt_off || !tags || !stats_buf) return GSSD_ERR_ARG;
    if (x && (x->world < 1 || x->world > GSSD_XCHG_MAX_RANKS || x->rank < 0 || x->rank >= x->world)) return GSSD_ERR_ARG;
    int rc = check_match_sizes(P, B, sum_G, g_max);
    if (rc) return rc;
    if (conf && (C < 2 || C > GSSD_MAX_CLASSES)) return GSSD_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(stats_buf, 0, sizeof(gssd_loss_stats) + sizeof(int32_t) * (size_t)B, st));
    MatchArgs a = {};
    a.priors = reinterpret_cast<const float4 *>(priors); a.P = P;
    a.conf = conf; a.C = C;
    a.gt = gt; a.gt_off = gt_off; a.threshold = threshold;
    a.tags = tags;
    a.stats = reinterpret_cast<uint32_t *>(stats_buf);
    a.num_pos = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(stats_buf) + sizeof(gssd_loss_stats));
    a.x = xdev_from(x);
    return conf ? launch_match<false, true>(a, B, g_max, st) : launch_match<false, false>(a, B, g_max, st);
}

extern "C" int gssd_mbox_match(const float *priors, int P, const float *conf, int C,
                               const float *gt, const int32_t *gt_off, int B, int sum_G, int g_max,
                               float threshold, uint16_t *tags, void *stats_buf, void *stream) {
    return mbox_match_impl(priors, P, conf, C, gt, gt_off, B, sum_G, g_max, threshold, tags, stats_buf, nullptr, stream);
}

extern "C" int gssd_mbox_match_x(const float *priors, int P, const float *conf, int C,
                                 const float *gt, const int32_t *gt_off, int B, int sum_G, int g_max,
                                 float threshold, uint16_t *tags, void *stats_buf, const gssd_xchg *x, void *stream) {
    if (!x) return GSSD_ERR_ARG;
    return mbox_match_impl(priors, P, conf, C, gt, gt_off, B, sum_G, g_max, threshold, tags, stats_buf, x, stream);
}

// ---- exchange buffers ---------------------------------------------------------------------------------------
extern "C" int gssd_xchg_create(void **xbuf_out, void *ipc_handle_out) {
    if (!xbuf_out || !ipc_handle_out) return GSSD_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == GSSD_XCHG_HANDLE_BYTES, "handle size");
    void *ptr = nullptr;
    GSSD_RETURN_IF_CUDA(cudaMalloc(&ptr, sizeof(XBuf)));
    cudaError_t e = cudaMemset(ptr, 0, sizeof(XBuf));
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(ipc_handle_out), ptr);
    if (e != cudaSuccess) { cudaFree(ptr); return (int)e; }
    *xbuf_out = ptr;
    return GSSD_OK;
}
extern "C" int gssd_xchg_open(const void *ipc_handle, void **peer_ptr_out) {
    if (!ipc_handle || !peer_ptr_out) return GSSD_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    GSSD_RETURN_IF_CUDA(cudaIpcOpenMemHandle(peer_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return GSSD_OK;
}
extern "C" int gssd_xchg_close(void *peer_ptr) { return peer_ptr ? (int)cudaIpcCloseMemHandle(peer_ptr) : GSSD_ERR_ARG; }
extern "C" int gssd_xchg_destroy(void *xbuf) { return xbuf ? (int)cudaFree(xbuf) : GSSD_ERR_ARG; }
