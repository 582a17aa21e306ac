// fused.cu — MultiBoxLoss.forward and its backward as ONE launch (multibox_loss.py:46-120 with the matching loop of
// lines 67-72 -> box_utils.match, box_utils.py:70-111, inside).
//
// Why one launch: at the batch the reference trains with (32 images) the whole problem is a few megabytes; two launches
// (match.cu then loss.cu) spend more time starting, draining and handing 16 bytes of statistics through global memory than
// moving data.  The only batch-wide dependencies of the loss are two scalars — the max of conf (box_utils.py:167) and the
// number of positives N (multibox_loss.py:117) — so all CTAs of the batch are made co-resident (cooperative launch) and
// meet twice through global memory: every CTA owns one word per value, (epoch << 32) | value, which it writes with a plain
// store and everybody polls — no atomics, no fences, nothing to reset; everything else is per image and stays inside one
// thread-block cluster.
//
// Per image: a cluster of S CTAs of NT threads; the priors are dealt to the CTAs in interleaved chunks of 256 (CTA r owns
// chunks r, r+S, ...: the large priors at the end of the list overlap most GT boxes, contiguous slices would leave the first
// CTAs waiting).  A thread owns the same priors in every phase, and the CTA keeps their conf rows, tags and mining keys in
// shared memory: conf is read from HBM exactly once.
//
//   A  conf rows -> shared memory (cp.async), local max -> this CTA's word of rendezvous 1
//   B  IoU sweep (as match.cu), per-GT best prior combined over the cluster (1 exchange), sequential force match
//   C  positives counted -> this CTA's word of rendezvous 2
//   D  wait for rendezvous 1 (long complete): mining keys with the batch-global max, first radix digit (11 bits) counted
//      on the fly
//   E  select of the min(ratio*num_pos, P-1) largest (key, lower index first) composites: the non-empty bins of every CTA
//      are pushed into every CTA's totals (1 exchange); the members of the bin that holds the cut (<= 256, typically
//      ~100) are pushed to every CTA (1 exchange) and ranked there by counting; more passes only when more than 256
//      composites share 11 / 22 / 32 ... leading bits (heavily tied keys) — the composite carries the prior index below the
//      key, so "more than 256 EQUAL keys" is just two more passes, with no special case
//   F  wait for rendezvous 2: smooth-L1 + CE over pos | neg, all gradients written (zeros for everybody else)
//   G  per-CTA partial sums in double; the last CTA to finish reduces them in a fixed order, divides by N, and advances the
//      epoch for the next launch
//
// Data-parallel jobs (gssd_xchg): every CTA stores its two words into the exchange buffer of EVERY rank over NVLink (8 bytes
// per peer and value) and the readers poll world x CTAs words of their LOCAL buffer: no aggregation step, no collective.  The
// max of conf is published microseconds after the kernel starts and is needed only after the IoU sweep; the positives are
// published before the select and needed after it: the exchange latency hides.
#include <mutex>
#include <vector>

#include "common.cuh"

GSSD_PHASE_DECL(fused)
#ifdef GSSD_PHASE_TIMING
// development: %globaltimer at entry and exit of every CTA (launch skew across the grid)
namespace gssd { __device__ unsigned long long g_fused_cta_ns[2][1024]; }
extern "C" __attribute__((visibility("default"))) int gssd_debug_fused_cta_ns(unsigned long long *out) {
    return (int)cudaMemcpyFromSymbol(out, gssd::g_fused_cta_ns, sizeof(unsigned long long) * 2 * 1024);
}
#define FUSED_CTA_STAMP(i) do { if (threadIdx.x == 0) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); \
        g_fused_cta_ns[i][blockIdx.y * gridDim.x + blockIdx.x] = _t; } } while (0)
#else
#define FUSED_CTA_STAMP(i) do { } while (0)
#endif

namespace gssd {

constexpr int FCHUNK = 256;          // priors per interleaved chunk
constexpr int FBINS = 2048;          // 11-bit digits
constexpr int FCAND = 256;           // composites that are ranked directly

struct FusedArgs {
    const float4 *loc; const float *conf; const float4 *priors;
    int B, P, C;
    const float *gt; const int32_t *gt_off;
    float threshold; int ratio; float var0, var1;
    FusedState *state;
    float *losses; float4 *grad_loc; float *grad_conf;
    uint8_t *pos_mask, *neg_mask; int32_t *num_pos;
    double *partials;                       // [2 * n_ctas]
    int S;                                  // CTAs per image (cluster size)
    int items;                              // priors of shared-memory storage per CTA (chunks per CTA * FCHUNK)
    XDev x;                                 // peer exchange (world == 0: off)
};

// 48-bit composite: mining key (ordered uint32) above, 0xffff - prior index below: larger = larger key, then LOWER index
__device__ __forceinline__ unsigned long long fcomp(uint32_t key, int p) {
    return ((unsigned long long)key << 16) | (unsigned long long)(0xffffu - (unsigned)p);
}
__device__ __forceinline__ int fpass_shift(int pass) { return pass == 0 ? 37 : pass == 1 ? 26 : pass == 2 ? 16 : pass == 3 ? 5 : 0; }
__device__ __forceinline__ int fpass_bits(int pass) { return pass == 2 ? 10 : pass == 4 ? 5 : 11; }

__device__ __forceinline__ float fsmooth_l1(float d, float &grad) {
    const float ad = fabsf(d);
    if (ad < 1.f) { grad = d; return __fmul_rn(__fmul_rn(0.5f, d), d); }
    grad = d > 0.f ? 1.f : -1.f;
    return __fsub_rn(ad, 0.5f);
}

// poll one word until it carries the epoch of this step; returns its value.  Every CTA of every rank is resident while we
// wait (cooperative launch), so only a rank that never reaches the criterion — or a state buffer shared by launches on two
// streams — can keep us here: after `timeout_ns` (0 = never) the kernel traps instead of wedging the GPU.
__device__ __forceinline__ uint32_t slot_wait(const unsigned long long *slot, uint32_t epoch, unsigned long long timeout_ns) {
    unsigned long long v, t0 = 0;
    unsigned spins = 0;
    while (true) {
        v = *reinterpret_cast<const volatile unsigned long long *>(slot);
        if ((uint32_t)(v >> 32) == epoch) break;
        if ((++spins & 0xff) == 0 && timeout_ns) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeout_ns) __trap();
        }
    }
    return (uint32_t)v;
}

// Registers are held to 64 per thread (two CTAs of 512 threads per SM): Detect runs beside the loss on a second stream with
// CTAs that fill an SM's register file, so the loss must be able to pack two CTAs onto each of the SMs Detect leaves free —
// otherwise the cooperative launch waits for Detect to finish and the two kernels serialise.
template <int NT, bool C2, bool GRADS>
__global__ void __launch_bounds__(NT, NT <= 512 ? 1024 / NT : 1) fused_kernel(FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned S = cluster.num_blocks();
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    constexpr int SLOTS = NT / FCHUNK;                           // chunks a CTA sweeps per trip
    const int g0 = a.gt_off[b];
    const int G = a.gt_off[b + 1] - g0;
    const int C = C2 ? 2 : a.C;
    const unsigned n_ctas = gridDim.x * gridDim.y;

    // ---- shared memory ------------------------------------------------------------------------------------------
    unsigned long long *sbest = reinterpret_cast<unsigned long long *>(smem_raw);
    const int Gp = (G + 1) & ~1;
    unsigned long long *sin = sbest + Gp;                                              // [S][Gp], clusters only
    float4 *sgt4 = reinterpret_cast<float4 *>(sin + (S > 1 ? S * Gp : 0));
    float *sarea = reinterpret_cast<float *>(sgt4 + G);
    float *slabel = sarea + G;
    int *sbp = reinterpret_cast<int *>(slabel + G);
    int *glist = sbp + G;
    uint32_t *hist = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(glist + G) + 15) & ~(uintptr_t)15);
    uint32_t *total = hist + FBINS;                                                    // [2][FBINS]
    unsigned long long *cand = reinterpret_cast<unsigned long long *>(total + 2 * FBINS);      // [S][FCAND]: one region per CTA of the image
    unsigned long long *flat = cand + (size_t)S * FCAND;                                       // [FCAND]
    uint32_t *keys = reinterpret_cast<uint32_t *>(flat + FCAND);
    float *conf_s = reinterpret_cast<float *>(keys + a.items);
    uint16_t *stag = reinterpret_cast<uint16_t *>(conf_s + (size_t)a.items * C);

    __shared__ int s_wi[NW];
    __shared__ float s_wf[NW];
    __shared__ float s_bbox[4][NW];
    __shared__ double s_red[2][NW];
    __shared__ int s_nlist;
    __shared__ uint32_t s_xmax_ord, s_digit, s_krem, s_eq, s_cand_count, s_wtot[NW];
    __shared__ int s_ntotal, s_npos_img, s_npos_cta;
    __shared__ uint32_t s_cand_n[8];
    __shared__ unsigned long long s_cut;

    FUSED_CTA_STAMP(0);
#ifdef GSSD_PHASE_TIMING
    if (a.ratio == -12345) return;                               // development: the launch + drain floor of this grid shape
#endif
    const int n_chunks = (a.P + FCHUNK - 1) / FCHUNK;
    const int my_chunks = (int)rank < n_chunks ? (n_chunks - (int)rank + (int)S - 1) / (int)S : 0;
    const int trips = (my_chunks + SLOTS - 1) / SLOTS;
    const int p_last = a.P - 1;
    const bool dbg = blockIdx.x == 0 && blockIdx.y == 0;
    GSSD_PHASE(fused, 0, dbg);

    // epoch of this step (read before anybody can have advanced it: it moves when the LAST CTA of the launch exits)
    const bool multi = a.x.world > 0;
    // two counters: the exchange buffer's (shared by the ranks' kernels of one consumer) tags the words that cross ranks, the
    // local state's tags everything that stays on this GPU (the state may serve several consumers on one stream)
    const uint32_t lepoch = *reinterpret_cast<const volatile uint32_t *>(&a.state->epoch) + 1;
    const uint32_t epoch = multi ? *reinterpret_cast<const volatile uint32_t *>(&a.x.peers[a.x.rank]->epoch) + 1 : lepoch;
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    // this CTA's word of rendezvous `kind` (0: conf max, 1: positives): one plain 8-byte store per destination
    auto publish = [&](int kind, uint32_t value) {
        const unsigned long long word = ((unsigned long long)epoch << 32) | value;
        if (!multi) {
            *reinterpret_cast<volatile unsigned long long *>(&a.state->slot[kind][cta]) = word;
        } else {
            for (int r = 0; r < a.x.world; ++r) {
                XBuf *dst = a.x.peers[r];
                if (kind == 0 && cta == 0)
                    *reinterpret_cast<volatile unsigned long long *>(&dst->hdr[epoch & 1][a.x.rank]) = ((unsigned long long)epoch << 32) | n_ctas;
                *reinterpret_cast<volatile unsigned long long *>(&dst->slot[epoch & 1][kind][a.x.rank][cta]) = word;
            }
        }
    };
    // all words of rendezvous `kind`, combined with MAX (kind 0) or SUM (kind 1) by the whole CTA; result in *s_out
    auto collect = [&](int kind, uint32_t *s_out) {
        uint32_t acc = 0;
        if (!multi) {
            for (unsigned i = tid; i < n_ctas; i += NT) {
                const uint32_t v = slot_wait(&a.state->slot[kind][i], epoch, 4000000000ull);
                acc = kind == 0 ? max(acc, v) : acc + v;
            }
        } else {
            // lane r of every warp learns how many CTAs rank r runs (all ranks polled at once: one L2 round trip, not `world` of
            // them in a row — that chain cost 3 us per rendezvous at 8 ranks), then the warp's threads share the world x CTAs words
            const XBuf *xl = a.x.peers[a.x.rank];
            const uint32_t n_mine = lane < a.x.world ? slot_wait(&xl->hdr[epoch & 1][lane], epoch, a.x.timeout_ns) : 0u;
            uint32_t incl = n_mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            for (uint32_t base = 0; base < total; base += NT) {                 // uniform trip count: the shuffles below need every lane
                const uint32_t j = base + tid;
                uint32_t r = 0, first = 0;
                for (int q = 0; q < a.x.world; ++q) {
                    const uint32_t end_q = __shfl_sync(FULL, incl, q), n_q = __shfl_sync(FULL, n_mine, q);
                    if (j >= end_q) { r = q + 1; } else if (j >= end_q - n_q) { r = q; first = end_q - n_q; }
                }
                if (j < total) {
                    const uint32_t v = slot_wait(&xl->slot[epoch & 1][kind][r][j - first], epoch, a.x.timeout_ns);
                    acc = kind == 0 ? max(acc, v) : acc + v;
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const uint32_t t = __shfl_xor_sync(FULL, acc, o);
            acc = kind == 0 ? max(acc, t) : acc + t;
        }
        if (lane == 0) s_wtot[warp] = acc;
        __syncthreads();
        acc = lane < NW ? s_wtot[lane] : 0u;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const uint32_t t = __shfl_xor_sync(FULL, acc, o);
            acc = kind == 0 ? max(acc, t) : acc + t;
        }
        if (tid == 0) *s_out = acc;
        __syncthreads();
    };

    // ---- A. conf rows -> shared memory, asynchronously; buffers of the select cleared; GT staged -------------------------
    for (int cj = 0; cj < my_chunks; ++cj) {
        const int ck = (int)rank + cj * (int)S;
        const int n_pr = min(FCHUNK, a.P - ck * FCHUNK);
        const int nf = n_pr * C;
        const float *src = a.conf + ((size_t)b * a.P + (size_t)ck * FCHUNK) * C;
        float *dst = conf_s + (size_t)cj * FCHUNK * C;
        const bool vec = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
        const int n4 = vec ? nf >> 2 : 0;
        for (int i = tid; i < n4; i += NT) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 4 * i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 4 * i) : "memory");
        }
        for (int i = 4 * n4 + tid; i < nf; i += NT) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst + i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src + i) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < 3 * FBINS; i += NT) hist[i] = 0;       // hist + both totals
    if (tid == 0) { s_cand_count = 0; s_npos_img = 0; s_nlist = G; }
    for (int g = tid; g < G; g += NT) {
        const float *row = a.gt + 5 * (size_t)(g0 + g);
        const float4 t = make_float4(row[0], row[1], row[2], row[3]);
        sgt4[g] = t;
        sarea[g] = box_area(t);
        slabel[g] = row[4];
        sbest[g] = 0x00000000ffffffffull;                        // an all-zero IoU row resolves to prior 0 (torch.max: first maximum)
        glist[g] = g;
    }
    GSSD_PHASE(fused, 12, dbg);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    GSSD_PHASE(fused, 13, dbg);
    if (S > 1) cluster_arrive();
    GSSD_PHASE(fused, 14, dbg);         // my receiving buffers are initialised; waited for before the first remote store

    // local max of conf -> batch max (box_utils.py:167): this CTA's word of rendezvous 1.  Only the last chunk of the image can
    // be partial; the full chunks are one contiguous run of floats in shared memory.
    {
        const bool owns_tail = my_chunks > 0 && (int)rank + (my_chunks - 1) * (int)S == n_chunks - 1 && (a.P % FCHUNK) != 0;
        const int n_full4 = (my_chunks - (owns_tail ? 1 : 0)) * FCHUNK * C / 4;              // FCHUNK * C is a multiple of 4
        const float4 *src4 = reinterpret_cast<const float4 *>(conf_s);
        float m0 = -INFINITY, m1 = -INFINITY;
        int i = tid;
        for (; i + NT < n_full4; i += 2 * NT) {
            const float4 u = src4[i], v = src4[i + NT];
            m0 = fmaxf(m0, fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w)));
            m1 = fmaxf(m1, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
        if (i < n_full4) { const float4 u = src4[i]; m0 = fmaxf(m0, fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w))); }
        if (owns_tail) {
            const float *src = conf_s + (size_t)(my_chunks - 1) * FCHUNK * C;
            const int nf = (a.P % FCHUNK) * C;
            for (int q = tid; q < nf; q += NT) m1 = fmaxf(m1, src[q]);
        }
        float cmax = warp_max(fmaxf(m0, m1));
        if (lane == 0) s_wf[warp] = cmax;
        __syncthreads();
        if (warp == 0) {
            cmax = warp_max(lane < NW ? s_wf[lane] : -INFINITY);
            if (lane == 0) publish(0, my_chunks > 0 ? f2ord(cmax) : 0u);
        }
    }
    GSSD_PHASE(fused, 1, dbg);

    // ---- B. IoU sweep (box_utils.py:88-96), as match.cu -----------------------------------------------------------------
    // optional: drop the GT boxes that cannot touch this CTA's priors
    if (G >= 8) {
        float bx1 = INFINITY, by1 = INFINITY, bx2 = -INFINITY, by2 = -INFINITY;
        for (int t = 0; t < trips; ++t) {
            const int cj = t * SLOTS + (tid >> 8);
            const int p = ((int)rank + cj * (int)S) * FCHUNK + (tid & 255);
            if (cj >= my_chunks || p >= a.P) continue;
            const float4 pb = point_form(a.priors[p]);
            bx1 = fminf(bx1, pb.x); by1 = fminf(by1, pb.y); bx2 = fmaxf(bx2, pb.z); by2 = fmaxf(by2, pb.w);
        }
        bx1 = warp_min(bx1); by1 = warp_min(by1); bx2 = warp_max(bx2); by2 = warp_max(by2);
        if (lane == 0) { s_bbox[0][warp] = bx1; s_bbox[1][warp] = by1; s_bbox[2][warp] = bx2; s_bbox[3][warp] = by2; }
        __syncthreads();
        if (warp == 0) {
            for (int w = 0; w < NW; ++w) {
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
    const bool warp_cull = n_list >= 8;
    auto prior_of = [&](int t) {                                 // clamped: a trip past the end re-reads a cached row
        const int cj = min(t * SLOTS + (tid >> 8), max(my_chunks - 1, 0));
        return a.priors[min(((int)rank + cj * (int)S) * FCHUNK + (tid & 255), p_last)];
    };
    if (G < 8) {
        // few GT boxes (the training case: 1-5): a thread's priors of up to five trips are swept TOGETHER, one GT box at a time —
        // five independent dependency chains per thread instead of one (the sweep was latency-bound: 7.0 k cycles at batch 32)
        constexpr int Q = 5;
        for (int t0 = 0; t0 < trips; t0 += Q) {
            float4 pb[Q];
            float ar[Q], best[Q];
            int bidx[Q], pcs[Q];
            bool on[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int cj = (t0 + q) * SLOTS + (tid >> 8);    // warp-uniform
                on[q] = t0 + q < trips && cj < my_chunks;
                pcs[q] = min(((int)rank + min(cj, max(my_chunks - 1, 0)) * (int)S) * FCHUNK + (tid & 255), p_last);
                pb[q] = point_form(a.priors[pcs[q]]);            // Q loads in flight; a lane past the end repeats the last prior
                ar[q] = box_area(pb[q]);
                best[q] = 0.f; bidx[q] = 0;                      // IoU >= 0: row 0 wins an all-zero column
            }
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                const float4 tg = sgt4[g];
                const float ta = sarea[g];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    if (!on[q]) continue;
                    const float iw = __fsub_rn(fminf(tg.z, pb[q].z), fmaxf(tg.x, pb[q].x));
                    const float ih = __fsub_rn(fminf(tg.w, pb[q].w), fmaxf(tg.y, pb[q].y));
                    if (iw > 0.f && ih > 0.f) {                  // the boxes overlap
                        const float inter = __fmul_rn(iw, ih);
                        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ta, ar[q]), inter));
                        if (iou > best[q]) { best[q] = iou; bidx[q] = g; }       // first max over GT (torch.max dim 0)
                        const unsigned bits = __float_as_uint(iou);
                        const unsigned long long key = ((unsigned long long)bits << 32) | (0xffffffffu - (unsigned)pcs[q]);
                        if (key > sbest[g]) {                    // best prior of this GT: max of (IoU bits, ~prior)
                            const unsigned act = __activemask();
                            const unsigned m = __reduce_max_sync(act, bits);
                            const unsigned who = __ballot_sync(act, bits == m);
                            if (lane == __ffs(who) - 1) atomicMax(&sbest[g], key);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q)
                if (on[q])
                    stag[((t0 + q) * SLOTS + (tid >> 8)) * FCHUNK + (tid & 255)] =
                        (uint16_t)(bidx[q] | (!(best[q] < a.threshold) ? 0x8000 : 0));                 // box_utils.py:108
        }
    }
    float4 nxt = prior_of(0);
    for (int t = 0; G >= 8 && t < trips; ++t) {
        const int cj = t * SLOTS + (tid >> 8);                   // warp-uniform
        const float4 cur = nxt;
        nxt = prior_of(t + 1);                                   // in flight while this trip's pairs are swept
        if (cj >= my_chunks) continue;
        const int p = ((int)rank + cj * (int)S) * FCHUNK + (tid & 255);
        const int pc = min(p, p_last);                           // a lane past the end repeats the last prior (its keys are ignored)
        const float4 pb = point_form(cur);
        const float area_b = box_area(pb);
        float best = 0.f;                                        // IoU >= 0: row 0 wins an all-zero column
        int bidx = 0;
        auto sweep_one = [&](int g) {
            const float4 tg = sgt4[g];
            const float iw = __fsub_rn(fminf(tg.z, pb.z), fmaxf(tg.x, pb.x));
            const float ih = __fsub_rn(fminf(tg.w, pb.w), fmaxf(tg.y, pb.y));
            if (iw > 0.f && ih > 0.f) {                          // the boxes overlap
                const float inter = __fmul_rn(iw, ih);
                const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sarea[g], area_b), inter));
                if (iou > best) { best = iou; bidx = g; }        // first max over GT (torch.max dim 0)
                const unsigned bits = __float_as_uint(iou);
                const unsigned long long key = ((unsigned long long)bits << 32) | (0xffffffffu - (unsigned)pc);
                if (key > sbest[g]) {                            // best prior of this GT: max of (IoU bits, ~prior)
                    const unsigned act = __activemask();
                    const unsigned m = __reduce_max_sync(act, bits);
                    const unsigned who = __ballot_sync(act, bits == m);
                    if (lane == __ffs(who) - 1) atomicMax(&sbest[g], key);
                }
            }
        };
        if (!warp_cull) {
            for (int q = 0; q < n_list; ++q) sweep_one(glist[q]);
        } else {
            float bx1 = pb.x, by1 = pb.y, bx2 = pb.z, by2 = pb.w;
            bx1 = warp_min(bx1); by1 = warp_min(by1); bx2 = warp_max(bx2); by2 = warp_max(by2);
            for (int gb = 0; gb < n_list; gb += 32) {
                const int q = gb + lane;
                bool hit = false;
                if (q < n_list) {
                    const float4 tg = sgt4[glist[q]];
                    hit = fminf(tg.z, bx2) > fmaxf(tg.x, bx1) && fminf(tg.w, by2) > fmaxf(tg.y, by1);
                }
                unsigned m = __ballot_sync(FULL, hit);
                while (m) {                                      // ascending GT row: first-max order is kept
                    const int qq = gb + __ffs(m) - 1;
                    m &= m - 1;
                    sweep_one(glist[qq]);
                }
            }
        }
        stag[cj * FCHUNK + (tid & 255)] = (uint16_t)(bidx | (!(best < a.threshold) ? 0x8000 : 0));   // box_utils.py:108
    }
    __syncthreads();
    GSSD_PHASE(fused, 2, dbg);

    // best prior per GT over the whole image, then the sequential force match (box_utils.py:101-105)
    if (S > 1) {
        cluster_wait();                                          // every CTA of the image is running, its buffers are initialised
        for (unsigned i = tid; i < (unsigned)G * S; i += NT) {
            const unsigned g = i / S, r = i - g * S;
            if (r != rank) cluster.map_shared_rank(sin, r)[rank * Gp + g] = sbest[g];
        }
        cluster.sync();
        for (int g = tid; g < G; g += NT) {
            unsigned long long m = sbest[g];
            for (unsigned r = 0; r < S; ++r)
                if (r != rank) m = max(m, sin[r * Gp + g]);
            sbest[g] = m;
        }
        __syncthreads();
    }
    for (int g = tid; g < G; g += NT) sbp[g] = (int)(0xffffffffu - (unsigned)(sbest[g] & 0xffffffffu));
    __syncthreads();
    if (tid == 0) {
        for (int g = 0; g < G; ++g) {                            // in GT order: the last GT wins a shared best prior
            const int bp = sbp[g];
            const int ck = bp / FCHUNK;
            if (ck % (int)S == (int)rank) stag[(ck / (int)S) * FCHUNK + bp % FCHUNK] = (uint16_t)(0x8000 | g);
        }
    }
    __syncthreads();

    // ---- C. positives of this CTA -> N; arrive at rendezvous 2 ----------------------------------------------------------
    {
        int npos = 0;
        for (int li = tid; li < my_chunks * FCHUNK; li += NT) {
            const int p = ((int)rank + (li >> 8) * (int)S) * FCHUNK + (li & 255);
            npos += (p < a.P) && (stag[li] & 0x8000);
        }
        npos = warp_sum(npos);
        if (lane == 0) s_wi[warp] = npos;
        __syncthreads();
        if (warp == 0) {
            const int tot = warp_sum(lane < NW ? s_wi[lane] : 0);
            if (lane == 0) {
                s_npos_cta = tot;
                if (S == 1) s_npos_img = tot;
                publish(1, (uint32_t)tot);
            }
        }
    }
    GSSD_PHASE(fused, 3, dbg);

    // ---- D. the batch max of conf, then the mining keys (multibox_loss.py:91-99) ---------------------------------------------
    collect(0, &s_xmax_ord);
    const float x_max = ord2f(s_xmax_ord);
    for (int li = tid; li < my_chunks * FCHUNK; li += NT) {
        const int p = ((int)rank + (li >> 8) * (int)S) * FCHUNK + (li & 255);
        if (p >= a.P) { keys[li] = 0u; continue; }
        float key = 0.f;
        if (!(stag[li] & 0x8000)) {                              // loss_c[pos] = 0
            if (C2) {
                const float2 x = *reinterpret_cast<const float2 *>(conf_s + 2 * li);
                const float s = __fadd_rn(expf(__fsub_rn(x.x, x_max)), expf(__fsub_rn(x.y, x_max)));
                key = __fsub_rn(__fadd_rn(logf(s), x_max), x.x);
            } else {
                const float *row = conf_s + (size_t)li * C;
                float s = 0.f;
                for (int c = 0; c < C; ++c) s = __fadd_rn(s, expf(__fsub_rn(row[c], x_max)));
                key = __fsub_rn(__fadd_rn(logf(s), x_max), row[0]);
            }
            key = __fadd_rn(key, 0.f);                           // -0 -> +0
        }
        const uint32_t ko = f2ord(key);
        keys[li] = ko;
        atomicAdd(&hist[ko >> 21], 1u);
    }
    __syncthreads();
    GSSD_PHASE(fused, 4, dbg);

    // ---- E. hard-negative selection (multibox_loss.py:102-106) ------------------------------------------------------------------
    if (S > 1) {
        for (int i = tid; i < FBINS; i += NT) {
            const uint32_t c = hist[i];
            if (c)
                for (unsigned r = 0; r < S; ++r) atomicAdd(&cluster.map_shared_rank(total, r)[i], c);
        }
        if (tid == 0)
            for (unsigned r = 0; r < S; ++r) atomicAdd(cluster.map_shared_rank(&s_npos_img, r), s_npos_cta);
        cluster.sync();
    }
    GSSD_PHASE(fused, 8, dbg);
    const int num_pos = s_npos_img;
    long long k_ll = (long long)a.ratio * num_pos;
    if (k_ll > a.P - 1) k_ll = a.P - 1;
    const bool have_sel = k_ll > 0;
    unsigned long long cut = ~0ull;
    if (have_sel) {
        uint32_t k_rem = (uint32_t)k_ll, eq = 0;
        unsigned long long prefix = 0;
        int pass = 0;
        while (true) {
            // the bin that holds the k_rem-th largest composite: suffix sums over the bins, a contiguous run of bins per thread
            const uint32_t *tot = hist + (S > 1 ? (1 + (pass & 1)) * FBINS : 0);          // total = hist + FBINS: one shared base
            const int nb = 1 << fpass_bits(pass);
            constexpr int BPT = FBINS / NT > 0 ? FBINS / NT : 1;
            uint32_t mine = 0;
            if (tid * BPT < nb)
#pragma unroll
                for (int q = 0; q < BPT; ++q) mine += tot[tid * BPT + q];
            uint32_t incl = mine;                                // inclusive suffix within the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_down_sync(FULL, incl, o);
                if (lane + o < 32) incl += t;
            }
            if (lane == 0) s_wtot[warp] = incl;
            __syncthreads();
            uint32_t above = (lane > warp && lane < NW) ? s_wtot[lane] : 0u;           // composites in the runs of the warps above
#pragma unroll
            for (int o = 16; o; o >>= 1) above += __shfl_xor_sync(FULL, above, o);
            uint32_t run = above + incl - mine;                  // composites in bins above this thread's run
            if (tid * BPT < nb && run < k_rem && k_rem <= run + mine) {
#pragma unroll
                for (int q = BPT - 1; q >= 0; --q) {
                    const uint32_t c = tot[tid * BPT + q];
                    if (run < k_rem && k_rem <= run + c) { s_digit = tid * BPT + q; s_krem = k_rem - run; s_eq = c; }
                    run += c;
                }
            }
            __syncthreads();
            prefix = (prefix << fpass_bits(pass)) | s_digit;
            k_rem = s_krem; eq = s_eq;
            GSSD_PHASE(fused, 9, dbg);
            if (eq <= (uint32_t)FCAND || pass == 4) break;
            ++pass;
            // next digit of the composites that share the prefix
            for (int i = tid; i < FBINS; i += NT) { hist[i] = 0; if (S > 1) total[((pass + 1) & 1) * FBINS + i] = 0; }
            __syncthreads();
            const int sh = fpass_shift(pass), bits = fpass_bits(pass);
            for (int li = tid; li < my_chunks * FCHUNK; li += NT) {
                const int p = ((int)rank + (li >> 8) * (int)S) * FCHUNK + (li & 255);
                if (p >= a.P) continue;
                const unsigned long long c = fcomp(keys[li], p);
                if ((c >> (sh + bits)) == prefix) atomicAdd(&hist[(uint32_t)(c >> sh) & ((1u << bits) - 1u)], 1u);
            }
            __syncthreads();
            if (S > 1) {
                for (int i = tid; i < (1 << bits); i += NT) {
                    const uint32_t c = hist[i];
                    if (c)
                        for (unsigned r = 0; r < S; ++r) atomicAdd(&cluster.map_shared_rank(total, r)[(pass & 1) * FBINS + i], c);
                }
                cluster.sync();
            }
        }
        // the members of the cut bin: compacted locally (shared-memory atomics), then copied into this CTA's region of every
        // CTA of the image with plain stores — a remote atomic that returns a slot costs a round trip per candidate
        const int sh = fpass_shift(pass);
        unsigned long long *mine = cand + (size_t)rank * FCAND;
        for (int li = tid; li < my_chunks * FCHUNK; li += NT) {
            const int p = ((int)rank + (li >> 8) * (int)S) * FCHUNK + (li & 255);
            if (p >= a.P) continue;
            const unsigned long long c = fcomp(keys[li], p);
            if ((c >> sh) == prefix) mine[atomicAdd(&s_cand_count, 1u)] = c;
        }
        __syncthreads();
        GSSD_PHASE(fused, 10, dbg);
        const uint32_t n_mine = s_cand_count;
        if (S > 1) {
            for (unsigned i = tid; i < n_mine * S; i += NT) {
                const unsigned r = i / n_mine, j = i - r * n_mine;
                if (r != rank) cluster.map_shared_rank(cand, r)[(size_t)rank * FCAND + j] = mine[j];
            }
            if (tid < S && tid != rank) cluster.map_shared_rank(s_cand_n, tid)[rank] = n_mine;
            if (tid == 0) s_cand_n[rank] = n_mine;
            cluster.sync();
            // flatten: region r holds s_cand_n[r] composites; eq of them in total
            if (tid < (int)eq) {
                int r = 0, j = tid;
                while (j >= (int)s_cand_n[r]) { j -= (int)s_cand_n[r]; ++r; }
                flat[tid] = cand[(size_t)r * FCAND + j];
            }
            __syncthreads();
        }
        GSSD_PHASE(fused, 11, dbg);
        const unsigned long long *list = cand + (S > 1 ? (size_t)S * FCAND : (size_t)rank * FCAND);   // flat or mine: one shared base
        {
            int R = 1;                                           // threads per candidate: as many as the CTA has to spare
            while (R < 32 && 2 * R * (int)eq <= NT) R <<= 1;
            const int ci = tid / R, part = tid % R;
            const unsigned long long c = ci < (int)eq ? list[ci] : 0ull;
            uint32_t cnt = 0;
            for (int j = part; j < (int)eq; j += R) cnt += list[j] > c;
            for (int o = R >> 1; o; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
            if (ci < (int)eq && part == 0 && cnt == k_rem - 1) s_cut = c;
        }
        __syncthreads();
        cut = s_cut;
    }
    GSSD_PHASE(fused, 5, dbg);

    // ---- F. N, then smooth-L1 (80-88), cross-entropy over pos | neg (108-113) and every gradient ----------------------------------------
    collect(1, reinterpret_cast<uint32_t *>(&s_ntotal));
    const int n_total = s_ntotal;
    const float n_f = (float)n_total;
    double acc_l = 0.0, acc_c = 0.0;
    for (int li = tid; li < my_chunks * FCHUNK; li += NT) {
        const int p = ((int)rank + (li >> 8) * (int)S) * FCHUNK + (li & 255);
        if (p >= a.P) continue;
        const size_t o = (size_t)b * a.P + p;
        const uint16_t tag = stag[li];
        const bool pos = tag & 0x8000;
        const bool neg = have_sel && fcomp(keys[li], p) >= cut;
        if (a.pos_mask) a.pos_mask[o] = pos;
        if (a.neg_mask) a.neg_mask[o] = neg;
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pos) {
            const float4 lt = encode_box(sgt4[tag & 0x7fff], a.priors[p], a.var0, a.var1);   // box_utils.py:114-135, never materialised
            const float4 l = ldg_stream(a.loc + o);
            const float l0 = fsmooth_l1(__fsub_rn(l.x, lt.x), g4.x), l1 = fsmooth_l1(__fsub_rn(l.y, lt.y), g4.y);
            const float l2 = fsmooth_l1(__fsub_rn(l.z, lt.z), g4.z), l3 = fsmooth_l1(__fsub_rn(l.w, lt.w), g4.w);
            acc_l += (double)l0 + (double)l1 + (double)l2 + (double)l3;
            g4.x = __fdiv_rn(g4.x, n_f); g4.y = __fdiv_rn(g4.y, n_f); g4.z = __fdiv_rn(g4.z, n_f); g4.w = __fdiv_rn(g4.w, n_f);
        }
        if (GRADS) __stcs(&a.grad_loc[o], g4);
        if (!(pos || neg)) {
            if (GRADS) {
                if (C2) __stcs(reinterpret_cast<float2 *>(a.grad_conf + o * 2), make_float2(0.f, 0.f));
                else for (int c = 0; c < C; ++c) a.grad_conf[o * C + c] = 0.f;
            }
            continue;
        }
        const int t = pos ? (int)__fadd_rn(slabel[tag & 0x7fff], 1.f) : 0;                  // box_utils.py:107
        if (C2) {
            const float2 x = *reinterpret_cast<const float2 *>(conf_s + 2 * li);
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
            const float *row = conf_s + (size_t)li * C;
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
    if (a.num_pos && rank == 0 && tid == 0) a.num_pos[b] = num_pos;
    GSSD_PHASE(fused, 6, dbg);

    // ---- G. finish -------------------------------------------------------------------------------------------------------------
    // Every CTA leaves its two partial sums (reduced in double, rounded once to float) in its own tagged words; CTA 0 of the
    // grid waits for all of them, adds them in CTA order (deterministic) in double and divides by N.  No fence, no atomic,
    // and nobody but CTA 0 waits.
    acc_l = warp_sum(acc_l); acc_c = warp_sum(acc_c);
    if (lane == 0) { s_red[0][warp] = acc_l; s_red[1][warp] = acc_c; }
    __syncthreads();
    if (tid == 0) {
        double l = 0.0, c = 0.0;
        for (int w = 0; w < NW; ++w) { l += s_red[0][w]; c += s_red[1][w]; }
        *reinterpret_cast<volatile unsigned long long *>(&a.state->slot[2][cta]) = ((unsigned long long)lepoch << 32) | __float_as_uint((float)l);
        *reinterpret_cast<volatile unsigned long long *>(&a.state->slot[3][cta]) = ((unsigned long long)lepoch << 32) | __float_as_uint((float)c);
    }
    if (cta == 0) {
        double l = 0.0, c = 0.0;
        for (unsigned i = tid; i < n_ctas; i += NT) {
            l += (double)__uint_as_float(slot_wait(&a.state->slot[2][i], lepoch, 4000000000ull));
            c += (double)__uint_as_float(slot_wait(&a.state->slot[3][i], lepoch, 4000000000ull));
        }
        l = warp_sum(l); c = warp_sum(c);
        __syncthreads();
        if (lane == 0) { s_red[0][warp] = l; s_red[1][warp] = c; }
        __syncthreads();
        if (tid == 0) {
            l = 0.0; c = 0.0;
            for (int w = 0; w < NW; ++w) { l += s_red[0][w]; c += s_red[1][w]; }
            a.losses[0] = __fdiv_rn((float)l, n_f);              // multibox_loss.py:117-119
            a.losses[1] = __fdiv_rn((float)c, n_f);
            // every CTA has published its last word: the next launch on this state is a new epoch
            if (multi) *reinterpret_cast<volatile uint32_t *>(&a.x.peers[a.x.rank]->epoch) = epoch;
            *reinterpret_cast<volatile uint32_t *>(&a.state->epoch) = lepoch;
        }
    }
    GSSD_PHASE(fused, 7, dbg);
    FUSED_CTA_STAMP(1);
}

static size_t fused_smem_bytes(int g_max, int S, int items, int C) {
    const size_t Gp = (size_t)((g_max + 1) & ~1);
    size_t b = Gp * 8 + (S > 1 ? (size_t)S * Gp * 8 : 0) + (size_t)g_max * (16 + 4 + 4 + 4 + 4) + 16;
    b += (size_t)3 * FBINS * 4 + (size_t)(S + 1) * FCAND * 8;
    b += (size_t)items * (4 + 4 * (size_t)C + 2) + 16;
    return b;
}

struct FusedPlan { int S, NT, items; size_t smem; };

template <int NT, bool C2, bool GR>
static const void *fused_fn() { return reinterpret_cast<const void *>(fused_kernel<NT, C2, GR>); }

static const void *fused_pick(int NT, bool c2, bool gr) {
    if (NT == 256) return c2 ? (gr ? fused_fn<256, true, true>() : fused_fn<256, true, false>()) : (gr ? fused_fn<256, false, true>() : fused_fn<256, false, false>());
    if (NT == 1024) return c2 ? (gr ? fused_fn<1024, true, true>() : fused_fn<1024, true, false>()) : (gr ? fused_fn<1024, false, true>() : fused_fn<1024, false, false>());
    return c2 ? (gr ? fused_fn<512, true, true>() : fused_fn<512, true, false>()) : (gr ? fused_fn<512, false, true>() : fused_fn<512, false, false>());
}

static FusedPlan fused_plan_uncached(int B, int P, int C, int g_max, const void **kern_out, bool c2, bool gr);

// the launch shape for which every CTA of the batch is resident at once, or S == 0 when there is none (memoised: the
// occupancy query is far slower than a launch)
static FusedPlan fused_plan(int B, int P, int C, int g_max, const void **kern_out, bool c2, bool gr) {
    struct Entry { int dev, B, P, C, g, gr; FusedPlan pl; const void *kern; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const int gb = (g_max + 7) & ~7;                             // shared memory grows with g_max: bucket it
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry &e : cache)
        if (e.dev == dev && e.B == B && e.P == P && e.C == C && e.g == gb && e.gr == (int)gr) { *kern_out = e.kern; return e.pl; }
    const void *kern = nullptr;
    const FusedPlan pl = fused_plan_uncached(B, P, C, gb, &kern, c2, gr);
    if (cache.size() < 4096) cache.push_back(Entry{dev, B, P, C, gb, (int)gr, pl, kern});
    *kern_out = kern;
    return pl;
}

static FusedPlan fused_plan_uncached(int B, int P, int C, int g_max, const void **kern_out, bool c2, bool gr) {
    FusedPlan pl = {0, 0, 0, 0};
    static const int forced_s = []{ const char *e = getenv("GSSD_FUSED_S"); return e ? atoi(e) : 0; }();
    static const int forced_nt = []{ const char *e = getenv("GSSD_FUSED_NT"); return e ? atoi(e) : 0; }();
    static const int off = []{ const char *e = getenv("GSSD_FUSED"); return e && atoi(e) == 0 ? 1 : 0; }();
    if (off) return pl;
    int dev = 0, sms = 148, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int n_chunks = ceil_div(P, FCHUNK);
    const int NT = (forced_nt == 256 || forced_nt == 512 || forced_nt == 1024) ? forced_nt : 512;
    // first choice: one CTA per SM (every phase at full speed), the widest cluster that allows it; second: two CTAs per SM;
    // last: whatever the occupancy calculator admits with one CTA per image
    for (int round = 0; round < 3; ++round) {
        for (int S = 8; S >= 1; S >>= 1) {
            if (forced_s && S != forced_s) continue;
            if (S > 1 && S > n_chunks) continue;
            if (!forced_s) {
                if (round == 0 && (long)B * S > sms) continue;
                if (round == 1 && (long)B * S > 2l * sms) continue;
                if (round == 2 && S > 1) continue;
            }
            const int items = ceil_div(n_chunks, S) * FCHUNK;
            // a CTA that owns more than 24 chunks has too few threads for its IoU sweep: measured, SSD512 priors with up to 32 GT at
            // batch 64 (S = 2, 48 chunks per CTA) take 94 us in one launch against 41 + 39 us in two, SSD300 priors at batch 64 (S = 2,
            // 18 chunks) 32 us against 16 + 23 us
            if (!forced_s && items > 24 * FCHUNK) continue;
            const size_t smem = fused_smem_bytes(g_max, S, items, C);
            if (smem > (size_t)optin - 2048) continue;
            const void *kern = fused_pick(NT, c2, gr);
            if (allow_max_smem(kern) != cudaSuccess) { cudaGetLastError(); continue; }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(S, B, 1);
            cfg.blockDim = dim3(NT, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
            if (clusters < B || (long)B * S > GSSD_FUSED_MAX_CTAS) continue;
            pl.S = S; pl.NT = NT; pl.items = items; pl.smem = smem;
            *kern_out = kern;
            return pl;
        }
    }
    return pl;
}

}  // namespace gssd

using namespace gssd;

extern "C" size_t gssd_fused_state_bytes(void) { return sizeof(FusedState); }

extern "C" int gssd_mbox_fused_supported(int B, int P, int C, int g_max) {
    if (B <= 0 || P <= 0 || C < 2 || C > GSSD_MAX_CLASSES || g_max <= 0 || g_max > GSSD_MAX_GT_PER_IMAGE || P > GSSD_MAX_PRIORS) return 0;
    const void *kern = nullptr;
    return fused_plan(B, P, C, g_max, &kern, C == 2, true).S > 0 ? 1 : 0;
}

extern "C" int gssd_mbox_loss_fused(const float *loc, const float *conf, const float *priors, int B, int P, int C,
                                    const float *gt, const int32_t *gt_off, int sum_G, int g_max,
                                    float threshold, int negpos_ratio, float var0, float var1,
                                    void *state, const gssd_xchg *x,
                                    float *losses, float *grad_loc, float *grad_conf,
                                    uint8_t *pos_mask, uint8_t *neg_mask, int32_t *num_pos,
                                    void *ws, size_t ws_bytes, void *stream) {
    if (x && (x->world < 1 || x->world > GSSD_XCHG_MAX_RANKS || x->rank < 0 || x->rank >= x->world)) return GSSD_ERR_ARG;
    if (!loc || !conf || !priors || !gt || !gt_off || !state || !losses || !ws) return GSSD_ERR_ARG;
#ifdef GSSD_PHASE_TIMING
    const bool floor_probe = negpos_ratio == -12345;
#else
    const bool floor_probe = false;
#endif
    if (B <= 0 || P <= 0 || sum_G <= 0 || g_max <= 0 || (negpos_ratio < 0 && !floor_probe)) return sum_G <= 0 && B > 0 ? GSSD_ERR_EMPTY : GSSD_ERR_ARG;
    if (C < 2 || C > GSSD_MAX_CLASSES) return GSSD_ERR_ARG;
    if ((grad_loc == nullptr) != (grad_conf == nullptr)) return GSSD_ERR_ARG;
    if (g_max > GSSD_MAX_GT_PER_IMAGE || P > GSSD_MAX_PRIORS) return GSSD_ERR_LIMIT;
    if (ws_bytes < gssd_workspace_bytes(GSSD_WS_LOSS, B, P, C, sum_G, 0)) return GSSD_ERR_WS;
    const bool gr = grad_loc != nullptr, c2 = C == 2;
    const void *kern = nullptr;
    const FusedPlan pl = fused_plan(B, P, C, g_max, &kern, c2, gr);
    if (pl.S == 0) return GSSD_ERR_UNSUPPORTED;
    if (pl.smem < fused_smem_bytes(g_max, pl.S, pl.items, C) || (long)B * pl.S > GSSD_FUSED_MAX_CTAS) return GSSD_ERR_UNSUPPORTED;
    FusedArgs a = {};
    a.loc = reinterpret_cast<const float4 *>(loc); a.conf = conf; a.priors = reinterpret_cast<const float4 *>(priors);
    a.B = B; a.P = P; a.C = C; a.gt = gt; a.gt_off = gt_off;
    a.threshold = threshold; a.ratio = negpos_ratio; a.var0 = var0; a.var1 = var1;
    a.state = reinterpret_cast<FusedState *>(state);
    a.losses = losses; a.grad_loc = reinterpret_cast<float4 *>(grad_loc); a.grad_conf = grad_conf;
    a.pos_mask = pos_mask; a.neg_mask = neg_mask; a.num_pos = num_pos;
    a.partials = reinterpret_cast<double *>(ws);
    a.S = pl.S; a.items = pl.items;
    a.x = xdev_from(x);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.S, B, 1);
    cfg.blockDim = dim3(pl.NT, 1, 1);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;                 // every CTA resident at once: the rendezvous cannot deadlock
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    void *params[] = {&a};
    static int coop_ok = []{ const char *e = getenv("GSSD_FUSED_COOP"); return e && atoi(e) == 0 ? 0 : 1; }();   // cluster + cooperative in one launch: dropped if the runtime refuses
    if (coop_ok) {
        cfg.numAttrs = 2;
        cudaError_t e = cudaLaunchKernelExC(&cfg, kern, params);
        if (e == cudaSuccess) { GSSD_AFTER_LAUNCH(); return GSSD_OK; }
        cudaGetLastError();
        if (e != cudaErrorNotSupported && e != cudaErrorInvalidValue && e != cudaErrorCooperativeLaunchTooLarge) return (int)e;
        coop_ok = 0;
    }
    cfg.numAttrs = 1;                                            // residency was checked with cudaOccupancyMaxActiveClusters
    GSSD_RETURN_IF_CUDA(cudaLaunchKernelExC(&cfg, kern, params));
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}
