// fused.cu — MultiBoxLoss.forward and its backward as ONE launch (multibox_loss.py:46-120 with the matching loop of
// lines 67-72 -> box_utils.match, box_utils.py:70-111, inside).
//
// Why one launch: at the batch the reference trains with (32 images) the whole problem is a few megabytes; two launches
// (match.cu then loss.cu) spend more time starting, draining and handing 16 bytes of statistics through global memory than
// moving data.  The only batch-wide dependencies of the loss are two scalars — the max of conf (box_utils.py:167) and the
// number of positives N (multibox_loss.py:117) — so all CTAs of the batch are made co-resident (cooperative launch) and
// meet at two counters in global memory; everything else is per image and stays inside one thread-block cluster.
//
// Per image: a cluster of S CTAs of NT threads; the priors are dealt to the CTAs in interleaved chunks of 256 (CTA r owns
// chunks r, r+S, ...: the large priors at the end of the list overlap most GT boxes, contiguous slices would leave the first
// CTAs waiting).  A thread owns the same priors in every phase, and the CTA keeps their conf rows, tags and mining keys in
// shared memory: conf is read from HBM exactly once.
//
//   A  conf rows -> shared memory (cp.async), local max -> atomicMax; arrive at rendezvous 1
//   B  IoU sweep (as match.cu), per-GT best prior combined over the cluster (1 exchange), sequential force match
//   C  positives counted -> atomicAdd N; arrive at rendezvous 2
//   D  wait for rendezvous 1 (long complete): mining keys with the batch-global max, first radix digit (11 bits) counted
//      on the fly
//   E  select of the min(ratio*num_pos, P-1) largest (key, lower index first) composites: the non-empty bins of every CTA
//      are pushed into every CTA's totals (1 exchange); the members of the bin that holds the cut (<= 256, typically
//      ~100) are pushed to every CTA (1 exchange) and ranked there by counting; more passes only when more than 256
//      composites share 11 / 22 / 32 ... leading bits (heavily tied keys) — the composite carries the prior index below the
//      key, so "more than 256 EQUAL keys" is just two more passes, with no special case
//   F  wait for rendezvous 2: smooth-L1 + CE over pos | neg, all gradients written (zeros for everybody else)
//   G  per-CTA partial sums in double; the last CTA to finish reduces them in a fixed order, divides by N, and resets the
//      rendezvous counters for the next launch
//
// Data-parallel jobs (gssd_xchg): the CTA that completes a rendezvous stores this rank's value, tagged with the step's
// epoch in the same 64-bit word, into every peer's exchange buffer over NVLink; CTAs then wait on their LOCAL buffer for all
// ranks' words instead of on the local counter.  The max of conf is published microseconds after the kernel starts and is
// needed only after the IoU sweep; N is published before the select and needed after it: the exchange latency hides.
#include <mutex>
#include <vector>

#include "common.cuh"

GSSD_PHASE_DECL(fused)

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

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// wait for a rendezvous counter of this GPU.  Every CTA is resident (cooperative launch), so this cannot deadlock; the bound
// (4 s) only turns a broken invariant — e.g. a state buffer shared by launches on two streams — into a trap instead of a hang
__device__ __forceinline__ void bar_wait(const uint32_t *counter, uint32_t target) {
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_gpu(counter) < target) {
        if ((++spins & 0x3ff) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}

// wait until every rank's word of this epoch is in the local exchange buffer; lanes < world of one warp; returns the word
__device__ __forceinline__ unsigned long long xchg_wait(const unsigned long long *slot, uint32_t epoch, unsigned long long timeout_ns) {
    unsigned long long v, t0 = 0;
    unsigned spins = 0;
    while (true) {
        v = ld_acquire_sys64(slot);
        if ((uint32_t)(v >> 32) == epoch) break;
        if ((++spins & 0xff) == 0 && timeout_ns) {                 // a rank that never arrives must not wedge the GPU
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeout_ns) __trap();
        }
    }
    return v;
}

template <int NT, bool C2, bool GRADS>
__global__ void __launch_bounds__(NT, 1) fused_kernel(FusedArgs a) {
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
    unsigned long long *cand = reinterpret_cast<unsigned long long *>(total + 2 * FBINS);
    uint32_t *keys = reinterpret_cast<uint32_t *>(cand + FCAND);
    float *conf_s = reinterpret_cast<float *>(keys + a.items);
    uint16_t *stag = reinterpret_cast<uint16_t *>(conf_s + (size_t)a.items * C);

    __shared__ int s_wi[NW];
    __shared__ float s_wf[NW];
    __shared__ float s_bbox[4][NW];
    __shared__ double s_red[2][NW];
    __shared__ int s_nlist;
    __shared__ uint32_t s_xmax_ord, s_digit, s_krem, s_eq, s_cand_count, s_wtot[NW];
    __shared__ int s_ntotal, s_npos_img, s_npos_cta;
    __shared__ unsigned long long s_cut;
    __shared__ bool s_last;

    const int n_chunks = (a.P + FCHUNK - 1) / FCHUNK;
    const int my_chunks = (int)rank < n_chunks ? (n_chunks - (int)rank + (int)S - 1) / (int)S : 0;
    const int trips = (my_chunks + SLOTS - 1) / SLOTS;
    const int p_last = a.P - 1;
    const bool dbg = blockIdx.x == 0 && blockIdx.y == 0;
    GSSD_PHASE(fused, 0, dbg);

    // epoch of this step in the peer exchange (read before anybody can have advanced it: it moves when the LAST CTA exits)
    uint32_t epoch = 0;
    if (a.x.world > 0) epoch = *reinterpret_cast<const volatile uint32_t *>(&a.x.peers[a.x.rank]->epoch) + 1;

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
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (S > 1) cluster_arrive();         // my receiving buffers are initialised; waited for before the first remote store

    // local max of conf -> batch max (box_utils.py:167), then arrive at rendezvous 1
    {
        float cmax = -INFINITY;
        for (int cj = 0; cj < my_chunks; ++cj) {
            const int ck = (int)rank + cj * (int)S;
            const int nf = min(FCHUNK, a.P - ck * FCHUNK) * C;
            const float *src = conf_s + (size_t)cj * FCHUNK * C;
            for (int i = tid; i < nf; i += NT) cmax = fmaxf(cmax, src[i]);
        }
        cmax = warp_max(cmax);
        if (lane == 0) s_wf[warp] = cmax;
        __syncthreads();
        if (tid == 0) {
            float mx = -INFINITY;
            for (int w = 0; w < NW; ++w) mx = fmaxf(mx, s_wf[w]);
            if (my_chunks > 0) atomicMax(&a.state->xmax_ord, f2ord(mx));
            __threadfence();
            const unsigned old = atomicAdd(&a.state->bar1, 1u);
            if (old == n_ctas - 1 && a.x.world > 0) {            // this rank's max is final: publish it to every rank
                const unsigned long long word = ((unsigned long long)epoch << 32) | atomicMax(&a.state->xmax_ord, 0u);
                for (int r = 0; r < a.x.world; ++r)
                    *reinterpret_cast<volatile unsigned long long *>(&a.x.peers[r]->xmax[epoch & 1][a.x.rank]) = word;
                __threadfence_system();
            }
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
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            bx1 = fminf(bx1, __shfl_xor_sync(FULL, bx1, o)); by1 = fminf(by1, __shfl_xor_sync(FULL, by1, o));
            bx2 = fmaxf(bx2, __shfl_xor_sync(FULL, bx2, o)); by2 = fmaxf(by2, __shfl_xor_sync(FULL, by2, o));
        }
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
    for (int t = 0; t < trips; ++t) {
        const int cj = t * SLOTS + (tid >> 8);                   // warp-uniform
        if (cj >= my_chunks) continue;
        const int p = ((int)rank + cj * (int)S) * FCHUNK + (tid & 255);
        const int pc = min(p, p_last);                           // a lane past the end repeats the last prior (its keys are ignored)
        const float4 pb = point_form(a.priors[pc]);
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
        if (G < 8) {
#pragma unroll 1
            for (int g = 0; g < G; ++g) sweep_one(g);
        } else if (!warp_cull) {
            for (int q = 0; q < n_list; ++q) sweep_one(glist[q]);
        } else {
            float bx1 = pb.x, by1 = pb.y, bx2 = pb.z, by2 = pb.w;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                bx1 = fminf(bx1, __shfl_xor_sync(FULL, bx1, o)); by1 = fminf(by1, __shfl_xor_sync(FULL, by1, o));
                bx2 = fmaxf(bx2, __shfl_xor_sync(FULL, bx2, o)); by2 = fmaxf(by2, __shfl_xor_sync(FULL, by2, o));
            }
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
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < NW; ++w) tot += s_wi[w];
            s_npos_cta = tot;
            if (S == 1) s_npos_img = tot;
            atomicAdd(&a.state->n_pos, tot);
            __threadfence();
            const unsigned old = atomicAdd(&a.state->bar2, 1u);
            if (old == n_ctas - 1 && a.x.world > 0) {
                const unsigned long long word = ((unsigned long long)epoch << 32) | (unsigned)atomicAdd(&a.state->n_pos, 0);
                for (int r = 0; r < a.x.world; ++r)
                    *reinterpret_cast<volatile unsigned long long *>(&a.x.peers[r]->npos[epoch & 1][a.x.rank]) = word;
                __threadfence_system();
            }
        }
    }
    GSSD_PHASE(fused, 3, dbg);

    // ---- D. the batch max of conf, then the mining keys (multibox_loss.py:91-99) ---------------------------------------------
    if (warp == 0) {
        uint32_t mo = 0;
        if (a.x.world > 0) {
            if (lane < a.x.world)
                mo = (uint32_t)xchg_wait(&a.x.peers[a.x.rank]->xmax[epoch & 1][lane], epoch, a.x.timeout_ns);
#pragma unroll
            for (int o = 16; o; o >>= 1) mo = max(mo, __shfl_xor_sync(FULL, mo, o));
        } else if (lane == 0) {
            bar_wait(&a.state->bar1, n_ctas);
            mo = ld_acquire_gpu(&a.state->xmax_ord);
        }
        if (lane == 0) s_xmax_ord = mo;
    }
    __syncthreads();
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
            const uint32_t *tot = S > 1 ? total + (pass & 1) * FBINS : hist;
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
            uint32_t above = 0;
            for (int w = warp + 1; w < NW; ++w) above += s_wtot[w];
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
        // the members of the cut bin go to every CTA of the image and are ranked there by counting
        const int sh = fpass_shift(pass);
        for (int li = tid; li < my_chunks * FCHUNK; li += NT) {
            const int p = ((int)rank + (li >> 8) * (int)S) * FCHUNK + (li & 255);
            if (p >= a.P) continue;
            const unsigned long long c = fcomp(keys[li], p);
            if ((c >> sh) == prefix) {
                if (S > 1) {
                    for (unsigned r = 0; r < S; ++r) {
                        const uint32_t slot = atomicAdd(cluster.map_shared_rank(&s_cand_count, r), 1u);
                        cluster.map_shared_rank(cand, r)[slot] = c;
                    }
                } else {
                    cand[atomicAdd(&s_cand_count, 1u)] = c;
                }
            }
        }
        if (S > 1) cluster.sync(); else __syncthreads();
        {
            constexpr int R = NT / FCAND;                        // threads per candidate
            const int ci = tid / R, part = tid % R;
            const unsigned long long c = ci < (int)eq ? cand[ci] : 0ull;
            uint32_t cnt = 0;
            for (int j = part; j < (int)eq; j += R) cnt += cand[j] > c;
#pragma unroll
            for (int o = R >> 1; o; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
            if (ci < (int)eq && part == 0 && cnt == k_rem - 1) s_cut = c;
        }
        __syncthreads();
        cut = s_cut;
    }
    GSSD_PHASE(fused, 5, dbg);

    // ---- F. N, then smooth-L1 (80-88), cross-entropy over pos | neg (108-113) and every gradient ----------------------------------------
    if (warp == 0) {
        int nt = 0;
        if (a.x.world > 0) {
            if (lane < a.x.world)
                nt = (int)(uint32_t)xchg_wait(&a.x.peers[a.x.rank]->npos[epoch & 1][lane], epoch, a.x.timeout_ns);
#pragma unroll
            for (int o = 16; o; o >>= 1) nt += __shfl_xor_sync(FULL, nt, o);
        } else if (lane == 0) {
            bar_wait(&a.state->bar2, n_ctas);
            nt = (int)ld_acquire_gpu(reinterpret_cast<const uint32_t *>(&a.state->n_pos));
        }
        if (lane == 0) s_ntotal = nt;
    }
    __syncthreads();
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
    acc_l = warp_sum(acc_l); acc_c = warp_sum(acc_c);
    if (lane == 0) { s_red[0][warp] = acc_l; s_red[1][warp] = acc_c; }
    __syncthreads();
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double l = 0.0, c = 0.0;
        for (int w = 0; w < NW; ++w) { l += s_red[0][w]; c += s_red[1][w]; }
        a.partials[2 * cta] = l; a.partials[2 * cta + 1] = c;
        __threadfence();
        s_last = atomicAdd(&a.state->done, 1u) == n_ctas - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double l = 0.0, c = 0.0;
        for (unsigned i = tid; i < n_ctas; i += NT) {            // fixed order -> deterministic
            l += __ldcg(&a.partials[2 * i]); c += __ldcg(&a.partials[2 * i + 1]);
        }
        l = warp_sum(l); c = warp_sum(c);
        if (lane == 0) { s_red[0][warp] = l; s_red[1][warp] = c; }
        __syncthreads();
        if (tid == 0) {
            l = 0.0; c = 0.0;
            for (int w = 0; w < NW; ++w) { l += s_red[0][w]; c += s_red[1][w]; }
            a.losses[0] = __fdiv_rn((float)l, n_f);              // multibox_loss.py:117-119
            a.losses[1] = __fdiv_rn((float)c, n_f);
            // everybody has left: the rendezvous counters are ready for the next launch on this state
            a.state->bar1 = 0; a.state->bar2 = 0; a.state->xmax_ord = 0; a.state->n_pos = 0; a.state->done = 0;
            if (a.x.world > 0) *reinterpret_cast<volatile uint32_t *>(&a.x.peers[a.x.rank]->epoch) = epoch;
        }
    }
    GSSD_PHASE(fused, 7, dbg);
}

static size_t fused_smem_bytes(int g_max, int S, int items, int C) {
    const size_t Gp = (size_t)((g_max + 1) & ~1);
    size_t b = Gp * 8 + (S > 1 ? (size_t)S * Gp * 8 : 0) + (size_t)g_max * (16 + 4 + 4 + 4 + 4) + 16;
    b += (size_t)3 * FBINS * 4 + (size_t)FCAND * 8;
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
    for (int S = 8; S >= 1; S >>= 1) {
        if (forced_s && S != forced_s) continue;
        if (S > 1 && S > n_chunks) continue;
        const int items = ceil_div(n_chunks, S) * FCHUNK;
        const size_t smem = fused_smem_bytes(g_max, S, items, C);
        if (smem > (size_t)optin - 2048) continue;
        // one CTA per SM keeps every phase at full speed; beyond that only what the occupancy calculator admits
        if (!forced_s && (long)B * S > sms) {
            if (S > 1) continue;
        }
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
        if (clusters < B) continue;
        pl.S = S; pl.NT = NT; pl.items = items; pl.smem = smem;
        *kern_out = kern;
        return pl;
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
    if (B <= 0 || P <= 0 || sum_G <= 0 || g_max <= 0 || negpos_ratio < 0) return sum_G <= 0 && B > 0 ? GSSD_ERR_EMPTY : GSSD_ERR_ARG;
    if (C < 2 || C > GSSD_MAX_CLASSES) return GSSD_ERR_ARG;
    if ((grad_loc == nullptr) != (grad_conf == nullptr)) return GSSD_ERR_ARG;
    if (g_max > GSSD_MAX_GT_PER_IMAGE || P > GSSD_MAX_PRIORS) return GSSD_ERR_LIMIT;
    if (ws_bytes < gssd_workspace_bytes(GSSD_WS_LOSS, B, P, C, sum_G, 0)) return GSSD_ERR_WS;
    const bool gr = grad_loc != nullptr, c2 = C == 2;
    const void *kern = nullptr;
    const FusedPlan pl = fused_plan(B, P, C, g_max, &kern, c2, gr);
    if (pl.S == 0) return GSSD_ERR_UNSUPPORTED;
    if (pl.smem < fused_smem_bytes(g_max, pl.S, pl.items, C)) return GSSD_ERR_UNSUPPORTED;
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
    static int coop_ok = 1;                                      // cluster + cooperative in one launch: dropped if the runtime refuses
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
