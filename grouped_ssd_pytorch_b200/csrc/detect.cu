// detect.cu — Detect.forward (detection_pytorch_ver_1point5.py:33-89) and box_utils.nms
// (box_utils.py:174-238).
//
// One CTA per (image, class).  The scores of the class are thresholded into order-preserving uint32
// keys in shared memory (0 = not a candidate); a radix select (select.cuh) finds the top_k candidates
// — the reference sorts all candidates and keeps the last top_k BEFORE suppression (box_utils.py:194-196)
// — only those <= top_k boxes are decoded (box_utils.py:139-157), sorted (bitonic, (score, index)
// descending = the reference's stable ascending sort read from its end), an upper-triangular IoU
// suppression bitmask is built by all threads and resolved by one warp, 32 candidates per step with
// shuffles only; surviving rows are written in descending score, the rest of the slab is zeroed
// (detection_pytorch_ver_1point5.py:56, 82-84).
#include <stdlib.h>

#include <math.h>

#include "select.cuh"

GSSD_PHASE_DECL(detect)

namespace gssd {

constexpr int DET_NT = 1024;

struct DetArgs {
    // Detect mode
    const float4 *loc; const float *conf; const float4 *priors;
    int B, P, C;
    float conf_thresh, var0, var1;
    float *out; int32_t *count; int32_t *keep_idx;
    // NMS mode
    const float4 *boxes; const float *scores; int64_t *keep;
    // common
    int top_k, k2;          // k2 = next power of two >= top_k
    float nms_thresh;
    // logits mode (gssd_detect_logits): conf holds raw class logits, the class score is softmax(conf + bias)[cl]
    int logits;
    float bias[GSSD_MAX_CLASSES];
    // two classes: a prior whose logit difference (x1 + b1) - (x0 + b0) is below `cull` cannot reach conf_thresh — its score is
    // sigmoid(difference) up to a few ulp — and skips the exact softmax (two expf and an IEEE divide); -inf switches the test off
    float cull;
};

struct DetShared {
    SelectShared sel;
    int n_cand;
    int n_sel;
    int n_keep;
    int k_final;                       // candidates that enter NMS (read by the helper CTAs of the cluster)
    uint32_t keep_bits[GSSD_MAX_TOP_K / 32];
};

constexpr int DET_CAND_CAP = 1024;      // candidates that are sorted directly, without a select pass

// ---- block-wide bitonic sort, descending, n2 = power of two >= 64 ------------------------------------------
// Strides >= 64 go through shared memory with a block barrier; everything below runs inside one warp on a
// 64-element block held in registers (2 per lane) with shuffles, so a 256-sort needs 4 block barriers
// instead of 36.
// sub-stages of phase `sz` that stay inside a 64-element block: stride 32 (the lane's own pair, only when
// sz >= 64) and strides 16..1 (partner lane = lane ^ stride).  e0 / e1 = global indices of v0 / v1.
__device__ __forceinline__ void warp_substages(unsigned long long &v0, unsigned long long &v1, int e0, int e1,
                                               int sz, int lane) {
    const bool d0 = (e0 & sz) == 0, d1 = (e1 & sz) == 0;             // descending sub-sequence?
    if (sz >= 64 && (v0 < v1) == d0) { const unsigned long long t = v0; v0 = v1; v1 = t; }
#pragma unroll
    for (int st = 16; st >= 1; st >>= 1) {
        if (st < sz) {
            const bool lower = (lane & st) == 0;
            cmpx(v0, shfl_xor_u64(v0, st), d0 == lower);
            cmpx(v1, shfl_xor_u64(v1, st), d1 == lower);
        }
    }
}

template <int NT>
__device__ void bitonic_sort_desc(unsigned long long *a, int n2) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_blocks = n2 >> 6;
    for (int size = 64; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride >= 64; stride >>= 1) {      // only for size >= 128
            for (int i = tid; i < (n2 >> 1); i += NT) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long x = a[lo], y = a[hi];
                if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
            }
            __syncthreads();
        }
        for (int blk = warp; blk < n_blocks; blk += NT / 32) {
            const int e0 = (blk << 6) + lane, e1 = e0 + 32;
            unsigned long long v0 = a[e0], v1 = a[e1];
            if (size == 64) {                                            // phases 2..32 live in the block too
                for (int sz = 2; sz < 64; sz <<= 1) warp_substages(v0, v1, e0, e1, sz, lane);
            }
            warp_substages(v0, v1, e0, e1, size, lane);
            a[e0] = v0; a[e1] = v1;
        }
        __syncthreads();
    }
}

// softmax over a row of class logits, the way torch's softmax kernel evaluates it (ssd_multiphase_custom_group.py:388):
// subtract the row max, exp, sum in class order, IEEE divide
__device__ __forceinline__ float softmax2_class1(float x0, float x1) {
    const float m = fmaxf(x0, x1);
    const float e0 = expf(__fsub_rn(x0, m)), e1 = expf(__fsub_rn(x1, m));
    return __fdiv_rn(e1, __fadd_rn(e0, e1));
}
__device__ __forceinline__ float softmax_class(const float *row, const float *bias, int C, int cl) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, __fadd_rn(__ldg(row + c), bias[c]));
    float sum = 0.f, mine = 0.f;
    for (int c = 0; c < C; ++c) {
        const float e = expf(__fsub_rn(__fadd_rn(__ldg(row + c), bias[c]), m));
        sum = __fadd_rn(sum, e);
        if (c == cl) mine = e;
    }
    return __fdiv_rn(mine, sum);
}

// dynamic smem:  [ u32 keys[n] | (aliased later) u32 mask[top_k][words] ]  u64 ckey[max(k2, CAP)]  float4 box[top_k]  float area[top_k]
// CL: launched as a cluster (small batches: helper CTAs for the bitmask, one round of loads in the threshold pass, registers
// unconstrained); !CL: two CTAs per SM for machine-filling batches.
// Both variants are held to 32 registers per thread: a 1024-thread CTA at 64 registers fills an SM's register file, and the
// loss kernel that runs beside Detect (bench step, training step) then has to wait for those SMs — measured on the batch-32
// step: 43.0 -> 41.6 us with Detect itself 22.9 -> 23.5 us (profiles/r2_detect_regs_ab.txt).
#ifndef GSSD_DET_CL_MINB
#define GSSD_DET_CL_MINB 2
#endif
template <bool NMS_MODE, bool CL>
__global__ void __launch_bounds__(DET_NT, CL ? GSSD_DET_CL_MINB : 2) detect_kernel(DetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DetShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Detect mode runs a cluster of S CTAs per (image, class): CTA 0 does the whole job, the others only help with the
    // suppression bitmask (the one phase that is bound by the issue rate of a single SM) through distributed shared memory
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned S = CL ? cluster.num_blocks() : 1u;
    const unsigned crank = CL ? cluster.block_rank() : 0u;
    const int cl = NMS_MODE ? 0 : blockIdx.x / S;
    const int b = NMS_MODE ? 0 : blockIdx.y;
    const int n = a.P;                           // number of scores scanned
    const int top_k = a.top_k;
    const int cap = max(a.k2, DET_CAND_CAP);

    const size_t region_a = max((size_t)n * 4, (size_t)top_k * ((top_k + 31) / 32) * 4);
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned long long *ckey = reinterpret_cast<unsigned long long *>(smem_raw + ((region_a + 15) & ~(size_t)15));
    float4 *sbox = reinterpret_cast<float4 *>(ckey + cap);
    float *sarea = reinterpret_cast<float *>(sbox + top_k);

    float *out_slab = NMS_MODE ? nullptr : a.out + ((size_t)b * a.C + cl) * top_k * 5;
    int32_t *idx_slab = (NMS_MODE || !a.keep_idx) ? nullptr : a.keep_idx + ((size_t)b * a.C + cl) * top_k;

    if (!NMS_MODE && cl == 0) {                  // background slab stays zero (the whole cluster leaves: no cluster barrier)
        if (crank != 0) return;
        for (int i = tid; i < top_k * 5; i += DET_NT) out_slab[i] = 0.f;
        if (idx_slab) for (int i = tid; i < top_k; i += DET_NT) idx_slab[i] = -1;
        if (a.count && tid == 0) a.count[b * a.C] = 0;
        return;
    }

    const bool dbg = cl == 1 && crank == 0 && blockIdx.y == 0;
    GSSD_PHASE(detect, 0, dbg);
    // ---- 1. threshold -> keys, and an optimistic compaction of the candidates -----------------------------
    if (tid == 0) { sh.n_cand = 0; sh.n_sel = 0; sh.k_final = 0; }
    __syncthreads();
    // C == 2 in a cluster: every CTA scans its own slice of the priors into its own (keys, candidate list); CTA 0 then pulls
    // the helpers' candidates (and, if there are too many for the direct sort, their keys) through distributed shared memory
    const bool split_scan = CL && S > 1 && !NMS_MODE && a.C == 2 && (n & 1) == 0;
    if (crank == 0 || split_scan) {
        // append one candidate per lane: warp-aggregated slot allocation
        auto push = [&](bool cand, uint32_t key, int p) {
            const unsigned m = __ballot_sync(FULL, cand);
            if (m) {
                int slot = 0;
                if (lane == 0) slot = atomicAdd(&sh.n_cand, __popc(m));
                slot = __shfl_sync(FULL, slot, 0) + __popc(m & ((1u << lane) - 1));
                if (cand && slot < DET_CAND_CAP) ckey[slot] = ((unsigned long long)key << 32) | (unsigned)p;
            }
        };
        constexpr int U = CL ? 3 : 4;            // loads in flight per thread: with 2 CTAs per image a CTA owns 2.1 float4 per thread at
                                                 // P = 8732 -> one round of loads
        if (!NMS_MODE && a.C == 2 && (n & 1) == 0) {
            // rows are (background, class 1) pairs: one float4 = two priors, the .y / .w lanes are ours
            const float4 *src = reinterpret_cast<const float4 *>(a.conf + (size_t)b * a.P * 2);
            const int n4 = n >> 1;
            const int per = split_scan ? (n4 + (int)S - 1) / (int)S : n4;
            const int q_lo = split_scan ? (int)crank * per : 0, q_hi = min(n4, q_lo + per);
            for (int base = q_lo; base < q_hi; base += DET_NT * U) {
                float4 sv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int q = base + u * DET_NT + tid;
                    sv[u] = q < q_hi ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (a.logits) {                                         // fused softmax: (.y, .w) become the class-1 scores
                        const float y0 = __fadd_rn(sv[u].x, a.bias[0]), y1 = __fadd_rn(sv[u].y, a.bias[1]);
                        const float z0 = __fadd_rn(sv[u].z, a.bias[0]), z1 = __fadd_rn(sv[u].w, a.bias[1]);
                        // most priors are far below the threshold: 0 stands for "not a candidate" (conf_thresh >= 0 whenever
                        // cull is finite); the exact torch-formula score is computed for the others only
                        sv[u].y = __fsub_rn(y1, y0) < a.cull ? 0.f : softmax2_class1(y0, y1);
                        sv[u].w = __fsub_rn(z1, z0) < a.cull ? 0.f : softmax2_class1(z0, z1);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int q = base + u * DET_NT + tid;
                    const bool ok = q < q_hi;
                    const bool c0 = ok && sv[u].y > a.conf_thresh, c1 = ok && sv[u].w > a.conf_thresh;   // strict >, line 69
                    const uint32_t k0 = c0 ? f2ord(sv[u].y) : 0u, k1 = c1 ? f2ord(sv[u].w) : 0u;
                    if (ok) *reinterpret_cast<uint2 *>(keys + 2 * q) = make_uint2(k0, k1);
                    push(c0, k0, 2 * q);
                    push(c1, k1, 2 * q + 1);
                }
            }
        } else {
            const float *src = NMS_MODE ? a.scores : a.conf + (size_t)b * a.P * a.C + cl;
            const int stride = NMS_MODE ? 1 : a.C;
            for (int base = 0; base < n; base += DET_NT * U) {
                float sv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int p = base + u * DET_NT + tid;
                    sv[u] = p < n ? __ldg(src + (size_t)p * stride) : 0.f;
                    if (!NMS_MODE && a.logits && p < n) sv[u] = softmax_class(a.conf + ((size_t)b * a.P + p) * a.C, a.bias, a.C, cl);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int p = base + u * DET_NT + tid;
                    const bool cand = p < n && (NMS_MODE || sv[u] > a.conf_thresh);  // strict >, line 69
                    const uint32_t key = cand ? f2ord(sv[u]) : 0u;
                    if (p < n) keys[p] = key;
                    push(cand, key, p);
                }
            }
        }
    }
    if (split_scan) {
        cluster.sync();                                              // every CTA's slice is scanned
        if (crank == 0) {
            int cnt[8];
            int total = 0;
            bool overflow = false;
            for (unsigned r = 0; r < S; ++r) {
                cnt[r] = r == 0 ? sh.n_cand : cluster.map_shared_rank(&sh, r)->n_cand;
                overflow |= cnt[r] > DET_CAND_CAP;
                total += cnt[r];
            }
            __syncthreads();                                         // everyone has read sh.n_cand before it is replaced
            if (total <= DET_CAND_CAP && !overflow) {                // the usual case: append the helpers' candidates
                int off = cnt[0];
                for (unsigned r = 1; r < S; ++r) {
                    const unsigned long long *src = cluster.map_shared_rank(ckey, r);
                    for (int i = tid; i < cnt[r]; i += DET_NT) ckey[off + i] = src[i];
                    off += cnt[r];
                }
            } else {                                                 // too many for the direct sort: the select needs every key
                const int n4 = n >> 1, per = (n4 + (int)S - 1) / (int)S;
                for (unsigned r = 1; r < S; ++r) {
                    const uint32_t *src = cluster.map_shared_rank(keys, r);
                    const int lo = 2 * min(n4, (int)r * per), hi = 2 * min(n4, ((int)r + 1) * per);
                    for (int i = lo + tid; i < hi; i += DET_NT) keys[i] = src[i];
                }
            }
            if (tid == 0) sh.n_cand = total;
        }
    }
    __syncthreads();
    const int n_cand = sh.n_cand;
    int k = min(n_cand, top_k);
    int words = (k + 31) / 32;                   // 32-candidate blocks actually in use

    GSSD_PHASE(detect, 1, dbg);
    if (crank == 0 && k > 0) {
        // ---- 2. the top_k candidates in descending (score, index) order (box_utils.py:194-196) ---------
        int n2;
        if (n_cand <= DET_CAND_CAP) {            // few candidates: sort them all
            n2 = 64; while (n2 < n_cand) n2 <<= 1;
            for (int i = n_cand + tid; i < n2; i += DET_NT) ckey[i] = 0ull;
        } else {                                 // many: radix select of the top_k, then sort those
            const unsigned long long cut = radix_select<DET_NT, false>(keys, n, 0u, (uint32_t)top_k, false, &sh.sel);
            n2 = max(a.k2, 64);
            for (int i = tid; i < n2; i += DET_NT) ckey[i] = 0ull;
            __syncthreads();
            for (int p = tid; p < n; p += DET_NT) {
                const uint32_t key = keys[p];
                const unsigned long long ck = ((unsigned long long)key << 32) | (unsigned)p;
                if (key != 0u && ck >= cut) ckey[atomicAdd(&sh.n_sel, 1)] = ck;
            }
        }
        __syncthreads();
        GSSD_PHASE(detect, 2, dbg);
        GSSD_PHASE(detect, 3, dbg);
        bitonic_sort_desc<DET_NT>(ckey, n2);
        GSSD_PHASE(detect, 4, dbg);
        // ---- 3. boxes of the candidates ------------------------------------------------------------------
        for (int i = tid; i < k; i += DET_NT) {
            unsigned p = (unsigned)(ckey[i] & 0xffffffffu);
            float4 bx = NMS_MODE ? a.boxes[p]
                                 : decode_box(a.loc[(size_t)b * a.P + p], a.priors[p], a.var0, a.var1);
            sbox[i] = bx;
            sarea[i] = box_area(bx);                               // box_utils.py:193
        }
        if (tid == 0) sh.k_final = k;
        __syncthreads();                                            // keys[] is dead from here: mask aliases it
    }
    GSSD_PHASE(detect, 5, dbg);
    // ---- 4. suppression bitmask: bit j of row i (j > i) = box j is removed when i is kept ----------------
    // The rows are dealt to the warps of ALL CTAs of the cluster: the helpers copy the boxes out of CTA 0's shared memory
    // and store their mask words straight into it (distributed shared memory).
    uint32_t *mask_dst = mask;
    if (CL && S > 1) {
        cluster.sync();                                              // CTA 0's boxes are final
        if (crank != 0) {
            const DetShared *sh0 = cluster.map_shared_rank(&sh, 0);
            k = sh0->k_final;
            words = (k + 31) / 32;
            const float4 *sbox0 = cluster.map_shared_rank(sbox, 0);
            const float *sarea0 = cluster.map_shared_rank(sarea, 0);
            for (int i = tid; i < k; i += DET_NT) { sbox[i] = sbox0[i]; sarea[i] = sarea0[i]; }
            mask_dst = cluster.map_shared_rank(mask, 0);
            __syncthreads();
        }
    }
    if (k > 0) {
        // one warp per (row, 32-column word), one IoU per lane, the word is the ballot
        const float thr = a.nms_thresh;
        const float eps = thr * 9.5367431640625e-07f;               // 2^-20 relative: >> the 2-ulp error of the fast divide
        for (int i = crank * (DET_NT / 32) + warp; i < k; i += S * (DET_NT / 32)) {
            const float4 bi = sbox[i];
            const float ai = sarea[i];
            for (int w = i >> 5; w < words; ++w) {                  // strictly-lower words are never read
                const int j = 32 * w + lane;
                bool sup = false;
                if (j > i && j < k) {
                    const float4 bj = sbox[j];
                    const float xx1 = fmaxf(bj.x, bi.x), yy1 = fmaxf(bj.y, bi.y);    // box_utils.py:220-223
                    const float xx2 = fminf(bj.z, bi.z), yy2 = fminf(bj.w, bi.w);
                    float ww = __fsub_rn(xx2, xx1), hh = __fsub_rn(yy2, yy1);
                    ww = ww < 0.f ? 0.f : ww; hh = hh < 0.f ? 0.f : hh;              // 229-230
                    const float inter = __fmul_rn(ww, hh);
                    const float uni = __fadd_rn(__fsub_rn(sarea[j], inter), ai);     // 233-234
                    // IoU = inter/uni (IEEE) and "kept iff IoU <= thr" (235-237, NaN -> removed).  The fast
                    // divide decides every case that is not within 2^-20 of the threshold; the rest takes
                    // the exact one.
                    const float q = __fdividef(inter, uni);
                    const bool sane = uni > 0.f && uni < 1e30f;
                    if (sane && q > thr + eps) sup = true;
                    else if (sane && q < thr - eps) sup = false;
                    else sup = !(__fdiv_rn(inter, uni) <= thr);
                }
                const unsigned bits = __ballot_sync(FULL, sup);
                if (lane == 0) mask_dst[i * words + w] = bits;
            }
        }
    }
    if (CL && S > 1) {
        cluster.sync();                                              // every mask word has landed in CTA 0
        if (crank != 0) return;
    } else {
        __syncthreads();
    }
    GSSD_PHASE(detect, 6, dbg);
    if (k > 0) {
        // ---- 5. greedy resolve by one warp: lane w owns removed-word w --------------------------------------
        if (warp == 0) {
            uint32_t removed = 0;                                   // word `lane` of the removed set
            for (int c = 0; c < words; ++c) {
                const int row = 32 * c + lane;
                const uint32_t diag = row < k ? mask[row * words + c] : 0u;
                uint32_t cur = __shfl_sync(FULL, removed, c);
                if (32 * c + 32 > k) cur |= ~0u << (k - 32 * c);    // rows beyond k do not exist
                // only rows that suppress something inside this block can change `cur`
                const uint32_t nz = __ballot_sync(FULL, diag != 0u);
                uint32_t todo = nz & ~cur;
                while (todo) {                                      // warp-uniform
                    const int t = __ffs(todo) - 1;
                    cur |= __shfl_sync(FULL, diag, t);
                    todo = nz & ~cur & ~((2u << t) - 1u);
                }
                const uint32_t kept = ~cur;
                if (lane == 0) sh.keep_bits[c] = kept;
                // rows kept in this block remove boxes of the later blocks: OR-reduce their words
                const bool mine_kept = (kept >> lane) & 1u;
                for (int w = c + 1; w < words; ++w) {
                    const uint32_t r = __reduce_or_sync(FULL, mine_kept ? mask[row * words + w] : 0u);
                    if (lane == w) removed |= r;
                }
            }
        }
        __syncthreads();
    }

    GSSD_PHASE(detect, 7, dbg);
    // ---- 6. emit ---------------------------------------------------------------------------------------------
    int n_keep = 0;
    if (k > 0) for (int c = 0; c < words; ++c) n_keep += __popc(sh.keep_bits[c]);
    if (NMS_MODE) {
        for (int i = tid; i < n; i += DET_NT) a.keep[i] = 0;        // box_utils.py:186
        __syncthreads();
        if (tid == 0) *a.count = n_keep;
    } else {
        for (int i = n_keep * 5 + tid; i < top_k * 5; i += DET_NT) out_slab[i] = 0.f;
        if (idx_slab) for (int i = n_keep + tid; i < top_k; i += DET_NT) idx_slab[i] = -1;
        if (a.count && tid == 0) a.count[b * a.C + cl] = n_keep;
    }
    for (int i = tid; i < k; i += DET_NT) {
        const int c = i >> 5, bit = i & 31;
        const uint32_t kb = sh.keep_bits[c];
        if (!((kb >> bit) & 1u)) continue;
        int r = __popc(kb & ((1u << bit) - 1));
        for (int q = 0; q < c; ++q) r += __popc(sh.keep_bits[q]);
        const unsigned long long ck = ckey[i];
        const unsigned p = (unsigned)(ck & 0xffffffffu);
        if (NMS_MODE) {
            a.keep[r] = (int64_t)p;
        } else {
            const float4 bx = sbox[i];
            float *o = out_slab + 5 * r;                            // detection_pytorch_ver_1point5.py:82-84
            o[0] = ord2f((uint32_t)(ck >> 32)); o[1] = bx.x; o[2] = bx.y; o[3] = bx.z; o[4] = bx.w;
            if (idx_slab) idx_slab[r] = (int32_t)p;
        }
    }
    __syncthreads();
    GSSD_PHASE(detect, 8, dbg);
}

static size_t det_smem_bytes(int n, int top_k, int k2) {
    const int words = (top_k + 31) / 32;
    size_t region_a = (size_t)n * 4;
    size_t m = (size_t)top_k * words * 4;
    if (m > region_a) region_a = m;
    region_a = (region_a + 15) & ~(size_t)15;
    const size_t cap = k2 > DET_CAND_CAP ? k2 : DET_CAND_CAP;
    return region_a + cap * 8 + (size_t)top_k * 16 + (size_t)top_k * 4 + 16;
}

static int next_pow2(int v) { int p = 2; while (p < v) p <<= 1; return p; }

template <bool NMS_MODE, bool CL>
static int launch_detect_as(DetArgs &a, dim3 grid, int S, size_t smem, cudaStream_t st) {
    auto kern = detect_kernel<NMS_MODE, CL>;
    GSSD_RETURN_IF_CUDA(allow_max_smem(reinterpret_cast<const void *>(kern)));
    if (!CL) {
        kern<<<grid, DET_NT, smem, st>>>(a);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid.x * S, grid.y, 1);
        cfg.blockDim = dim3(DET_NT, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        GSSD_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    }
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

template <bool NMS_MODE>
static int launch_detect(DetArgs &a, dim3 grid, cudaStream_t st) {
    a.k2 = next_pow2(a.top_k);
    size_t smem = det_smem_bytes(a.P, a.top_k, a.k2);
    if (smem > 227 * 1024) return GSSD_ERR_LIMIT;
    // helper CTAs per (image, class) for the bitmask: only while they find idle SMs (small batches)
    int S = 1;
    if (!NMS_MODE) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long active = (long)grid.y * (grid.x > 1 ? grid.x - 1 : 1);    // class 0 leaves at once
        // measured inside the whole step at batch 32 (Detect beside match + loss): 2 helpers/image 50.0 us per step, none 53.7,
        // 4 helpers 57.0 (they take SMs from the loss kernels); 4 only when the batch leaves most of the GPU idle anyway
        if (active <= 8) S = 4; else if (active * 2 <= sms) S = 2;
        static const int forced = []{ const char *e = getenv("GSSD_DETECT_CLUSTER"); return e ? atoi(e) : 0; }();
        if (forced == 1 || forced == 2 || forced == 4) S = forced;
    }
    // cluster variant: at 32 registers two of these CTAs would fit one SM, and the scheduler then packs the CTAs of a cluster onto
    // it (measured: batch 64, 22.7 -> 30.6 us); asking for more than half an SM's shared memory keeps it at one per SM while
    // leaving room (registers, threads, ~110 KB) for a CTA of the loss kernel beside it
    if (!NMS_MODE && S > 1) return launch_detect_as<NMS_MODE, true>(a, grid, S, smem > 116 * 1024 ? smem : (size_t)116 * 1024, st);
    return launch_detect_as<NMS_MODE, false>(a, grid, 1, smem, st);
}

}  // namespace gssd

using namespace gssd;

static int detect_impl(const float *loc, const float *conf, const float *class_bias, bool logits, const float *priors, int B, int P, int C,
                       int top_k, float conf_thresh, float nms_thresh, float var0, float var1,
                       float *out, int32_t *count, int32_t *keep_idx, void *stream) {
    if (!loc || !conf || !priors || !out) return GSSD_ERR_ARG;
    if (B <= 0 || P <= 0 || C < 1 || top_k <= 0) return GSSD_ERR_ARG;
    if (nms_thresh <= 0) return GSSD_ERR_VALUE;                     // detection_pytorch_ver_1point5.py:39-40
    if (P > GSSD_MAX_PRIORS || top_k > GSSD_MAX_TOP_K || C > GSSD_MAX_CLASSES) return GSSD_ERR_LIMIT;
    DetArgs a = {};
    a.loc = reinterpret_cast<const float4 *>(loc); a.conf = conf; a.priors = reinterpret_cast<const float4 *>(priors);
    a.B = B; a.P = P; a.C = C; a.conf_thresh = conf_thresh; a.var0 = var0; a.var1 = var1;
    a.out = out; a.count = count; a.keep_idx = keep_idx;
    a.top_k = top_k; a.nms_thresh = nms_thresh;
    a.logits = logits ? 1 : 0;
    a.cull = -INFINITY;
    if (logits && C == 2 && conf_thresh > 0.f && conf_thresh < 1.f) {
        // score > thr  <=>  logit difference > log(thr / (1 - thr)) in exact arithmetic; the computed score is within a few ulp
        // (5e-7 relative) of sigmoid(difference), which a margin of 1e-2 in the difference covers with room for every thr <= 0.999 (checked on 4e6 pairs per threshold with torch.softmax)
        const double l = log((double)conf_thresh / (1.0 - (double)conf_thresh)) - 1e-2;
        if (1.0 - (double)conf_thresh >= 1e-3) a.cull = nextafterf((float)l, -INFINITY);
    }
    if (logits && class_bias) for (int c = 0; c < C; ++c) a.bias[c] = class_bias[c];
    return launch_detect<false>(a, dim3(C, B, 1), (cudaStream_t)stream);
}

extern "C" int gssd_detect(const float *loc, const float *conf, const float *priors, int B, int P, int C,
                           int top_k, float conf_thresh, float nms_thresh, float var0, float var1,
                           float *out, int32_t *count, int32_t *keep_idx, void *stream) {
    return detect_impl(loc, conf, nullptr, false, priors, B, P, C, top_k, conf_thresh, nms_thresh, var0, var1, out, count, keep_idx, stream);
}

extern "C" int gssd_detect_logits(const float *loc, const float *conf_logits, const float *class_bias_host, const float *priors,
                                  int B, int P, int C, int top_k, float conf_thresh, float nms_thresh, float var0, float var1,
                                  float *out, int32_t *count, int32_t *keep_idx, void *stream) {
    if (C < 2) return GSSD_ERR_ARG;
    return detect_impl(loc, conf_logits, class_bias_host, true, priors, B, P, C, top_k, conf_thresh, nms_thresh, var0, var1, out, count,
                       keep_idx, stream);
}

extern "C" int gssd_nms(const float *boxes, const float *scores, int n, float overlap, int top_k,
                        int64_t *keep, int32_t *count, void *ws, size_t ws_bytes, void *stream) {
    (void)ws; (void)ws_bytes;
    if (!count || top_k <= 0 || n < 0) return GSSD_ERR_ARG;
    if (n == 0) {                                                   // box_utils.py:187-188
        GSSD_RETURN_IF_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), (cudaStream_t)stream));
        return GSSD_OK;
    }
    if (!boxes || !scores || !keep) return GSSD_ERR_ARG;
    if (n > GSSD_MAX_PRIORS || top_k > GSSD_MAX_TOP_K) return GSSD_ERR_LIMIT;
    DetArgs a = {};
    a.boxes = reinterpret_cast<const float4 *>(boxes); a.scores = scores; a.keep = keep; a.count = count;
    a.P = n; a.C = 1; a.B = 1;
    a.top_k = top_k; a.nms_thresh = overlap;
    return launch_detect<true>(a, dim3(1, 1, 1), (cudaStream_t)stream);
}
