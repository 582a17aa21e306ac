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
#include "select.cuh"

namespace gssd {

constexpr int DET_NT = 512;

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
};

struct DetShared {
    SelectShared sel;
    int n_cand;
    int n_sel;
    int n_keep;
    uint32_t keep_bits[GSSD_MAX_TOP_K / 32];
};

// dynamic smem:  [ u32 keys[n] | (aliased later) u32 mask[top_k][words] ]  u64 ckey[k2]  float4 box[top_k]  float area[top_k]
template <bool NMS_MODE>
__global__ void __launch_bounds__(DET_NT) detect_kernel(DetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DetShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cl = NMS_MODE ? 0 : blockIdx.x;
    const int b = NMS_MODE ? 0 : blockIdx.y;
    const int n = a.P;                           // number of scores scanned
    const int top_k = a.top_k, k2 = a.k2;

    const size_t region_a = max((size_t)n * 4, (size_t)top_k * ((top_k + 31) / 32) * 4);
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned long long *ckey = reinterpret_cast<unsigned long long *>(smem_raw + ((region_a + 15) & ~(size_t)15));
    float4 *sbox = reinterpret_cast<float4 *>(ckey + k2);
    float *sarea = reinterpret_cast<float *>(sbox + top_k);

    float *out_slab = NMS_MODE ? nullptr : a.out + ((size_t)b * a.C + cl) * top_k * 5;
    int32_t *idx_slab = (NMS_MODE || !a.keep_idx) ? nullptr : a.keep_idx + ((size_t)b * a.C + cl) * top_k;

    if (!NMS_MODE && cl == 0) {                  // background slab stays zero
        for (int i = tid; i < top_k * 5; i += DET_NT) out_slab[i] = 0.f;
        if (idx_slab) for (int i = tid; i < top_k; i += DET_NT) idx_slab[i] = -1;
        if (a.count && tid == 0) a.count[b * a.C] = 0;
        return;
    }

    // ---- 1. threshold -> keys ------------------------------------------------------------------------
    if (tid == 0) { sh.n_cand = 0; sh.n_sel = 0; }
    __syncthreads();
    int mine = 0;
    for (int p = tid; p < n; p += DET_NT) {
        float s;
        bool cand;
        if (NMS_MODE) { s = a.scores[p]; cand = true; }
        else { s = a.conf[((size_t)b * a.P + p) * a.C + cl]; cand = s > a.conf_thresh; }   // strict >, line 69
        keys[p] = cand ? f2ord(s) : 0u;
        mine += cand;
    }
    mine = warp_sum(mine);
    if (lane == 0 && mine) atomicAdd(&sh.n_cand, mine);
    __syncthreads();
    const int n_cand = sh.n_cand;
    const int k = min(n_cand, top_k);
    const int words = (k + 31) / 32;             // 32-candidate blocks actually in use

    if (k > 0) {
        // ---- 2. top-k candidates (box_utils.py:194-196) ---------------------------------------------
        SelectResult sel;
        sel.v = 1u; sel.need = 0; sel.eq = 0; sel.tie_cut = 0; sel.low_first = false;   // key >= 1: every candidate
        if (n_cand > top_k) sel = radix_select<DET_NT, false>(keys, n, (uint32_t)top_k, false, &sh.sel);
        for (int i = tid; i < k2; i += DET_NT) ckey[i] = 0ull;
        __syncthreads();
        for (int p = tid; p < n; p += DET_NT) {
            uint32_t key = keys[p];
            if (key != 0u && sel.selected(key, p)) {
                int slot = atomicAdd(&sh.n_sel, 1);
                ckey[slot] = ((unsigned long long)key << 32) | (unsigned)p;
            }
        }
        __syncthreads();
        // ---- 3. bitonic sort, descending (score, index) -------------------------------------------------
        for (int size = 2; size <= k2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = tid; i < k2 / 2; i += DET_NT) {
                    int lo = 2 * i - (i & (stride - 1));
                    int hi = lo + stride;
                    bool desc = (lo & size) == 0;
                    unsigned long long x = ckey[lo], y = ckey[hi];
                    if ((x < y) == desc) { ckey[lo] = y; ckey[hi] = x; }
                }
                __syncthreads();
            }
        }
        // ---- 4. boxes of the candidates ------------------------------------------------------------------
        for (int i = tid; i < k; i += DET_NT) {
            unsigned p = (unsigned)(ckey[i] & 0xffffffffu);
            float4 bx = NMS_MODE ? a.boxes[p]
                                 : decode_box(a.loc[(size_t)b * a.P + p], a.priors[p], a.var0, a.var1);
            sbox[i] = bx;
            sarea[i] = box_area(bx);                               // box_utils.py:193
        }
        __syncthreads();                                            // keys[] is dead from here: mask aliases it
        // ---- 5. suppression bitmask: bit j of row i (j > i) = box j is removed when i is kept ------------
        for (int item = tid; item < k * words; item += DET_NT) {
            const int i = item / words, w = item - i * words;
            uint32_t bits = 0;
            if (32 * w + 31 > i) {
                const float4 bi = sbox[i];
                const float ai = sarea[i];
                const int j0 = max(32 * w, i + 1), j1 = min(32 * w + 32, k);
                for (int j = j0; j < j1; ++j) {
                    const float4 bj = sbox[j];
                    float xx1 = fmaxf(bj.x, bi.x), yy1 = fmaxf(bj.y, bi.y);          // box_utils.py:220-223
                    float xx2 = fminf(bj.z, bi.z), yy2 = fminf(bj.w, bi.w);
                    float ww = __fsub_rn(xx2, xx1), hh = __fsub_rn(yy2, yy1);
                    ww = ww < 0.f ? 0.f : ww; hh = hh < 0.f ? 0.f : hh;              // 229-230
                    float inter = __fmul_rn(ww, hh);
                    float uni = __fadd_rn(__fsub_rn(sarea[j], inter), ai);           // 233-234
                    float iou = __fdiv_rn(inter, uni);
                    if (!(iou <= a.nms_thresh)) bits |= 1u << (j - 32 * w);          // 237 (NaN -> removed)
                }
            }
            mask[i * words + w] = bits;
        }
        __syncthreads();
        // ---- 6. greedy resolve by one warp: lane w owns removed-word w --------------------------------------
        if (warp == 0) {
            uint32_t removed = 0;                                   // word `lane` of the removed set
            for (int c = 0; c < words; ++c) {
                const int row = 32 * c + lane;
                uint32_t diag = row < k ? mask[row * words + c] : 0u;
                uint32_t cur = __shfl_sync(FULL, removed, c);
                if (32 * c + 32 > k) cur |= ~0u << (k - 32 * c);    // rows beyond k do not exist
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    uint32_t d = __shfl_sync(FULL, diag, t);
                    if (!((cur >> t) & 1u)) cur |= d;
                }
                const uint32_t kept = ~cur;
                if (lane == 0) sh.keep_bits[c] = kept;
                if (lane > c && lane < words) {
                    uint32_t acc = 0;
                    for (uint32_t m = kept; m; m &= m - 1) acc |= mask[(32 * c + __ffs(m) - 1) * words + lane];
                    removed |= acc;
                }
            }
        }
        __syncthreads();
    }

    // ---- 7. emit ---------------------------------------------------------------------------------------------
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
}

static size_t det_smem_bytes(int n, int top_k, int k2) {
    const int words = (top_k + 31) / 32;
    size_t region_a = (size_t)n * 4;
    size_t m = (size_t)top_k * words * 4;
    if (m > region_a) region_a = m;
    region_a = (region_a + 15) & ~(size_t)15;
    return region_a + (size_t)k2 * 8 + (size_t)top_k * 16 + (size_t)top_k * 4 + 16;
}

static int next_pow2(int v) { int p = 2; while (p < v) p <<= 1; return p; }

template <bool NMS_MODE>
static int launch_detect(DetArgs &a, dim3 grid, cudaStream_t st) {
    a.k2 = next_pow2(a.top_k);
    size_t smem = det_smem_bytes(a.P, a.top_k, a.k2);
    if (smem > 227 * 1024) return GSSD_ERR_LIMIT;
    auto kern = detect_kernel<NMS_MODE>;
    if (smem > 48 * 1024)
        GSSD_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, DET_NT, smem, st>>>(a);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

}  // namespace gssd

using namespace gssd;

extern "C" int gssd_detect(const float *loc, const float *conf, const float *priors, int B, int P, int C,
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
    return launch_detect<false>(a, dim3(C, B, 1), (cudaStream_t)stream);
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
