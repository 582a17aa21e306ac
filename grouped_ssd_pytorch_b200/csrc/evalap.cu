// evalap.cu — the consumer of Detect's output in the reference's evaluator, on the GPU:
//   gssd_collect_detections : test_ap_iobb.py:126-149 for a whole batch — class slab, `score > 0`, boxes scaled to the image,
//                             `score > thresh`, an image-id column in front — as a prefix count per image + one compaction
//   gssd_ap_match           : test_ap_iobb.py:251-297 — per detection, in descending score inside its image, the ground-truth
//                             box of largest IoU (and of largest IoBB = intersection over the DETECTION's area), greedy
//                             true/false-positive assignment with one "already detected" flag per box and threshold
//   gssd_ap_sort            : the global descending-score order of make_pred (test_ap_iobb.py:213-223), a stable LSD radix sort
//   gssd_ap_curve           : test_ap_iobb.py:299-326 + voc_ap (10-41): cumulative TP / FP in that order, precision / recall in
//                             float64, then the 11-point VOC-07 metric or the area under the precision envelope
// Float64 throughout where the reference computes in float64 (`astype(float)`): the TP/FP decisions are bit-exact.
#include "common.cuh"

namespace gssd {

// ---- collect ------------------------------------------------------------------------------------------------------------------
// Detect writes the rows of an (image, class) slab in descending score and zero-pads it, so the rows that pass `score > 0`
// and `score > thresh` are a prefix of the slab.
__global__ void __launch_bounds__(256) collect_count_kernel(const float *__restrict__ out, int B, int C, int top_k, int cls, float thresh,
                                                            int32_t *__restrict__ counts) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float *slab = out + ((size_t)warp * C + cls) * top_k * 5;
    int n = 0;
    for (int r = lane; r < top_k; r += 32) {
        const float s = slab[5 * r];
        n += (s > 0.f && s > thresh) ? 1 : 0;
    }
    n = warp_sum(n);
    if (lane == 0) counts[warp] = n;
}

// exclusive scan of counts[B] -> offsets[B + 1] by one block
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t *__restrict__ counts, int B, int32_t *__restrict__ offsets) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int i = base + tid;
        const int v = i < B ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, w, o); if (lane >= o) w += t; }
            s_w[lane] = w;
        }
        __syncthreads();
        const int excl = s_carry + (warp ? s_w[warp - 1] : 0) + incl - v;
        if (i < B) offsets[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) offsets[B] = s_carry;
}

__global__ void __launch_bounds__(256) collect_write_kernel(const float *__restrict__ out, int B, int C, int top_k, int cls, float width,
                                                            float height, int first_id, const int32_t *__restrict__ offsets,
                                                            float *__restrict__ rows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float *slab = out + ((size_t)warp * C + cls) * top_k * 5;
    const int o0 = offsets[warp], n = offsets[warp + 1] - o0;
    for (int r = lane; r < n; r += 32) {
        float *dst = rows + (size_t)(o0 + r) * 6;
        dst[0] = (float)(first_id + warp);
        dst[1] = slab[5 * r];
        dst[2] = __fmul_rn(slab[5 * r + 1], width);  dst[3] = __fmul_rn(slab[5 * r + 2], height);      // detections[:, 1:] * scale
        dst[4] = __fmul_rn(slab[5 * r + 3], width);  dst[5] = __fmul_rn(slab[5 * r + 4], height);
    }
}

// ---- greedy TP / FP assignment per image ------------------------------------------------------------------------------------------
// One warp per image.  det rows (id, score, x1, y1, x2, y2) of the image are contiguous and in descending score (the order
// make_pred's global sort visits them in); lanes split the image's ground-truth boxes.  flags[t][j] live in shared memory.
constexpr int AP_MAX_GT = 128, AP_MAX_THR = 16;

__global__ void __launch_bounds__(128) ap_match_kernel(const float *__restrict__ rows, const int32_t *__restrict__ det_off,
                                                       const float *__restrict__ gt, const int32_t *__restrict__ gt_off, int n_img,
                                                       const double *__restrict__ thr, int n_iou, int n_iobb, int n_det,
                                                       uint8_t *__restrict__ tp) {
    __shared__ uint32_t s_flag[4][AP_MAX_THR][AP_MAX_GT / 32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x * 4 + w;
    if (img >= n_img) return;
    const int n_thr = n_iou + n_iobb;
    for (int i = lane; i < AP_MAX_THR * (AP_MAX_GT / 32); i += 32) (&s_flag[w][0][0])[i] = 0u;
    __syncwarp();
    const int d0 = det_off[img], d1 = det_off[img + 1], g0 = gt_off[img], G = gt_off[img + 1] - g0;
    for (int d = d0; d < d1; ++d) {
        const double bx0 = rows[6 * (size_t)d + 2], by0 = rows[6 * (size_t)d + 3], bx1 = rows[6 * (size_t)d + 4], by1 = rows[6 * (size_t)d + 5];
        const double area_d = (bx1 - bx0) * (by1 - by0);
        // np.max / np.argmax over the boxes of the image: first maximum
        double best_iou = -INFINITY, best_iobb = -INFINITY;
        int j_iou = 0x7fffffff, j_iobb = 0x7fffffff;
        for (int j = lane; j < G; j += 32) {
            const float *g = gt + 4 * (size_t)(g0 + j);
            const double gx0 = g[0], gy0 = g[1], gx1 = g[2], gy1 = g[3];
            const double iw = fmax(fmin(gx1, bx1) - fmax(gx0, bx0), 0.0), ih = fmax(fmin(gy1, by1) - fmax(gy0, by0), 0.0);
            const double inter = iw * ih;
            const double uni = area_d + (gx1 - gx0) * (gy1 - gy0) - inter;           // test_ap_iobb.py:268-270
            const double iou = inter / uni, iobb = inter / area_d;
            // a NaN never wins np.max unless... np.max propagates NaN; the reference's boxes have positive area, keep IEEE compares
            if (iou > best_iou) { best_iou = iou; j_iou = j; }
            if (iobb > best_iobb) { best_iobb = iobb; j_iobb = j; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double v = __shfl_xor_sync(FULL, best_iou, o); const int jj = __shfl_xor_sync(FULL, j_iou, o);
            if (v > best_iou || (v == best_iou && jj < j_iou)) { best_iou = v; j_iou = jj; }
            const double u = __shfl_xor_sync(FULL, best_iobb, o); const int ju = __shfl_xor_sync(FULL, j_iobb, o);
            if (u > best_iobb || (u == best_iobb && ju < j_iobb)) { best_iobb = u; j_iobb = ju; }
        }
        if (lane < n_thr) {
            const bool is_iou = lane < n_iou;
            const double ov = is_iou ? best_iou : best_iobb;
            const int jm = is_iou ? j_iou : j_iobb;
            uint8_t code = 0;                                                           // an image without boxes: neither TP nor FP (258)
            if (G > 0) {
                code = 2;                                                               // false positive unless ...
                if (ov > thr[lane]) {                                                   // 277 / 288: strict >
                    uint32_t &word = s_flag[w][lane][jm >> 5];
                    if (!((word >> (jm & 31)) & 1u)) { code = 1; word |= 1u << (jm & 31); }     // ... the box is detected for the first time
                }
            }
            tp[(size_t)lane * n_det + d] = code;
        }
        __syncwarp();
    }
}

// ---- stable LSD radix sort of (key, index) pairs, 8 bits per pass ---------------------------------------------------------------
constexpr int RS_TILE = 1024;

__global__ void __launch_bounds__(RS_TILE) rs_hist_kernel(const uint32_t *__restrict__ keys, int n, int shift, uint32_t *__restrict__ hist /* [256][blocks] */) {
    __shared__ uint32_t h[256];
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * RS_TILE + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    __syncthreads();
    if (threadIdx.x < 256) hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan over the digit-major histogram (one block; `total` entries)
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *__restrict__ hist, int total) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += 1024) {
        const int i = base + tid;
        const uint32_t v = i < total ? hist[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, w, o); if (lane >= o) w += t; }
            s_w[lane] = w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + (warp ? s_w[warp - 1] : 0u) + incl - v;
        if (i < total) hist[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_TILE) rs_scatter_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int n, int shift,
                                                             const uint32_t *__restrict__ hist, uint32_t *__restrict__ keys_out,
                                                             uint32_t *__restrict__ vals_out) {
    __shared__ uint16_t wcount[32][256];                                         // elements of digit d in warp w of this tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * 256; i += RS_TILE) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const int i = blockIdx.x * RS_TILE + tid;
    const bool ok = i < n;
    const uint32_t key = ok ? keys[i] : 0u;
    const uint32_t d = (key >> shift) & 255u;
    const unsigned act = __ballot_sync(FULL, ok);
    unsigned peers = 0;
    if (ok) {
        peers = __match_any_sync(act, d);
        if (lane == __ffs(peers) - 1) wcount[warp][d] = (uint16_t)__popc(peers);
    }
    __syncthreads();
    if (tid < 256) {                                                             // exclusive prefix over the warps, per digit
        uint32_t run = 0;
        for (int w = 0; w < 32; ++w) { const uint32_t c = wcount[w][tid]; wcount[w][tid] = (uint16_t)run; run += c; }
    }
    __syncthreads();
    if (ok) {
        const uint32_t pos = hist[(size_t)d * gridDim.x + blockIdx.x] + wcount[warp][d] + __popc(peers & ((1u << lane) - 1u));
        keys_out[pos] = key;
        vals_out[pos] = vals[i];
    }
}

__global__ void ap_keys_kernel(const float *__restrict__ rows, int n, uint32_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = ~f2ord(rows[6 * (size_t)i + 1]); idx[i] = (uint32_t)i; }   // ascending key = descending score
}

// ---- precision / recall and the AP of one threshold per block -------------------------------------------------------------------------
// block t: tp flags of threshold t visited in sorted order; cumulative sums carried over chunks of 1024
__global__ void __launch_bounds__(1024) ap_curve_kernel(const uint8_t *__restrict__ tp, const uint32_t *__restrict__ order, int n, double npos,
                                                        int use_07, const double *__restrict__ rec_thr /* [11] */, double *__restrict__ prec_ws /* [n_thr][n] */,
                                                        double *__restrict__ rec_ws, double *__restrict__ ap_out) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    __shared__ double s_red[32];
    __shared__ double s_run;
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *flags = tp + (size_t)t * n;
    double *prec = prec_ws + (size_t)t * n, *rec = rec_ws + (size_t)t * n;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    double pmax[11];
#pragma unroll
    for (int q = 0; q < 11; ++q) pmax[q] = 0.0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        // cumulative TP in the high, FP in the low 32 bits
        const uint32_t code = i < n ? (uint32_t)flags[order[i]] : 0u;
        const unsigned long long v = ((unsigned long long)(code == 1u) << 32) | (unsigned long long)(code == 2u);
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(FULL, w, o); if (lane >= o) w += u; }
            s_w[lane] = w;
        }
        __syncthreads();
        const unsigned long long tp_cum = s_carry + (warp ? s_w[warp - 1] : 0ull) + incl;
        if (i < n) {
            const double tpc = (double)(uint32_t)(tp_cum >> 32), fpc = (double)(uint32_t)tp_cum;
            const double r = tpc / npos;                                           // test_ap_iobb.py:304
            const double p = tpc / fmax(tpc + fpc, 2.220446049250313e-16);         // 307: np.finfo(np.float64).eps
            prec[i] = p; rec[i] = r;
            if (use_07) {
#pragma unroll
                for (int q = 0; q < 11; ++q) if (r >= rec_thr[q]) pmax[q] = fmax(pmax[q], p);   // voc_ap 17-22
            }
        }
        __syncthreads();
        if (tid == 1023) s_carry = tp_cum;
        __syncthreads();
    }
    if (use_07) {
        double ap = 0.0;
        for (int q = 0; q < 11; ++q) {
            double m = pmax[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
            if (lane == 0) s_red[warp] = m;
            __syncthreads();
            if (tid == 0) {
                double mm = 0.0;
                for (int w = 0; w < 32; ++w) mm = fmax(mm, s_red[w]);
                ap = ap + mm / 11.;                                                // voc_ap 23
            }
            __syncthreads();
        }
        if (tid == 0) ap_out[t] = ap;
        return;
    }
    // area under the precision envelope (voc_ap 24-40): mrec = [0, rec, 1], mpre = [0, prec, 0]; envelope = suffix maximum; sum over
    // the points where recall changes of (mrec[i+1] - mrec[i]) * mpre[i+1].  Walk backwards in chunks carrying the running maximum.
    if (tid == 0) s_run = 0.0;                                                   // mpre[n+1] = 0
    __syncthreads();
    double acc = 0.0;
    for (int top = n; top > 0; top -= 1024) {
        const int i = top - 1 - tid;                                               // this thread's index into prec / rec (descending)
        double v = i >= 0 ? prec[i] : 0.0;
        // inclusive suffix max in visiting order (thread 0 = highest index)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(FULL, v, o); if (lane >= o) v = fmax(v, u); }
        if (lane == 31) s_red[warp] = v;
        __syncthreads();
        if (warp == 0) {
            double w = s_red[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(FULL, w, o); if (lane >= o) w = fmax(w, u); }
            s_red[lane] = w;
        }
        __syncthreads();
        double env = fmax(v, s_run);
        if (warp) env = fmax(env, s_red[warp - 1]);
        // mrec index of prec[i] is i + 1; the term at position k = i (0-based in mrec[1:] != mrec[:-1]) is (mrec[i+1] - mrec[i]) * mpre[i+1]
        if (i >= 0) {
            const double r_here = rec[i], r_prev = i > 0 ? rec[i - 1] : 0.0;
            if (r_here != r_prev) acc += (r_here - r_prev) * env;
        }
        __syncthreads();
        if (tid == 1023 || i == 0) s_run = env;                                    // the lowest index of the chunk carries the maximum down
        __syncthreads();
    }
    // the last interval: mrec[n+1] = 1 against rec[n-1], with mpre[n+1] = 0 -> contributes 0
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < 32; ++w) s += s_red[w];
        ap_out[t] = s;
    }
}

}  // namespace gssd

using namespace gssd;

extern "C" int gssd_collect_detections(const float *detect_out, int B, int C, int top_k, int class_index, float width, float height,
                                       float thresh, int first_image_id, float *rows, int32_t *offsets /* [B+1] */,
                                       int32_t *counts_ws /* [B] */, void *stream) {
    if (!detect_out || !rows || !offsets || !counts_ws) return GSSD_ERR_ARG;
    if (B <= 0 || C <= 0 || top_k <= 0 || class_index < 0 || class_index >= C) return GSSD_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = ceil_div(B * 32, 256);
    collect_count_kernel<<<blocks, 256, 0, st>>>(detect_out, B, C, top_k, class_index, thresh, counts_ws);
    GSSD_AFTER_LAUNCH();
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts_ws, B, offsets);
    GSSD_AFTER_LAUNCH();
    collect_write_kernel<<<blocks, 256, 0, st>>>(detect_out, B, C, top_k, class_index, width, height, first_image_id, offsets, rows);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" size_t gssd_ap_workspace_bytes(int n_det, int n_thr) {
    if (n_det <= 0 || n_thr <= 0) return 0;
    const size_t n = (size_t)n_det, blocks = (n + RS_TILE - 1) / RS_TILE;
    // keys x2, idx x2, histogram, tp flags, prec + rec
    return 4 * n * 4 + 256 * blocks * 4 + (size_t)n_thr * n + 2 * (size_t)n_thr * n * 8 + 1024;
}

extern "C" int gssd_ap_eval(const float *rows, const int32_t *det_off, const float *gt_boxes, const int32_t *gt_off, int n_img, int n_det,
                            const double *thresholds /* device [n_iou + n_iobb] */, int n_iou, int n_iobb, int npos, int use_07_metric,
                            const double *rec_points /* device [11] */, double *ap_out /* device [n_iou + n_iobb] */,
                            uint8_t *tp_out /* device [n_thr][n_det] or NULL */, uint32_t *order_out /* device [n_det] or NULL */,
                            void *ws, size_t ws_bytes, void *stream) {
    if (!rows || !det_off || !gt_boxes || !gt_off || !thresholds || !ap_out || !ws || (use_07_metric && !rec_points)) return GSSD_ERR_ARG;
    const int n_thr = n_iou + n_iobb;
    if (n_img <= 0 || n_det <= 0 || n_iou < 0 || n_iobb < 0 || n_thr <= 0 || npos <= 0) return GSSD_ERR_ARG;
    if (n_thr > AP_MAX_THR) return GSSD_ERR_LIMIT;
    if (ws_bytes < gssd_ap_workspace_bytes(n_det, n_thr)) return GSSD_ERR_WS;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)n_det;
    const int blocks = (int)((n + RS_TILE - 1) / RS_TILE);
    uint8_t *base = reinterpret_cast<uint8_t *>(ws);
    uint32_t *k0 = reinterpret_cast<uint32_t *>(base), *k1 = k0 + n, *v0 = k1 + n, *v1 = v0 + n;
    uint32_t *hist = v1 + n;
    double *prec = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(hist + 256 * (size_t)blocks) + 15) & ~(uintptr_t)15);
    double *rec = prec + (size_t)n_thr * n;
    uint8_t *tp = reinterpret_cast<uint8_t *>(rec + (size_t)n_thr * n);
    // 1. TP / FP flags, image by image (test_ap_iobb.py:251-297)
    ap_match_kernel<<<ceil_div(n_img, 4), 128, 0, st>>>(rows, det_off, gt_boxes, gt_off, n_img, thresholds, n_iou, n_iobb, n_det, tp);
    GSSD_AFTER_LAUNCH();
    // 2. global descending-score order (make_pred, 213-223): stable, so equal scores keep (image, rank) order
    ap_keys_kernel<<<ceil_div(n_det, 256), 256, 0, st>>>(rows, n_det, k0, v0);
    GSSD_AFTER_LAUNCH();
    for (int pass = 0; pass < 4; ++pass) {
        rs_hist_kernel<<<blocks, RS_TILE, 0, st>>>(k0, n_det, 8 * pass, hist);
        GSSD_AFTER_LAUNCH();
        rs_scan_kernel<<<1, 1024, 0, st>>>(hist, 256 * blocks);
        GSSD_AFTER_LAUNCH();
        rs_scatter_kernel<<<blocks, RS_TILE, 0, st>>>(k0, v0, n_det, 8 * pass, hist, k1, v1);
        GSSD_AFTER_LAUNCH();
        uint32_t *t = k0; k0 = k1; k1 = t; t = v0; v0 = v1; v1 = t;
    }
    // 3. cumulative sums, precision / recall, AP per threshold (299-326, voc_ap)
    ap_curve_kernel<<<n_thr, 1024, 0, st>>>(tp, v0, n_det, (double)npos, use_07_metric, rec_points, prec, rec, ap_out);
    GSSD_AFTER_LAUNCH();
    if (tp_out) GSSD_RETURN_IF_CUDA(cudaMemcpyAsync(tp_out, tp, (size_t)n_thr * n, cudaMemcpyDeviceToDevice, st));
    if (order_out) GSSD_RETURN_IF_CUDA(cudaMemcpyAsync(order_out, v0, n * 4, cudaMemcpyDeviceToDevice, st));
    return GSSD_OK;
}
