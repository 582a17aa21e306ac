// bnrelu.cu — training-mode BatchNorm2d + ReLU of the backbone layers that stay NCHW fp32 (the "-> BN -> ReLU" after every grouped
// convolution of models/ssd_multiphase_custom_group.py:434-460, applied at :254-259 / 300-301), forward and backward.
//
// Why: in the reference's training step (batch 32, 300 x 300) these are the largest single cost — torch.profiler on a B200:
// cuDNN's batch-norm backward 13.0 ms + forward 4.8 ms + the separate ReLU / threshold_backward passes ~3 ms of a 53 ms step —
// although they are plain streaming work: a [32, 64, 300, 300] fp32 activation is 737 MB, the forward needs three passes over it
// (statistics; read + write) and the backward five (two reductions' inputs; x, dy, dx), 0.34 / 0.57 ms at the HBM roofline against
// 0.68 / 1.86 ms (+ the ReLU passes) measured for the library kernels.  HBM-bound: one CTA per (image, channel) plane, 16-byte
// loads and stores (scalar head / tail where a plane does not start or end on 16 bytes), fp32 partial sums per thread, double
// atomics per channel; the ReLU mask is recomputed from x in the backward (same fused multiply-add as the forward), so the
// activations are not read again.  Also here: the backward of the max pools between those layers as a gather (below).
#include "common.cuh"

namespace gssd {

constexpr int BNR_NT = 256;

__device__ __forceinline__ void block_sum2_to_double(float a, float b, double *dst) {
    __shared__ float sa[BNR_NT / 32], sb[BNR_NT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(FULL, a, o); b += __shfl_xor_sync(FULL, b, o); }
    if (lane == 0) { sa[warp] = a; sb[warp] = b; }
    __syncthreads();
    if (warp == 0) {
        double da = lane < BNR_NT / 32 ? (double)sa[lane] : 0.0, db = lane < BNR_NT / 32 ? (double)sb[lane] : 0.0;
#pragma unroll
        for (int o = 4; o; o >>= 1) { da += __shfl_xor_sync(FULL, da, o); db += __shfl_xor_sync(FULL, db, o); }
        if (lane == 0) { atomicAdd(dst, da); atomicAdd(dst + 1, db); }
    }
}

// y = x*a + b coefficients of channel c from the accumulated (sum, sum of squares)
struct BnCoef { float mean, rstd, a, b; };
__device__ __forceinline__ BnCoef bn_coef(const double *sums, int c, double inv_count, float eps, const float *gamma, const float *beta) {
    const double mean = sums[2 * c] * inv_count;
    double var = sums[2 * c + 1] * inv_count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    BnCoef k;
    k.mean = (float)mean;
    k.rstd = (float)(1.0 / sqrt(var + (double)eps));
    k.a = k.rstd * gamma[c];
    k.b = beta[c] - k.mean * k.a;
    return k;
}

// One plane = HW contiguous floats starting `off` floats into a 16-byte-aligned tensor: scalar head up to the next 16-byte boundary,
// float4 body, scalar tail (75 x 75 and 19 x 19 planes are not multiples of four floats: a scalar-only version ran those layers at a
// third of the bandwidth).  f1(i): element i; f4(i): elements i..i+3, i is 16-byte aligned.
template <typename F1, typename F4>
__device__ __forceinline__ void plane_loop(int HW, size_t off, bool aligned, F1 f1, F4 f4) {
    const int head = aligned ? min(HW, (int)((4 - (off & 3)) & 3)) : HW;
    for (int i = threadIdx.x; i < head; i += BNR_NT) f1(i);
    const int n4 = (HW - head) >> 2;
    for (int i = threadIdx.x; i < n4; i += BNR_NT) f4(head + 4 * i);
    for (int i = head + 4 * n4 + threadIdx.x; i < HW; i += BNR_NT) f1(i);
}

// ---- forward -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BNR_NT) bnr_stats_kernel(const float *__restrict__ x, int C, int HW, bool aligned, double *__restrict__ sums) {
    const size_t plane = blockIdx.x, off = plane * HW;
    const float *p = x + off;
    float s = 0.f, ss = 0.f;
    plane_loop(HW, off, aligned,
               [&](int i) { const float v = __ldg(p + i); s += v; ss += v * v; },
               [&](int i) {
                   const float4 v = __ldg(reinterpret_cast<const float4 *>(p + i));
                   s += (v.x + v.y) + (v.z + v.w);
                   ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
               });
    block_sum2_to_double(s, ss, sums + 2 * (plane % C));
}

__global__ void __launch_bounds__(BNR_NT) bnr_apply_kernel(const float *__restrict__ x, int C, int HW, bool aligned, const double *__restrict__ sums,
                                                           double inv_count, double unbias, float eps, const float *__restrict__ gamma,
                                                           const float *__restrict__ beta, int relu, float *__restrict__ y,
                                                           float *__restrict__ save, float *running_mean, float *running_var, float momentum,
                                                           const float *__restrict__ mean_shift) {
    const size_t plane = blockIdx.x, off = plane * HW;
    const int c = (int)(plane % C);
    const BnCoef k = bn_coef(sums, c, inv_count, eps, gamma, beta);
    if (plane < (size_t)C && threadIdx.x == 0) {                           // the planes of image 0 publish the channel's statistics
        save[2 * c] = k.mean; save[2 * c + 1] = k.rstd;
        if (running_mean != nullptr) {                                       // nn.BatchNorm2d: running = (1 - m)*running + m*batch, unbiased variance
            const double mean = sums[2 * c] * inv_count;
            double var = sums[2 * c + 1] * inv_count - mean * mean;
            var = var < 0.0 ? 0.0 : var;
            // mean_shift: the bias of the convolution in front, which the caller left out (it cancels in the normalisation)
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * ((float)mean + (mean_shift ? mean_shift[c] : 0.f));
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
        }
    }
    const float *p = x + off;
    float *q = y + off;
    const float lo = relu ? 0.f : -INFINITY;
    plane_loop(HW, off, aligned,
               [&](int i) { q[i] = fmaxf(fmaf(__ldcs(p + i), k.a, k.b), lo); },
               [&](int i) {
                   const float4 v = __ldcs(reinterpret_cast<const float4 *>(p + i));
                   *reinterpret_cast<float4 *>(q + i) = make_float4(fmaxf(fmaf(v.x, k.a, k.b), lo), fmaxf(fmaf(v.y, k.a, k.b), lo),
                                                                    fmaxf(fmaf(v.z, k.a, k.b), lo), fmaxf(fmaf(v.w, k.a, k.b), lo));
               });
}

// ---- backward ----------------------------------------------------------------------------------------------------
// g = dy where the forward's output was positive (recomputed: x*a + b > 0); sums: (sum g, sum g*x_hat) per channel
__global__ void __launch_bounds__(BNR_NT) bnr_bwd_reduce_kernel(const float *__restrict__ x, const float *__restrict__ dy, int C, int HW, bool aligned,
                                                                const float *__restrict__ save, const float *__restrict__ gamma,
                                                                const float *__restrict__ beta, int relu, double *__restrict__ sums) {
    const size_t plane = blockIdx.x, off = plane * HW;
    const int c = (int)(plane % C);
    const float mean = save[2 * c], rstd = save[2 * c + 1], a = rstd * gamma[c], b = beta[c] - mean * a;
    const float *p = x + off, *d = dy + off;
    float sg = 0.f, sgx = 0.f;
    auto one = [&](float xv, float dv) {
        const float g = (!relu || fmaf(xv, a, b) > 0.f) ? dv : 0.f;
        sg += g;
        sgx += g * ((xv - mean) * rstd);
    };
    plane_loop(HW, off, aligned,
               [&](int i) { one(__ldg(p + i), __ldg(d + i)); },
               [&](int i) {
                   const float4 v = __ldg(reinterpret_cast<const float4 *>(p + i)), w = __ldg(reinterpret_cast<const float4 *>(d + i));
                   one(v.x, w.x); one(v.y, w.y); one(v.z, w.z); one(v.w, w.w);
               });
    block_sum2_to_double(sg, sgx, sums + 2 * c);
}

// dx = gamma*rstd*(g - mean(g) - x_hat*mean(g*x_hat));  d_gamma = sum g*x_hat, d_beta = sum g
__global__ void __launch_bounds__(BNR_NT) bnr_bwd_apply_kernel(const float *__restrict__ x, const float *__restrict__ dy, int C, int HW, bool aligned,
                                                               const float *__restrict__ save, const float *__restrict__ gamma,
                                                               const float *__restrict__ beta, int relu, const double *__restrict__ sums,
                                                               double inv_count, float *__restrict__ dx, float *__restrict__ d_gamma,
                                                               float *__restrict__ d_beta) {
    const size_t plane = blockIdx.x, off = plane * HW;
    const int c = (int)(plane % C);
    const float mean = save[2 * c], rstd = save[2 * c + 1], a = rstd * gamma[c], b = beta[c] - mean * a;
    const float mg = (float)(sums[2 * c] * inv_count), mgx = (float)(sums[2 * c + 1] * inv_count);
    if (plane < (size_t)C && threadIdx.x == 0) { d_beta[c] = (float)sums[2 * c]; d_gamma[c] = (float)sums[2 * c + 1]; }
    const float *p = x + off, *d = dy + off;
    float *q = dx + off;
    auto one = [&](float xv, float dv) {
        const float g = (!relu || fmaf(xv, a, b) > 0.f) ? dv : 0.f;
        return a * (g - mg - ((xv - mean) * rstd) * mgx);
    };
    plane_loop(HW, off, aligned,
               [&](int i) { q[i] = one(__ldcs(p + i), __ldcs(d + i)); },
               [&](int i) {
                   const float4 v = __ldcs(reinterpret_cast<const float4 *>(p + i)), w = __ldcs(reinterpret_cast<const float4 *>(d + i));
                   __stcs(reinterpret_cast<float4 *>(q + i), make_float4(one(v.x, w.x), one(v.y, w.y), one(v.z, w.z), one(v.w, w.w)));
               });
}

// ---- nn.MaxPool2d backward on NCHW -------------------------------------------------------------------------------
// torch's max_pool_backward_nchw is the next largest streaming cost of the step (3.2 ms for the five pools at batch 32).  With the
// forward's argmax indices (flat h*W + w per output, what F.max_pool2d(..., return_indices=True) returns) the gradient is a
// scatter without collisions for the 2x2 / stride 2 pools — one thread per window writes its four input pixels, 8-byte stores
// along the row — and a gather for overlapping windows (3x3 / stride 1): an input pixel sums dy of the <= ceil(k/s)^2 windows
// that contain it and picked it.  No atomics, dx written exactly once.  (A first, gather-only version with 64-bit index
// arithmetic per element took 2.6 ms for the five pools.)
template <bool EVEN_W>
__global__ void __launch_bounds__(256) maxpool2x2_bwd_kernel(const float *__restrict__ dy, const long long *__restrict__ idx, int H, int W,
                                                             int OH, int OW, float *__restrict__ dx) {
    const int o = blockIdx.x * 256 + threadIdx.x;
    if (o >= OH * OW) return;
    const size_t plane = blockIdx.y;
    const int oy = o / OW, ox = o - oy * OW, y0 = 2 * oy, x0 = 2 * ox;
    const float g = __ldg(dy + plane * OH * OW + o);
    const int t = (int)(__ldg(idx + plane * OH * OW + o) - ((long long)y0 * W + x0));       // 0, 1, W or W + 1
    float *q = dx + plane * H * W + (size_t)y0 * W + x0;
    const bool row1 = y0 + 1 < H, col1 = x0 + 1 < W;                                           // ceil_mode windows may hang over the edge
    if (EVEN_W) {
        __stcs(reinterpret_cast<float2 *>(q), make_float2(t == 0 ? g : 0.f, t == 1 ? g : 0.f));
        if (row1) __stcs(reinterpret_cast<float2 *>(q + W), make_float2(t == W ? g : 0.f, t == W + 1 ? g : 0.f));
    } else {
        q[0] = t == 0 ? g : 0.f;
        if (col1) q[1] = t == 1 ? g : 0.f;
        if (row1) { q[W] = t == W ? g : 0.f; if (col1) q[W + 1] = t == W + 1 ? g : 0.f; }
    }
    // pixels no window covers (floor mode on odd extents) have no gradient
    if (ox == OW - 1)
        for (int x = 2 * OW; x < W; ++x) { dx[plane * H * W + (size_t)y0 * W + x] = 0.f; if (row1) dx[plane * H * W + (size_t)(y0 + 1) * W + x] = 0.f; }
    if (oy == OH - 1)
        for (int y = 2 * OH; y < H; ++y) {
            float *r = dx + plane * H * W + (size_t)y * W;
            r[x0] = 0.f; if (col1) r[x0 + 1] = 0.f;
            if (ox == OW - 1) for (int x = 2 * OW; x < W; ++x) r[x] = 0.f;
        }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float *__restrict__ dy, const long long *__restrict__ idx, int H, int W,
                                                          int OH, int OW, int k, int s, int pad, float *__restrict__ dx) {
    const int rem = blockIdx.x * 256 + threadIdx.x;
    if (rem >= H * W) return;
    const size_t plane = blockIdx.y;
    const int y = rem / W, x = rem - y * W;
    const int oy0 = y + pad - k + 1 > 0 ? (y + pad - k + s) / s : 0, oy1 = min(OH - 1, (y + pad) / s);      // ceil((y + pad - k + 1) / s) .. floor((y + pad) / s)
    const int ox0 = x + pad - k + 1 > 0 ? (x + pad - k + s) / s : 0, ox1 = min(OW - 1, (x + pad) / s);
    const float *d = dy + plane * OH * OW;
    const long long *ix = idx + plane * OH * OW;
    float g = 0.f;
    for (int oy = oy0; oy <= oy1; ++oy)
        for (int ox = ox0; ox <= ox1; ++ox)
            if (__ldg(ix + oy * OW + ox) == (long long)rem) g += __ldg(d + oy * OW + ox);
    __stcs(dx + plane * H * W + rem, g);
}

// ---- the same on channels-last (NHWC) tensors ---------------------------------------------------------------------
// With the backbone in torch.channels_last cuDNN runs its convolutions on NHWC tensor-core kernels and drops its NCHW <-> NHWC
// conversion passes (6 ms of the batch-32 step): the activations then are [rows = N*H*W, C] with the channels innermost.  A thread
// owns one channel quad (C/4 divides the CTA size, so a grid-stride loop over 16-byte elements never changes a thread's channels):
// coefficients live in registers, every access is a coalesced 16-byte load / store.
constexpr int BNH_NT = 256;

__device__ __forceinline__ void nhwc_reduce8_to_double(float (&v)[8], int q, int quad, double *sums /* [C][2] */) {
    __shared__ float red[BNH_NT][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = v[j];
    __syncthreads();
    if ((int)threadIdx.x < q) {                                   // thread `quad` adds the row groups of its channels
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = 0.f;
        for (int r = threadIdx.x; r < BNH_NT; r += q)
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] += red[r][j];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(sums + 2 * (4 * quad + j), (double)t[j]);
            atomicAdd(sums + 2 * (4 * quad + j) + 1, (double)t[4 + j]);
        }
    }
}

__global__ void __launch_bounds__(BNH_NT) bnh_stats_kernel(const float4 *__restrict__ x, int q, size_t n4, double *__restrict__ sums) {
    const int quad = threadIdx.x % q;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (size_t i = (size_t)blockIdx.x * BNH_NT + threadIdx.x; i < n4; i += (size_t)gridDim.x * BNH_NT) {
        const float4 t = __ldg(x + i);
        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
        v[4] += t.x * t.x; v[5] += t.y * t.y; v[6] += t.z * t.z; v[7] += t.w * t.w;
    }
    nhwc_reduce8_to_double(v, q, quad, sums);
}

__global__ void __launch_bounds__(BNH_NT) bnh_apply_kernel(const float4 *__restrict__ x, int C, size_t n4, const double *__restrict__ sums,
                                                           double inv_count, double unbias, float eps, const float *__restrict__ gamma,
                                                           const float *__restrict__ beta, int relu, float4 *__restrict__ y,
                                                           float *__restrict__ save, float *running_mean, float *running_var, float momentum,
                                                           const float *__restrict__ mean_shift) {
    const int q = C / 4, quad = threadIdx.x % q;
    BnCoef k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = bn_coef(sums, 4 * quad + j, inv_count, eps, gamma, beta);
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += BNH_NT) {
            const BnCoef kc = bn_coef(sums, c, inv_count, eps, gamma, beta);
            save[2 * c] = kc.mean; save[2 * c + 1] = kc.rstd;
            if (running_mean != nullptr) {
                const double mean = sums[2 * c] * inv_count;
                double var = sums[2 * c + 1] * inv_count - mean * mean;
                var = var < 0.0 ? 0.0 : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * ((float)mean + (mean_shift ? mean_shift[c] : 0.f));
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
            }
        }
    }
    const float lo = relu ? 0.f : -INFINITY;
    for (size_t i = (size_t)blockIdx.x * BNH_NT + threadIdx.x; i < n4; i += (size_t)gridDim.x * BNH_NT) {
        const float4 t = __ldcs(x + i);
        y[i] = make_float4(fmaxf(fmaf(t.x, k[0].a, k[0].b), lo), fmaxf(fmaf(t.y, k[1].a, k[1].b), lo),
                           fmaxf(fmaf(t.z, k[2].a, k[2].b), lo), fmaxf(fmaf(t.w, k[3].a, k[3].b), lo));
    }
}

struct BnhChan { float mean, rstd, a, b; };
__device__ __forceinline__ BnhChan bnh_chan(const float *save, const float *gamma, const float *beta, int c) {
    BnhChan k;
    k.mean = save[2 * c]; k.rstd = save[2 * c + 1]; k.a = k.rstd * gamma[c]; k.b = beta[c] - k.mean * k.a;
    return k;
}

__global__ void __launch_bounds__(BNH_NT) bnh_bwd_reduce_kernel(const float4 *__restrict__ x, const float4 *__restrict__ dy, int q, size_t n4,
                                                                const float *__restrict__ save, const float *__restrict__ gamma,
                                                                const float *__restrict__ beta, int relu, double *__restrict__ sums) {
    const int quad = threadIdx.x % q;
    BnhChan k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = bnh_chan(save, gamma, beta, 4 * quad + j);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (size_t i = (size_t)blockIdx.x * BNH_NT + threadIdx.x; i < n4; i += (size_t)gridDim.x * BNH_NT) {
        const float4 t = __ldg(x + i), d = __ldg(dy + i);
        const float xs[4] = {t.x, t.y, t.z, t.w}, ds[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float g = (!relu || fmaf(xs[j], k[j].a, k[j].b) > 0.f) ? ds[j] : 0.f;
            v[j] += g;
            v[4 + j] += g * ((xs[j] - k[j].mean) * k[j].rstd);
        }
    }
    nhwc_reduce8_to_double(v, q, quad, sums);
}

__global__ void __launch_bounds__(BNH_NT) bnh_bwd_apply_kernel(const float4 *__restrict__ x, const float4 *__restrict__ dy, int C, size_t n4,
                                                               const float *__restrict__ save, const float *__restrict__ gamma,
                                                               const float *__restrict__ beta, int relu, const double *__restrict__ sums,
                                                               double inv_count, float4 *__restrict__ dx, float *__restrict__ d_gamma,
                                                               float *__restrict__ d_beta) {
    const int q = C / 4, quad = threadIdx.x % q;
    BnhChan k[4];
    float mg[4], mgx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        k[j] = bnh_chan(save, gamma, beta, 4 * quad + j);
        mg[j] = (float)(sums[2 * (4 * quad + j)] * inv_count);
        mgx[j] = (float)(sums[2 * (4 * quad + j) + 1] * inv_count);
    }
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += BNH_NT) { d_beta[c] = (float)sums[2 * c]; d_gamma[c] = (float)sums[2 * c + 1]; }
    for (size_t i = (size_t)blockIdx.x * BNH_NT + threadIdx.x; i < n4; i += (size_t)gridDim.x * BNH_NT) {
        const float4 t = __ldcs(x + i), d = __ldcs(dy + i);
        const float xs[4] = {t.x, t.y, t.z, t.w}, ds[4] = {d.x, d.y, d.z, d.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float g = (!relu || fmaf(xs[j], k[j].a, k[j].b) > 0.f) ? ds[j] : 0.f;
            o[j] = k[j].a * (g - mg[j] - ((xs[j] - k[j].mean) * k[j].rstd) * mgx[j]);
        }
        __stcs(dx + i, make_float4(o[0], o[1], o[2], o[3]));
    }
}

// nn.MaxPool2d backward, channels last: one thread per (output pixel, channel quad) for the 2x2 / stride 2 pools, per (input pixel,
// channel quad) otherwise; indices are torch's (flat h*W + w inside the (image, channel) plane, whatever the memory format)
__global__ void __launch_bounds__(256) maxpool2x2_nhwc_bwd_kernel(const float4 *__restrict__ dy, const longlong4 *__restrict__ idx, int H, int W,
                                                                  int OH, int OW, int q, float4 *__restrict__ dx, size_t total) {
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const int quad = (int)(e % q);
        size_t o = e / q;
        const int ox = (int)(o % OW); o /= OW;
        const int oy = (int)(o % OH);
        const size_t n = o / OH;
        const float4 g = __ldg(dy + e);
        const longlong4 ix = idx[e];
        const int y0 = 2 * oy, x0 = 2 * ox;
        const long long base = (long long)y0 * W + x0;
        const int t0 = (int)(ix.x - base), t1 = (int)(ix.y - base), t2 = (int)(ix.z - base), t3 = (int)(ix.w - base);
        float4 *p = dx + ((n * H + y0) * W + x0) * q + quad;
        const bool row1 = y0 + 1 < H, col1 = x0 + 1 < W;
        auto pick = [&](int t) { return make_float4(t0 == t ? g.x : 0.f, t1 == t ? g.y : 0.f, t2 == t ? g.z : 0.f, t3 == t ? g.w : 0.f); };
        __stcs(p, pick(0));
        if (col1) __stcs(p + q, pick(1));
        if (row1) { __stcs(p + (size_t)W * q, pick(W)); if (col1) __stcs(p + (size_t)(W + 1) * q, pick(W + 1)); }
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);                    // pixels no window covers (floor mode on odd extents)
        if (ox == OW - 1)
            for (int x = 2 * OW; x < W; ++x) { dx[((n * H + y0) * W + x) * q + quad] = z; if (row1) dx[((n * H + y0 + 1) * W + x) * q + quad] = z; }
        if (oy == OH - 1)
            for (int y = 2 * OH; y < H; ++y) {
                dx[((n * H + y) * W + x0) * q + quad] = z; if (col1) dx[((n * H + y) * W + x0 + 1) * q + quad] = z;
                if (ox == OW - 1) for (int x = 2 * OW; x < W; ++x) dx[((n * H + y) * W + x) * q + quad] = z;
            }
    }
}

__global__ void __launch_bounds__(256) maxpool_nhwc_bwd_kernel(const float4 *__restrict__ dy, const longlong4 *__restrict__ idx, int H, int W,
                                                               int OH, int OW, int k, int s, int pad, int q, float4 *__restrict__ dx, size_t total) {
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const int quad = (int)(e % q);
        size_t o = e / q;
        const int x = (int)(o % W); o /= W;
        const int y = (int)(o % H);
        const size_t n = o / H;
        const long long me = (long long)y * W + x;
        const int oy0 = y + pad - k + 1 > 0 ? (y + pad - k + s) / s : 0, oy1 = min(OH - 1, (y + pad) / s);
        const int ox0 = x + pad - k + 1 > 0 ? (x + pad - k + s) / s : 0, ox1 = min(OW - 1, (x + pad) / s);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int oy = oy0; oy <= oy1; ++oy)
            for (int ox = ox0; ox <= ox1; ++ox) {
                const size_t w = ((n * OH + oy) * OW + ox) * q + quad;
                const longlong4 ix = idx[w];
                const float4 d = __ldg(dy + w);
                if (ix.x == me) g.x += d.x;
                if (ix.y == me) g.y += d.y;
                if (ix.z == me) g.z += d.z;
                if (ix.w == me) g.w += d.w;
            }
        __stcs(dx + e, g);
    }
}

static unsigned bnh_grid(size_t n4) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t want = (n4 + BNH_NT - 1) / BNH_NT;
    return (unsigned)(want < (size_t)sms * 16 ? want : (size_t)sms * 16);
}
static int bnh_check(long rows, int C, const void *a, const void *b, const void *c) {
    if (rows <= 0 || C <= 0) return GSSD_ERR_ARG;
    if (C % 4 || BNH_NT % (C / 4) || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15))
        return GSSD_ERR_LIMIT;                                       // C/4 must divide the CTA size: C in {4, 8, ..., 1024}, a power of two times 4
    return GSSD_OK;
}

static int bnr_check(int N, int C, int HW) {
    if (N <= 0 || C <= 0 || HW <= 0) return GSSD_ERR_ARG;
    if ((long)N * C > 2147483647l) return GSSD_ERR_LIMIT;
    return GSSD_OK;
}
static bool bnr_aligned(const void *a, const void *b, const void *c) {          // the tensors' bases (planes are handled by plane_loop)
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

}  // namespace gssd

using namespace gssd;

extern "C" int gssd_bn_relu_nchw_fwd(const float *x, const float *gamma, const float *beta, int N, int C, int HW, float eps, int relu,
                                     float *y, float *save_mean_rstd, float *running_mean, float *running_var, float momentum,
                                     const float *mean_shift, double *ws, void *stream) {
    if (!x || !gamma || !beta || !y || !save_mean_rstd || !ws) return GSSD_ERR_ARG;
    if ((running_mean == nullptr) != (running_var == nullptr)) return GSSD_ERR_ARG;
    int rc = bnr_check(N, C, HW);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
    const double count = (double)N * HW, unbias = count > 1 ? count / (count - 1) : 1.0;
    const unsigned planes = (unsigned)((long)N * C);
    const bool al = bnr_aligned(x, y, x);
    bnr_stats_kernel<<<planes, BNR_NT, 0, st>>>(x, C, HW, al, ws);
    GSSD_AFTER_LAUNCH();
    bnr_apply_kernel<<<planes, BNR_NT, 0, st>>>(x, C, HW, al, ws, 1.0 / count, unbias, eps, gamma, beta, relu, y, save_mean_rstd, running_mean,
                                                running_var, momentum, mean_shift);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_bn_relu_nchw_bwd(const float *x, const float *dy, const float *gamma, const float *beta, const float *save_mean_rstd,
                                     int N, int C, int HW, int relu, float *dx, float *d_gamma, float *d_beta, double *ws, void *stream) {
    if (!x || !dy || !gamma || !beta || !save_mean_rstd || !dx || !d_gamma || !d_beta || !ws) return GSSD_ERR_ARG;
    int rc = bnr_check(N, C, HW);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
    const double count = (double)N * HW;
    const unsigned planes = (unsigned)((long)N * C);
    const bool al = bnr_aligned(x, dy, dx);
    bnr_bwd_reduce_kernel<<<planes, BNR_NT, 0, st>>>(x, dy, C, HW, al, save_mean_rstd, gamma, beta, relu, ws);
    GSSD_AFTER_LAUNCH();
    bnr_bwd_apply_kernel<<<planes, BNR_NT, 0, st>>>(x, dy, C, HW, al, save_mean_rstd, gamma, beta, relu, ws, 1.0 / count, dx, d_gamma, d_beta);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_maxpool_nchw_bwd(const float *dy, const int64_t *indices, int planes, int H, int W, int OH, int OW, int kernel, int stride,
                                     int pad, float *dx, void *stream) {
    if (!dy || !indices || !dx) return GSSD_ERR_ARG;
    if (planes <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || kernel <= 0 || stride <= 0 || pad < 0 || 2 * pad > kernel) return GSSD_ERR_ARG;
    if (planes > 65535 * 32) return GSSD_ERR_LIMIT;
    cudaStream_t st = (cudaStream_t)stream;
    const long long *ix = reinterpret_cast<const long long *>(indices);
    // gridDim.y is limited to 65535: planes beyond that go into further launches
    for (int p0 = 0; p0 < planes; p0 += 65535) {
        const int np = planes - p0 < 65535 ? planes - p0 : 65535;
        const float *dyp = dy + (size_t)p0 * OH * OW;
        const long long *ixp = ix + (size_t)p0 * OH * OW;
        float *dxp = dx + (size_t)p0 * H * W;
        if (kernel == 2 && stride == 2 && pad == 0) {
            const dim3 grid((OH * OW + 255) / 256, np);
            if ((W & 1) == 0 && (reinterpret_cast<uintptr_t>(dxp) & 7) == 0) maxpool2x2_bwd_kernel<true><<<grid, 256, 0, st>>>(dyp, ixp, H, W, OH, OW, dxp);
            else maxpool2x2_bwd_kernel<false><<<grid, 256, 0, st>>>(dyp, ixp, H, W, OH, OW, dxp);
        } else {
            maxpool_bwd_kernel<<<dim3((H * W + 255) / 256, np), 256, 0, st>>>(dyp, ixp, H, W, OH, OW, kernel, stride, pad, dxp);
        }
        GSSD_AFTER_LAUNCH();
    }
    return GSSD_OK;
}

extern "C" int gssd_bn_relu_nhwc_fwd(const float *x, const float *gamma, const float *beta, long rows, int C, float eps, int relu, float *y,
                                     float *save_mean_rstd, float *running_mean, float *running_var, float momentum, const float *mean_shift, double *ws,
                                     void *stream) {
    if (!x || !gamma || !beta || !y || !save_mean_rstd || !ws) return GSSD_ERR_ARG;
    if ((running_mean == nullptr) != (running_var == nullptr)) return GSSD_ERR_ARG;
    int rc = bnh_check(rows, C, x, y, x);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
    const size_t n4 = (size_t)rows * (C / 4);
    const double count = (double)rows, unbias = count > 1 ? count / (count - 1) : 1.0;
    const unsigned grid = bnh_grid(n4);
    bnh_stats_kernel<<<grid, BNH_NT, 0, st>>>(reinterpret_cast<const float4 *>(x), C / 4, n4, ws);
    GSSD_AFTER_LAUNCH();
    bnh_apply_kernel<<<grid, BNH_NT, 0, st>>>(reinterpret_cast<const float4 *>(x), C, n4, ws, 1.0 / count, unbias, eps, gamma, beta, relu,
                                              reinterpret_cast<float4 *>(y), save_mean_rstd, running_mean, running_var, momentum, mean_shift);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_bn_relu_nhwc_bwd(const float *x, const float *dy, const float *gamma, const float *beta, const float *save_mean_rstd,
                                     long rows, int C, int relu, float *dx, float *d_gamma, float *d_beta, double *ws, void *stream) {
    if (!x || !dy || !gamma || !beta || !save_mean_rstd || !dx || !d_gamma || !d_beta || !ws) return GSSD_ERR_ARG;
    int rc = bnh_check(rows, C, x, dy, dx);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
    const size_t n4 = (size_t)rows * (C / 4);
    const unsigned grid = bnh_grid(n4);
    bnh_bwd_reduce_kernel<<<grid, BNH_NT, 0, st>>>(reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(dy), C / 4, n4,
                                                   save_mean_rstd, gamma, beta, relu, ws);
    GSSD_AFTER_LAUNCH();
    bnh_bwd_apply_kernel<<<grid, BNH_NT, 0, st>>>(reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(dy), C, n4, save_mean_rstd,
                                                  gamma, beta, relu, ws, 1.0 / (double)rows, reinterpret_cast<float4 *>(dx), d_gamma, d_beta);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_maxpool_nhwc_bwd(const float *dy, const int64_t *indices, int N, int C, int H, int W, int OH, int OW, int kernel, int stride,
                                     int pad, float *dx, void *stream) {
    if (!dy || !indices || !dx) return GSSD_ERR_ARG;
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || kernel <= 0 || stride <= 0 || pad < 0 || 2 * pad > kernel) return GSSD_ERR_ARG;
    if (C % 4 || ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) || (reinterpret_cast<uintptr_t>(indices) & 31)) return GSSD_ERR_LIMIT;
    const int q = C / 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (kernel == 2 && stride == 2 && pad == 0) {
        const size_t total = (size_t)N * OH * OW * q;
        maxpool2x2_nhwc_bwd_kernel<<<bnh_grid(total), 256, 0, st>>>(reinterpret_cast<const float4 *>(dy), reinterpret_cast<const longlong4 *>(indices),
                                                                   H, W, OH, OW, q, reinterpret_cast<float4 *>(dx), total);
    } else {
        const size_t total = (size_t)N * H * W * q;
        maxpool_nhwc_bwd_kernel<<<bnh_grid(total), 256, 0, st>>>(reinterpret_cast<const float4 *>(dy), reinterpret_cast<const longlong4 *>(indices),
                                                                H, W, OH, OW, kernel, stride, pad, q, reinterpret_cast<float4 *>(dx), total);
    }
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}
