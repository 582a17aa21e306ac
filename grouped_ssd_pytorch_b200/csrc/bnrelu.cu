// bnrelu.cu — training-mode BatchNorm2d + ReLU of the backbone layers that stay NCHW fp32 (the "-> BN -> ReLU" after every grouped
// convolution of models/ssd_multiphase_custom_group.py:434-460, applied at :254-259 / 300-301), forward and backward.
//
// Why: in the reference's training step (batch 32, 300 x 300) these are the largest single cost — torch.profiler on a B200:
// cuDNN's batch-norm backward 13.0 ms + forward 4.8 ms + the separate ReLU / threshold_backward passes ~3 ms of a 53 ms step —
// although they are plain streaming work: a [32, 64, 300, 300] fp32 activation is 737 MB, the forward needs three passes over it
// (statistics; read + write) and the backward five (two reductions' inputs; x, dy, dx), 0.34 / 0.57 ms at the HBM roofline against
// 0.68 / 1.86 ms (+ the ReLU passes) measured for the library kernels.  HBM-bound: one CTA per (image, channel) plane, 16-byte
// loads and stores, fp32 partial sums per thread, double atomics per channel; the ReLU mask is recomputed from x in the backward
// (same fused multiply-add as the forward), so the activations are not read again.
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

// ---- forward -----------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(BNR_NT) bnr_stats_kernel(const float *__restrict__ x, int C, int HW, double *__restrict__ sums) {
    const size_t plane = blockIdx.x;
    const float *p = x + plane * HW;
    float s = 0.f, ss = 0.f;
    if (VEC) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        for (int i = threadIdx.x; i < HW / 4; i += BNR_NT) {
            const float4 v = __ldg(p4 + i);
            s += (v.x + v.y) + (v.z + v.w);
            ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
    } else {
        for (int i = threadIdx.x; i < HW; i += BNR_NT) { const float v = __ldg(p + i); s += v; ss += v * v; }
    }
    block_sum2_to_double(s, ss, sums + 2 * (plane % C));
}

template <bool VEC>
__global__ void __launch_bounds__(BNR_NT) bnr_apply_kernel(const float *__restrict__ x, int C, int HW, const double *__restrict__ sums,
                                                           double inv_count, double unbias, float eps, const float *__restrict__ gamma,
                                                           const float *__restrict__ beta, int relu, float *__restrict__ y,
                                                           float *__restrict__ save, float *running_mean, float *running_var, float momentum) {
    const size_t plane = blockIdx.x;
    const int c = (int)(plane % C);
    const BnCoef k = bn_coef(sums, c, inv_count, eps, gamma, beta);
    if (plane < (size_t)C && threadIdx.x == 0) {                           // the planes of image 0 publish the channel's statistics
        save[2 * c] = k.mean; save[2 * c + 1] = k.rstd;
        if (running_mean != nullptr) {                                       // nn.BatchNorm2d: running = (1 - m)*running + m*batch, unbiased variance
            const double mean = sums[2 * c] * inv_count;
            double var = sums[2 * c + 1] * inv_count - mean * mean;
            var = var < 0.0 ? 0.0 : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
        }
    }
    const float *p = x + plane * HW;
    float *q = y + plane * HW;
    const float lo = relu ? 0.f : -INFINITY;
    if (VEC) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        float4 *q4 = reinterpret_cast<float4 *>(q);
        for (int i = threadIdx.x; i < HW / 4; i += BNR_NT) {
            const float4 v = __ldcs(p4 + i);
            float4 o;
            o.x = fmaxf(fmaf(v.x, k.a, k.b), lo); o.y = fmaxf(fmaf(v.y, k.a, k.b), lo);
            o.z = fmaxf(fmaf(v.z, k.a, k.b), lo); o.w = fmaxf(fmaf(v.w, k.a, k.b), lo);
            q4[i] = o;
        }
    } else {
        for (int i = threadIdx.x; i < HW; i += BNR_NT) q[i] = fmaxf(fmaf(__ldcs(p + i), k.a, k.b), lo);
    }
}

// ---- backward ----------------------------------------------------------------------------------------------------
// g = dy where the forward's output was positive (recomputed: x*a + b > 0); sums: (sum g, sum g*x_hat) per channel
template <bool VEC>
__global__ void __launch_bounds__(BNR_NT) bnr_bwd_reduce_kernel(const float *__restrict__ x, const float *__restrict__ dy, int C, int HW,
                                                                const float *__restrict__ save, const float *__restrict__ gamma,
                                                                const float *__restrict__ beta, int relu, double *__restrict__ sums) {
    const size_t plane = blockIdx.x;
    const int c = (int)(plane % C);
    const float mean = save[2 * c], rstd = save[2 * c + 1], a = rstd * gamma[c], b = beta[c] - mean * a;
    const float *p = x + plane * HW, *d = dy + plane * HW;
    float sg = 0.f, sgx = 0.f;
    auto one = [&](float xv, float dv) {
        const float g = (!relu || fmaf(xv, a, b) > 0.f) ? dv : 0.f;
        sg += g;
        sgx += g * ((xv - mean) * rstd);
    };
    if (VEC) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p), *d4 = reinterpret_cast<const float4 *>(d);
        for (int i = threadIdx.x; i < HW / 4; i += BNR_NT) {
            const float4 v = __ldg(p4 + i), w = __ldg(d4 + i);
            one(v.x, w.x); one(v.y, w.y); one(v.z, w.z); one(v.w, w.w);
        }
    } else {
        for (int i = threadIdx.x; i < HW; i += BNR_NT) one(__ldg(p + i), __ldg(d + i));
    }
    block_sum2_to_double(sg, sgx, sums + 2 * c);
}

// dx = gamma*rstd*(g - mean(g) - x_hat*mean(g*x_hat));  d_gamma = sum g*x_hat, d_beta = sum g
template <bool VEC>
__global__ void __launch_bounds__(BNR_NT) bnr_bwd_apply_kernel(const float *__restrict__ x, const float *__restrict__ dy, int C, int HW,
                                                               const float *__restrict__ save, const float *__restrict__ gamma,
                                                               const float *__restrict__ beta, int relu, const double *__restrict__ sums,
                                                               double inv_count, float *__restrict__ dx, float *__restrict__ d_gamma,
                                                               float *__restrict__ d_beta) {
    const size_t plane = blockIdx.x;
    const int c = (int)(plane % C);
    const float mean = save[2 * c], rstd = save[2 * c + 1], a = rstd * gamma[c], b = beta[c] - mean * a;
    const float mg = (float)(sums[2 * c] * inv_count), mgx = (float)(sums[2 * c + 1] * inv_count);
    if (plane < (size_t)C && threadIdx.x == 0) { d_beta[c] = (float)sums[2 * c]; d_gamma[c] = (float)sums[2 * c + 1]; }
    const float *p = x + plane * HW, *d = dy + plane * HW;
    float *q = dx + plane * HW;
    auto one = [&](float xv, float dv) {
        const float g = (!relu || fmaf(xv, a, b) > 0.f) ? dv : 0.f;
        return a * (g - mg - ((xv - mean) * rstd) * mgx);
    };
    if (VEC) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p), *d4 = reinterpret_cast<const float4 *>(d);
        float4 *q4 = reinterpret_cast<float4 *>(q);
        for (int i = threadIdx.x; i < HW / 4; i += BNR_NT) {
            const float4 v = __ldcs(p4 + i), w = __ldcs(d4 + i);
            __stcs(q4 + i, make_float4(one(v.x, w.x), one(v.y, w.y), one(v.z, w.z), one(v.w, w.w)));
        }
    } else {
        for (int i = threadIdx.x; i < HW; i += BNR_NT) q[i] = one(__ldcs(p + i), __ldcs(d + i));
    }
}

static int bnr_check(int N, int C, int HW) {
    if (N <= 0 || C <= 0 || HW <= 0) return GSSD_ERR_ARG;
    if ((long)N * C > 2147483647l) return GSSD_ERR_LIMIT;
    return GSSD_OK;
}
static bool bnr_vec(int HW, const void *a, const void *b, const void *c) {
    return (HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

}  // namespace gssd

using namespace gssd;

extern "C" int gssd_bn_relu_nchw_fwd(const float *x, const float *gamma, const float *beta, int N, int C, int HW, float eps, int relu,
                                     float *y, float *save_mean_rstd, float *running_mean, float *running_var, float momentum,
                                     double *ws, void *stream) {
    if (!x || !gamma || !beta || !y || !save_mean_rstd || !ws) return GSSD_ERR_ARG;
    if ((running_mean == nullptr) != (running_var == nullptr)) return GSSD_ERR_ARG;
    int rc = bnr_check(N, C, HW);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
    const double count = (double)N * HW, unbias = count > 1 ? count / (count - 1) : 1.0;
    const unsigned planes = (unsigned)((long)N * C);
    if (bnr_vec(HW, x, y, x)) {
        bnr_stats_kernel<true><<<planes, BNR_NT, 0, st>>>(x, C, HW, ws);
        GSSD_AFTER_LAUNCH();
        bnr_apply_kernel<true><<<planes, BNR_NT, 0, st>>>(x, C, HW, ws, 1.0 / count, unbias, eps, gamma, beta, relu, y, save_mean_rstd,
                                                          running_mean, running_var, momentum);
    } else {
        bnr_stats_kernel<false><<<planes, BNR_NT, 0, st>>>(x, C, HW, ws);
        GSSD_AFTER_LAUNCH();
        bnr_apply_kernel<false><<<planes, BNR_NT, 0, st>>>(x, C, HW, ws, 1.0 / count, unbias, eps, gamma, beta, relu, y, save_mean_rstd,
                                                           running_mean, running_var, momentum);
    }
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
    if (bnr_vec(HW, x, dy, dx)) {
        bnr_bwd_reduce_kernel<true><<<planes, BNR_NT, 0, st>>>(x, dy, C, HW, save_mean_rstd, gamma, beta, relu, ws);
        GSSD_AFTER_LAUNCH();
        bnr_bwd_apply_kernel<true><<<planes, BNR_NT, 0, st>>>(x, dy, C, HW, save_mean_rstd, gamma, beta, relu, ws, 1.0 / count, dx, d_gamma, d_beta);
    } else {
        bnr_bwd_reduce_kernel<false><<<planes, BNR_NT, 0, st>>>(x, dy, C, HW, save_mean_rstd, gamma, beta, relu, ws);
        GSSD_AFTER_LAUNCH();
        bnr_bwd_apply_kernel<false><<<planes, BNR_NT, 0, st>>>(x, dy, C, HW, save_mean_rstd, gamma, beta, relu, ws, 1.0 / count, dx, d_gamma, d_beta);
    }
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}
