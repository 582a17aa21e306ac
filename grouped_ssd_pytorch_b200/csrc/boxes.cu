// boxes.cu — the small box_utils entry points (box_utils.py:4-67, 114-168), PriorBox
// (prior_box.py:32-172) and L2Norm (l2norm.py:19-23).  One thread per box / per cell / per pixel.
#include "common.cuh"

namespace gssd {

enum { OP_POINT_FORM, OP_CENTER_SIZE };

template <int OP>
__global__ void __launch_bounds__(256) unary_box_kernel(const float4 *in, int n, float4 *out) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float4 b = in[i];
    if (OP == OP_POINT_FORM) out[i] = point_form(b);
    else out[i] = make_float4(__fmul_rn(__fadd_rn(b.z, b.x), 0.5f), __fmul_rn(__fadd_rn(b.w, b.y), 0.5f),
                              __fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));            // box_utils.py:24-25 intent
}

template <bool IOU>
__global__ void __launch_bounds__(256) pair_kernel(const float4 *a, int A, const float4 *b, int Bn, float *out) {
    int j = blockIdx.x * 256 + threadIdx.x;
    int i = blockIdx.y;
    if (j >= Bn) return;
    float4 x = a[i], y = b[j];
    out[(size_t)i * Bn + j] = IOU ? box_iou_exact(x, box_area(x), y, box_area(y)) : box_inter(x, y);
}

template <bool ENC>
__global__ void __launch_bounds__(256) codec_kernel(const float4 *x, const float4 *priors, int n, float v0, float v1,
                                                    float4 *out) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    out[i] = ENC ? encode_box(x[i], priors[i], v0, v1) : decode_box(x[i], priors[i], v0, v1);
}

// ---- log_sum_exp with the tensor-wide max (box_utils.py:160-168) ---------------------------------
__global__ void __launch_bounds__(256) tensor_max_kernel(const float *x, size_t n, uint32_t *out_ord) {
    float m = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) m = fmaxf(m, x[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_ord, f2ord(m));
}

__global__ void __launch_bounds__(256) lse_kernel(const float *x, int rows, int C, const uint32_t *max_ord, float *out) {
    int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= rows) return;
    const float x_max = ord2f(*max_ord);
    const float *row = x + (size_t)r * C;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = __fadd_rn(s, expf(__fsub_rn(row[c], x_max)));
    out[r] = __fadd_rn(logf(s), x_max);
}

// ---- PriorBox ------------------------------------------------------------------------------------------
struct PriorLaunch {
    gssd_prior_cfg cfg;
    int cell_off[GSSD_MAX_FEATURE_MAPS + 1];   // first cell of each map
    int box_off[GSSD_MAX_FEATURE_MAPS + 1];    // first box of each map
    int per_cell[GSSD_MAX_FEATURE_MAPS];
};

__device__ __forceinline__ float clip01(float v, int clip) { return clip ? fminf(fmaxf(v, 0.f), 1.f) : v; }

__global__ void __launch_bounds__(128) priorbox_kernel(const __grid_constant__ PriorLaunch L, float4 *out) {
    const gssd_prior_cfg &c = L.cfg;
    int cell = blockIdx.x * 128 + threadIdx.x;
    if (cell >= L.cell_off[c.n_maps]) return;
    int k = 0;
    while (cell >= L.cell_off[k + 1]) ++k;
    const int f = c.feature_maps[k];
    const int local = cell - L.cell_off[k];
    const int i = local / f, j = local - i * f;                    // product(range(f), repeat=2): i row, j col
    float4 *o = out + L.box_off[k] + (size_t)local * L.per_cell[k];
    const int clip = c.clip;
#define EMIT(a_, b_, c_, d_) do { *o++ = make_float4(clip01((float)(a_), clip), clip01((float)(b_), clip), \
                                                      clip01((float)(c_), clip), clip01((float)(d_), clip)); } while (0)
    if (c.version != GSSD_PRIOR_LEGACY) {
        const double f_k = __ddiv_rn(c.min_dim, c.steps[k]);                       // prior_box.py:38
        const double cx = __ddiv_rn(j + 0.5, f_k), cy = __ddiv_rn(i + 0.5, f_k);   // 40-41
        const double s_k = __ddiv_rn(c.min_sizes[k], c.min_dim);                   // 45
        EMIT(cx, cy, s_k, s_k);
        const double s_kp = __dsqrt_rn(__dmul_rn(s_k, __ddiv_rn(c.max_sizes[k], c.min_dim)));   // 50
        EMIT(cx, cy, s_kp, s_kp);
        for (int a = 0; a < c.n_ar[k]; ++a) {
            const double r = __dsqrt_rn(c.aspect_ratios[k][a]);
            const double big = __dmul_rn(s_k, r), small = __ddiv_rn(s_k, r);
            if (c.version == GSSD_PRIOR_V2) { EMIT(cx, cy, big, small); EMIT(cx, cy, small, big); }   // 54-56
            else { EMIT(cx, cy, big, big); EMIT(cx, cy, small, small); }                            // 84-85
        }
    } else {                                                                       // 141-167
        const double step = __ddiv_rn(c.min_dim, (double)f);
        const double c_x = __dmul_rn(j + 0.5, step), c_y = __dmul_rn(i + 0.5, step);
        const double s = c.min_dim;
        double c_w = __ddiv_rn(c.min_sizes[k], 2.0), c_h = c_w;
#define CORNER() EMIT(__ddiv_rn(__dsub_rn(c_x, c_w), s), __ddiv_rn(__dsub_rn(c_y, c_h), s), \
                      __ddiv_rn(__dadd_rn(c_x, c_w), s), __ddiv_rn(__dadd_rn(c_y, c_h), s))
        CORNER();
        if (c.max_sizes[k] > 0) {
            c_w = c_h = __ddiv_rn(__dsqrt_rn(__dmul_rn(c.min_sizes[k], c.max_sizes[k])), 2.0);
            CORNER();
        }
        for (int a = 0; a < c.n_ar[k]; ++a) {
            const double ar = c.aspect_ratios[k][a];
            if (!(fabs(ar - 1) < 1e-6)) {
                c_w = __ddiv_rn(__dmul_rn(c.min_sizes[k], __dsqrt_rn(ar)), 2.0);
                c_h = __ddiv_rn(__ddiv_rn(c.min_sizes[k], __dsqrt_rn(ar)), 2.0);
                CORNER();
            }
        }
#undef CORNER
    }
#undef EMIT
}

static int prior_cfg_check(const gssd_prior_cfg *c) {
    if (!c || c->n_maps <= 0 || c->n_maps > GSSD_MAX_FEATURE_MAPS) return GSSD_ERR_ARG;
    if (c->version < GSSD_PRIOR_V2 || c->version > GSSD_PRIOR_LEGACY) return GSSD_ERR_ARG;
    for (int k = 0; k < c->n_maps; ++k)
        if (c->n_ar[k] < 0 || c->n_ar[k] > GSSD_MAX_ASPECT_RATIOS || c->feature_maps[k] <= 0) return GSSD_ERR_ARG;
    for (int i = 0; i < 2; ++i)
        if (c->variance[i] <= 0) return GSSD_ERR_VALUE;              // prior_box.py:28-30
    return GSSD_OK;
}

static int boxes_per_cell(const gssd_prior_cfg *c, int k) {
    if (c->version != GSSD_PRIOR_LEGACY) return 2 + 2 * c->n_ar[k];
    int n = 1 + (c->max_sizes[k] > 0 ? 1 : 0);
    for (int a = 0; a < c->n_ar[k]; ++a) {
        double d = c->aspect_ratios[k][a] - 1;
        if (!((d < 0 ? -d : d) < 1e-6)) ++n;
    }
    return n;
}

// ---- L2Norm ---------------------------------------------------------------------------------------------
constexpr int L2_PX = 32, L2_CG = 8;     // 32 pixels x 8 channel groups per CTA

__global__ void __launch_bounds__(L2_PX * L2_CG) l2norm_fwd_kernel(const float *x, const float *w, int Cn, int HW,
                                                                   float eps, float *y, float *norm_out) {
    __shared__ float red[L2_CG][L2_PX];
    const int tx = threadIdx.x & (L2_PX - 1), ty = threadIdx.x / L2_PX;
    const int px = blockIdx.x * L2_PX + tx, b = blockIdx.y;
    const bool ok = px < HW;
    const float *xb = x + (size_t)b * Cn * HW + px;
    float s = 0.f;
    if (ok) for (int c = ty; c < Cn; c += L2_CG) { float v = xb[(size_t)c * HW]; s += v * v; }
    red[ty][tx] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int g = 0; g < L2_CG; ++g) tot += red[g][tx];
    const float norm = __fadd_rn(sqrtf(tot), eps);                   // l2norm.py:20
    if (!ok) return;
    if (ty == 0 && norm_out) norm_out[(size_t)b * HW + px] = norm;
    float *yb = y + (size_t)b * Cn * HW + px;
    for (int c = ty; c < Cn; c += L2_CG)
        yb[(size_t)c * HW] = __fmul_rn(w[c], __fdiv_rn(xb[(size_t)c * HW], norm));   // 21-22
}

// gx = w*gy/n - x * D / (n^2 * s),  D = sum_c w*gy*x, s = n - eps;  gw partial = sum_px gy*x/n
__global__ void __launch_bounds__(L2_PX * L2_CG) l2norm_bwd_kernel(const float *x, const float *w, const float *norm,
                                                                   const float *gy, int Cn, int HW, float eps,
                                                                   float *gx, float *gw_part) {
    __shared__ float red[L2_CG][L2_PX];
    const int tx = threadIdx.x & (L2_PX - 1), ty = threadIdx.x / L2_PX;
    const int px = blockIdx.x * L2_PX + tx, b = blockIdx.y;
    const bool ok = px < HW;
    const size_t base = (size_t)b * Cn * HW + px;
    const float n = ok ? norm[(size_t)b * HW + px] : 1.f;
    float d = 0.f;
    if (ok) for (int c = ty; c < Cn; c += L2_CG) d += w[c] * gy[base + (size_t)c * HW] * x[base + (size_t)c * HW];
    red[ty][tx] = d;
    __syncthreads();
    float D = 0.f;
#pragma unroll
    for (int g = 0; g < L2_CG; ++g) D += red[g][tx];
    const float s = n - eps;
    const float k2 = D / (n * n * s);
    float *part = gw_part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * Cn;
    for (int c = ty; c < Cn; c += L2_CG) {
        float g = 0.f, xv = 0.f;
        if (ok) {
            g = gy[base + (size_t)c * HW]; xv = x[base + (size_t)c * HW];
            gx[base + (size_t)c * HW] = w[c] * g / n - xv * k2;
        }
        float t = ok ? g * xv / n : 0.f;                              // reduce over the 32 pixels of the warp
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
        if (tx == 0) part[c] = t;
    }
}

__global__ void __launch_bounds__(256) l2norm_gw_reduce_kernel(const float *part, int n_part, int Cn, float *gw) {
    int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= Cn) return;
    double s = 0.0;
    for (int i = 0; i < n_part; ++i) s += part[(size_t)i * Cn + c];
    gw[c] = (float)s;
}

}  // namespace gssd

using namespace gssd;

#define ST(s) ((cudaStream_t)(s))

extern "C" int gssd_priorbox_count(const gssd_prior_cfg *c) {
    int rc = prior_cfg_check(c);
    if (rc) return rc;
    long total = 0;
    for (int k = 0; k < c->n_maps; ++k) total += (long)c->feature_maps[k] * c->feature_maps[k] * boxes_per_cell(c, k);
    return (int)total;
}

extern "C" int gssd_priorbox(const gssd_prior_cfg *c, float *out, void *stream) {
    int rc = prior_cfg_check(c);
    if (rc) return rc;
    if (!out) return GSSD_ERR_ARG;
    PriorLaunch L;
    L.cfg = *c;
    L.cell_off[0] = 0; L.box_off[0] = 0;
    for (int k = 0; k < c->n_maps; ++k) {
        int cells = c->feature_maps[k] * c->feature_maps[k];
        L.per_cell[k] = boxes_per_cell(c, k);
        L.cell_off[k + 1] = L.cell_off[k] + cells;
        L.box_off[k + 1] = L.box_off[k] + cells * L.per_cell[k];
    }
    int cells = L.cell_off[c->n_maps];
    priorbox_kernel<<<ceil_div(cells, 128), 128, 0, ST(stream)>>>(L, reinterpret_cast<float4 *>(out));
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

#define F4(p) reinterpret_cast<const float4 *>(p)
#define F4W(p) reinterpret_cast<float4 *>(p)

extern "C" int gssd_point_form(const float *boxes, int n, float *out, void *stream) {
    if (n < 0 || (n && (!boxes || !out))) return GSSD_ERR_ARG;
    if (!n) return GSSD_OK;
    unary_box_kernel<OP_POINT_FORM><<<ceil_div(n, 256), 256, 0, ST(stream)>>>(F4(boxes), n, F4W(out));
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_center_size(const float *boxes, int n, float *out, void *stream) {
    if (n < 0 || (n && (!boxes || !out))) return GSSD_ERR_ARG;
    if (!n) return GSSD_OK;
    unary_box_kernel<OP_CENTER_SIZE><<<ceil_div(n, 256), 256, 0, ST(stream)>>>(F4(boxes), n, F4W(out));
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

template <bool IOU>
static int pair_launch(const float *a, int A, const float *b, int Bn, float *out, void *stream) {
    if (A < 0 || Bn < 0 || ((A && Bn) && (!a || !b || !out))) return GSSD_ERR_ARG;
    if (!A || !Bn) return GSSD_OK;
    if (A > 65535) return GSSD_ERR_LIMIT;
    pair_kernel<IOU><<<dim3(ceil_div(Bn, 256), A), 256, 0, ST(stream)>>>(F4(a), A, F4(b), Bn, out);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_intersect(const float *a, int A, const float *b, int Bn, float *out, void *stream) {
    return pair_launch<false>(a, A, b, Bn, out, stream);
}
extern "C" int gssd_jaccard(const float *a, int A, const float *b, int Bn, float *out, void *stream) {
    return pair_launch<true>(a, A, b, Bn, out, stream);
}

extern "C" int gssd_encode(const float *matched, const float *priors, int n, float v0, float v1, float *out, void *stream) {
    if (n < 0 || (n && (!matched || !priors || !out))) return GSSD_ERR_ARG;
    if (!n) return GSSD_OK;
    codec_kernel<true><<<ceil_div(n, 256), 256, 0, ST(stream)>>>(F4(matched), F4(priors), n, v0, v1, F4W(out));
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_decode(const float *loc, const float *priors, int n, float v0, float v1, float *out, void *stream) {
    if (n < 0 || (n && (!loc || !priors || !out))) return GSSD_ERR_ARG;
    if (!n) return GSSD_OK;
    codec_kernel<false><<<ceil_div(n, 256), 256, 0, ST(stream)>>>(F4(loc), F4(priors), n, v0, v1, F4W(out));
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_log_sum_exp(const float *x, int rows, int C, float *out, void *ws, size_t ws_bytes, void *stream) {
    if (rows <= 0 || C <= 0 || !x || !out || !ws) return GSSD_ERR_ARG;
    if (ws_bytes < sizeof(uint32_t)) return GSSD_ERR_WS;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(ws, 0, sizeof(uint32_t), ST(stream)));
    size_t n = (size_t)rows * C;
    int blocks = (int)((n + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    tensor_max_kernel<<<blocks, 256, 0, ST(stream)>>>(x, n, reinterpret_cast<uint32_t *>(ws));
    GSSD_AFTER_LAUNCH();
    lse_kernel<<<ceil_div(rows, 256), 256, 0, ST(stream)>>>(x, rows, C, reinterpret_cast<const uint32_t *>(ws), out);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_l2norm_fwd(const float *x, const float *weight, int B, int Cn, int HW, float eps,
                               float *y, float *norm, void *stream) {
    if (!x || !weight || !y || B <= 0 || Cn <= 0 || HW <= 0) return GSSD_ERR_ARG;
    if (B > 65535) return GSSD_ERR_LIMIT;
    l2norm_fwd_kernel<<<dim3(ceil_div(HW, L2_PX), B), L2_PX * L2_CG, 0, ST(stream)>>>(x, weight, Cn, HW, eps, y, norm);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" size_t gssd_l2norm_bwd_ws_bytes(int B, int Cn, int HW) {
    return (size_t)B * ceil_div(HW, L2_PX) * Cn * sizeof(float);
}

extern "C" int gssd_l2norm_bwd(const float *x, const float *weight, const float *norm, const float *gy,
                               int B, int Cn, int HW, float eps, float *gx, float *gw,
                               void *ws, size_t ws_bytes, void *stream) {
    if (!x || !weight || !norm || !gy || !gx || !gw || !ws || B <= 0 || Cn <= 0 || HW <= 0) return GSSD_ERR_ARG;
    if (B > 65535) return GSSD_ERR_LIMIT;
    if (ws_bytes < gssd_l2norm_bwd_ws_bytes(B, Cn, HW)) return GSSD_ERR_WS;
    dim3 grid(ceil_div(HW, L2_PX), B);
    l2norm_bwd_kernel<<<grid, L2_PX * L2_CG, 0, ST(stream)>>>(x, weight, norm, gy, Cn, HW, eps, gx, reinterpret_cast<float *>(ws));
    GSSD_AFTER_LAUNCH();
    l2norm_gw_reduce_kernel<<<ceil_div(Cn, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const float *>(ws), grid.x * grid.y, Cn, gw);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}
