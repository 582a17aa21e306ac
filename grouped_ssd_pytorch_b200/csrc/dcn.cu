// dcn.cu — modulated deformable convolution (DCNv2), the operator GSSD++ calls through layers/dcn_v2_custom.py:49-55 and
// 84-88 (`dcn_v2._DCNv2.apply`, CharlesShang/DCNv2: not vendored in the reference, no version pinned).
// The published algorithm — deformable im2col (bilinear samples at p + tap + offset, times the modulation mask) followed by
// a GEMM with the [c_out, c_in*9] filter matrix — is restated here B200-first:
//
//   forward   gssd_dcn_columns      x (PM bf16) + offset + mask -> columns, bf16 [rows, 9*c_in], k = tap*c_in + c
//             gssd_conv_igemm       columns as a 1x1 convolution with 9*c_in input channels: the tcgen05/TMEM kernel of gconv.cu
//   backward  gssd_conv_igemm       d_columns = dY * W          (1x1, c_out -> 9*c_in, transposed filter matrix)
//             gssd_conv_wgrad       dW = dY^T * columns         (split-K tcgen05 kernel of gconv_bwd.cu)
//             gssd_dcn_columns_bwd  d_columns -> d_input (fp32 vector reductions into a PM buffer), d_offset, d_mask
//
// The columns are HBM traffic the fused alternative (gather straight into the A operand's shared-memory tiles) would avoid, but
// they are small against the GEMM: 9*c_in*2 bytes per pixel written once and read once, 18 KB at c_in = 1024, for
// 2*9*c_in*c_out = 9.4 MFLOP per pixel at c_out = 512 — 520 flop/byte, far on the tensor side of the ridge.
//
// Sampling rule (parity is pinned against torchvision.ops.deform_conv2d, the stand-in SURVEY App. A names; it agrees with DCNv2's
// dmcn_im2col_bilinear): a sample at (y, x) is 0 unless -1 < y < H and -1 < x < W; inside, the four neighbours floor/floor+1
// contribute (1-ly)(1-lx), (1-ly)lx, ly(1-lx), ly*lx, neighbours outside the image count as 0.  PM's zero border IS that rule
// for every in-range sample, so the gather needs one predicate.  The offset gradient follows torchvision's
// get_coordinate_weight (neighbours are validated one by one, no in-range gate); the two differ from each other only for
// samples at exactly y = -1 or x = -1.
#include <cuda_bf16.h>

#include "common.cuh"

namespace gssd {

__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f[2 * j] = __uint_as_float(w[j] << 16);                    // bf16 -> fp32 is a shift
        f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&b);
}
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct DcnSample {
    float w1, w2, w3, w4;      // bilinear weights of (y0,x0) (y0,x0+1) (y0+1,x0) (y0+1,x0+1)
    float ly, lx;
    int y0, x0;
    bool in;                   // -1 < y < H and -1 < x < W
    bool near;                 // every neighbour index fits in an int and lies within two pixels of the image
};
__device__ __forceinline__ DcnSample dcn_sample(float sy, float sx, int h, int w) {
    DcnSample s;
    s.in = sy > -1.f && sy < (float)h && sx > -1.f && sx < (float)w;
    s.near = sy > -2.f && sy < (float)(h + 1) && sx > -2.f && sx < (float)(w + 1);
    const float fy = floorf(sy), fx = floorf(sx);
    s.y0 = s.near ? (int)fy : 0;
    s.x0 = s.near ? (int)fx : 0;
    s.ly = sy - fy; s.lx = sx - fx;
    const float hy = 1.f - s.ly, hx = 1.f - s.lx;
    s.w1 = hy * hx; s.w2 = hy * s.lx; s.w3 = s.ly * hx; s.w4 = s.ly * s.lx;
    return s;
}

// ---- per-pixel sampling records ---------------------------------------------------------------------------------
// Everything about a (tap, deformable group) sample is the same for all lanes of the warp that owns the pixel, so it is computed
// ONCE, one sample per lane, and kept in shared memory; the channel loop then reads it by broadcast (the first versions
// evaluated the sample in every lane at every step: 186 warp instructions per 512 bytes of columns, issue-bound).
constexpr int DCN_WARPS = 4;
constexpr int DCN_REC = 10;    // words per record: w1..w4, ly, lx, mask, r00 (int), flags (int), pad
constexpr int DCN_MAX_DG = 16;
enum : unsigned {
    DCN_IN = 1u,               // -1 < y < H and -1 < x < W
    DCN_PIX0 = 2u,             // bits 1-4: neighbour k is an image pixel (receives a gradient)
    DCN_BUF0 = 32u,            // bits 5-8: neighbour k exists in the PM buffer (its zero border stands for "outside the image")
};

__device__ __forceinline__ void dcn_make_records(float *rec, float *om, const float *__restrict__ offset, const float *__restrict__ mask,
                                                 int n, int y, int xx, int h, int w, int dg, int lane) {
    const size_t hw = (size_t)h * w, pix = (size_t)y * w + xx;
    const int nom = 27 * dg;
    __syncwarp();
    for (int i = lane; i < nom; i += 32)
        om[i] = i < 18 * dg ? __ldg(offset + ((size_t)n * 18 * dg + i) * hw + pix) : __ldg(mask + ((size_t)n * 9 * dg + (i - 18 * dg)) * hw + pix);
    __syncwarp();
    for (int i = lane; i < 9 * dg; i += 32) {                    // i = tap * dg + g: the order of the channel loop
        const int tap = i / dg, g = i - tap * dg, ky = tap / 3, kx = tap - 3 * ky;
        const float oy = om[(g * 9 + tap) * 2], ox = om[(g * 9 + tap) * 2 + 1];
        const DcnSample s = dcn_sample((float)(y - 1 + ky) + oy, (float)(xx - 1 + kx) + ox, h, w);
        const bool yl = s.near && s.y0 >= -1 && s.y0 <= h, yh = s.near && s.y0 + 1 <= h;      // near: y0 >= -2, so y0 + 1 >= -1
        const bool xl = s.near && s.x0 >= -1 && s.x0 <= w, xh = s.near && s.x0 + 1 <= w;
        const bool iy0 = s.y0 >= 0 && s.y0 < h, iy1 = s.y0 + 1 >= 0 && s.y0 + 1 < h;
        const bool ix0 = s.x0 >= 0 && s.x0 < w, ix1 = s.x0 + 1 >= 0 && s.x0 + 1 < w;
        const unsigned flags = (s.in ? DCN_IN : 0u) | (iy0 && ix0 ? DCN_PIX0 : 0u) | (iy0 && ix1 ? DCN_PIX0 << 1 : 0u) |
                               (iy1 && ix0 ? DCN_PIX0 << 2 : 0u) | (iy1 && ix1 ? DCN_PIX0 << 3 : 0u) | (yl && xl ? DCN_BUF0 : 0u) |
                               (yl && xh ? DCN_BUF0 << 1 : 0u) | (yh && xl ? DCN_BUF0 << 2 : 0u) | (yh && xh ? DCN_BUF0 << 3 : 0u);
        float *r = rec + i * DCN_REC;
        r[0] = s.w1; r[1] = s.w2; r[2] = s.w3; r[3] = s.w4; r[4] = s.ly; r[5] = s.lx;
        r[6] = om[18 * dg + g * 9 + tap];
        r[7] = __int_as_float((int)(((long)n * (h + 2) + s.y0 + 1) * (w + 2) + s.x0 + 1));       // PM row of neighbour (y0, x0); may lie two rows outside
        r[8] = __uint_as_float(flags);
    }
    __syncwarp();
}

// ---- forward: deformable im2col ---------------------------------------------------------------------------------
// One warp per padded pixel row, the warps of a CTA on consecutive pixels: the 9 taps x 4 neighbours of a pixel and of its
// neighbours along the row overlap, so part of the gathers are L1 hits.  The loop over (tap, deformable group, 256-channel pass)
// is software-pipelined: the four 16-byte gathers of step t+1 are in flight while step t is blended, converted and stored.
// 8 channels per lane, 16-byte streaming stores into the column row.  Border rows of the columns are written as zeros (they meet
// the zero border of dY in the weight-gradient GEMM).
struct DcnStep {
    uint4 v[4];                // the four neighbours' 8 channels (zeros when the sample is out of range)
    float a[4];                // mask * bilinear weight
    int dst;                   // element offset inside the column row, < 0: this lane has no channels in this pass
};

__device__ __forceinline__ void dcn_fetch(DcnStep &st, int t, int npass, int dg, int cpg, int c_in, int wp, int lane, const float *rec,
                                          const __nv_bfloat16 *__restrict__ x) {
    const int pass = t % npass, i = t / npass, tap = i / dg, g = i - tap * dg;
    const int c = pass * 256 + lane * 8;
    const float *r = rec + i * DCN_REC;
    const float m = r[6];
    st.a[0] = m * r[0]; st.a[1] = m * r[1]; st.a[2] = m * r[2]; st.a[3] = m * r[3];
    const bool act = c < cpg;
    st.dst = act ? tap * c_in + g * cpg + c : -1;
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (act && (__float_as_uint(r[8]) & DCN_IN)) {
        const __nv_bfloat16 *p = x + (long)__float_as_int(r[7]) * c_in + g * cpg + c;
        st.v[0] = __ldg(reinterpret_cast<const uint4 *>(p));
        st.v[1] = __ldg(reinterpret_cast<const uint4 *>(p + c_in));
        st.v[2] = __ldg(reinterpret_cast<const uint4 *>(p + (size_t)wp * c_in));
        st.v[3] = __ldg(reinterpret_cast<const uint4 *>(p + (size_t)(wp + 1) * c_in));
    } else {
        st.v[0] = z; st.v[1] = z; st.v[2] = z; st.v[3] = z;
    }
}

__global__ void __launch_bounds__(DCN_WARPS * 32) dcn_columns_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ offset,
                                                                     const float *__restrict__ mask, int c_in, int h, int w, int dg,
                                                                     __nv_bfloat16 *__restrict__ col, long rows) {
    extern __shared__ float s_dcn[];                                         // [warps][27*dg + 9*dg*DCN_REC]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hp = h + 2, wp = w + 2, cpg = c_in / dg;
    const int npass = (cpg + 255) / 256, T = 9 * dg * npass;
    float *om = s_dcn + warp * (27 + 9 * DCN_REC) * dg, *rec = om + 27 * dg;
    for (long row = (long)blockIdx.x * DCN_WARPS + warp; row < rows; row += (long)gridDim.x * DCN_WARPS) {
        const int n = (int)(row / (hp * wp)), rem = (int)(row - (long)n * hp * wp), py = rem / wp, px = rem - py * wp;
        __nv_bfloat16 *dst = col + (size_t)row * 9 * c_in;
        if (py < 1 || py > h || px < 1 || px > w) {
            for (int c = lane * 8; c < 9 * c_in; c += 256) __stcs(reinterpret_cast<uint4 *>(dst + c), make_uint4(0, 0, 0, 0));
            continue;
        }
        dcn_make_records(rec, om, offset, mask, n, py - 1, px - 1, h, w, dg, lane);
        DcnStep nxt;
        dcn_fetch(nxt, 0, npass, dg, cpg, c_in, wp, lane, rec, x);
        for (int t = 0; t < T; ++t) {
            const DcnStep cur = nxt;
            if (t + 1 < T) dcn_fetch(nxt, t + 1, npass, dg, cpg, c_in, wp, lane, rec, x);
            float r[8], v[8];
            unpack8(cur.v[0], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = cur.a[0] * v[j];
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                unpack8(cur.v[k], v);
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] += cur.a[k] * v[j];
            }
            if (cur.dst >= 0)
                __stcs(reinterpret_cast<uint4 *>(dst + cur.dst), make_uint4(pack2(r[0], r[1]), pack2(r[2], r[3]), pack2(r[4], r[5]), pack2(r[6], r[7])));
        }
    }
}

// ---- backward of the columns -----------------------------------------------------------------------------------------
// One warp per interior pixel, the same pipelined loop with the step's d_columns loaded beside the gathers.  d_input goes
// into a PM-shaped fp32 buffer by 16-byte vector reductions (channels innermost: a warp adds 1 KB runs) — the L1 / L2
// reduction path retires about four fp32 elements per cycle and SM, which bounds this kernel (ncu: L1/TEX 69 % busy, 8 RED.128
// per lane and step).  d_offset / d_mask are warp sums over the channels of a deformable group: the three sums are reduced
// together in 6 shuffles (after two exchange steps each quarter of the warp holds one of them) instead of 3 x 5.  The offset
// gradient takes the neighbours that exist one by one (torchvision's get_coordinate_weight), without the in-range gate.
struct DcnStepB {
    uint4 v[4], d;             // neighbours, d_columns
    const float *rec;          // the sample's record
    long q;                    // element offset of neighbour (y0, x0) in dx, group and channel included
    int i;                     // tap * dg + g
    bool act, last;            // lane has channels in this pass; last pass of this (tap, group)
};

__device__ __forceinline__ void dcn_fetch_b(DcnStepB &st, int t, int npass, int dg, int cpg, int c_in, int wp, int lane, const float *rec,
                                            const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ dsrc) {
    const int pass = t % npass, i = t / npass, tap = i / dg, g = i - tap * dg;
    const int c = pass * 256 + lane * 8;
    const float *r = rec + i * DCN_REC;
    st.rec = r; st.i = i; st.act = c < cpg; st.last = pass == npass - 1;
    const unsigned flags = __float_as_uint(r[8]);
    st.q = (long)__float_as_int(r[7]) * c_in + g * cpg + c;
    const __nv_bfloat16 *p = x + st.q;
    const uint4 z = make_uint4(0, 0, 0, 0);
    st.d = st.act ? __ldcs(reinterpret_cast<const uint4 *>(dsrc + (size_t)tap * c_in + g * cpg + c)) : z;
    st.v[0] = st.act && (flags & DCN_BUF0) ? __ldg(reinterpret_cast<const uint4 *>(p)) : z;
    st.v[1] = st.act && (flags & (DCN_BUF0 << 1)) ? __ldg(reinterpret_cast<const uint4 *>(p + c_in)) : z;
    st.v[2] = st.act && (flags & (DCN_BUF0 << 2)) ? __ldg(reinterpret_cast<const uint4 *>(p + (size_t)wp * c_in)) : z;
    st.v[3] = st.act && (flags & (DCN_BUF0 << 3)) ? __ldg(reinterpret_cast<const uint4 *>(p + (size_t)(wp + 1) * c_in)) : z;
}

__global__ void __launch_bounds__(DCN_WARPS * 32) dcn_columns_bwd_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ offset,
                                                                         const float *__restrict__ mask, const __nv_bfloat16 *__restrict__ dcol,
                                                                         int c_in, int h, int w, int dg, float *__restrict__ dx,
                                                                         float *__restrict__ d_offset, float *__restrict__ d_mask, long pixels) {
    extern __shared__ float s_dcn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hp = h + 2, wp = w + 2, cpg = c_in / dg;
    const int npass = (cpg + 255) / 256, T = 9 * dg * npass;
    const size_t hw = (size_t)h * w;
    float *om = s_dcn + warp * (27 + 9 * DCN_REC) * dg, *rec = om + 27 * dg;
    for (long pixel = (long)blockIdx.x * DCN_WARPS + warp; pixel < pixels; pixel += (long)gridDim.x * DCN_WARPS) {
        const int n = (int)(pixel / (long)hw), rem = (int)(pixel - (long)n * hw), y = rem / w, xx = rem - y * w;
        const size_t pix = (size_t)y * w + xx;
        const __nv_bfloat16 *dsrc = dcol + (((size_t)n * hp + y + 1) * wp + xx + 1) * 9 * c_in;
        dcn_make_records(rec, om, offset, mask, n, y, xx, h, w, dg, lane);
        DcnStepB nxt;
        dcn_fetch_b(nxt, 0, npass, dg, cpg, c_in, wp, lane, rec, x, dsrc);
        float sval = 0.f, sdy = 0.f, sdx = 0.f;
        for (int t = 0; t < T; ++t) {
            const DcnStepB cur = nxt;
            if (t + 1 < T) dcn_fetch_b(nxt, t + 1, npass, dg, cpg, c_in, wp, lane, rec, x, dsrc);
            const float w1 = cur.rec[0], w2 = cur.rec[1], w3 = cur.rec[2], w4 = cur.rec[3], ly = cur.rec[4], lx = cur.rec[5], m = cur.rec[6];
            const unsigned flags = __float_as_uint(cur.rec[8]);
            float d[8], v1[8], v2[8], v3[8], v4[8];
            unpack8(cur.d, d); unpack8(cur.v[0], v1); unpack8(cur.v[1], v2); unpack8(cur.v[2], v3); unpack8(cur.v[3], v4);
            const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sdy += d[j] * (lx * (v4[j] - v2[j]) + hx * (v3[j] - v1[j]));
                sdx += d[j] * (ly * (v4[j] - v3[j]) + hy * (v2[j] - v1[j]));
            }
            if (cur.act && (flags & DCN_IN)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) sval += d[j] * (w1 * v1[j] + w2 * v2[j] + w3 * v3[j] + w4 * v4[j]);
                float *q = dx + cur.q;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((flags & (DCN_PIX0 << k)) && cur.rec[k] != 0.f) {      // (samples on the pixel grid — the module's zero-initialised
                                                                                // offsets — touch one neighbour, not four)
                        float *tq = q + (size_t)((k >> 1) * wp + (k & 1)) * c_in;
                        const float a = m * cur.rec[k];
                        red_add_v4(tq, a * d[0], a * d[1], a * d[2], a * d[3]);
                        red_add_v4(tq + 4, a * d[4], a * d[5], a * d[6], a * d[7]);
                    }
                }
            }
            if (cur.last) {                                      // last pass of this (tap, group): reduce the three sums together
                const bool hi = lane & 16;
                float a = hi ? sdx : sval, b = hi ? 0.f : sdy;
                a += __shfl_xor_sync(FULL, hi ? sval : sdx, 16);
                b += __shfl_xor_sync(FULL, hi ? sdy : 0.f, 16);
                const bool q8 = lane & 8;
                float k = q8 ? b : a;
                k += __shfl_xor_sync(FULL, q8 ? a : b, 8);
                k += __shfl_xor_sync(FULL, k, 4);
                k += __shfl_xor_sync(FULL, k, 2);
                k += __shfl_xor_sync(FULL, k, 1);
                // lane 0: sum of sval, lane 8: sdy, lane 16: sdx
                const int tap = cur.i / dg, g = cur.i - tap * dg;
                const size_t ob = ((size_t)(n * dg + g) * 18 + 2 * tap) * hw + pix;
                if (lane == 0) d_mask[((size_t)(n * dg + g) * 9 + tap) * hw + pix] = k;
                if (lane == 8) d_offset[ob] = m * k;
                if (lane == 16) d_offset[ob + hw] = m * k;
                sval = sdy = sdx = 0.f;
            }
        }
    }
}

static int dcn_check(int n_img, int c_in, int h, int w, int dg) {
    if (n_img <= 0 || c_in <= 0 || h <= 0 || w <= 0 || dg <= 0) return GSSD_ERR_ARG;
    if (c_in % dg) return GSSD_ERR_ARG;
    if ((c_in / dg) % 8 || w > 512 || dg > DCN_MAX_DG) return GSSD_ERR_LIMIT;
    if ((long)n_img * (h + 2) * (w + 2) > (1l << 30) / 9) return GSSD_ERR_LIMIT;
    return GSSD_OK;
}
static int dcn_grid(long warps) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long want = (warps + DCN_WARPS - 1) / DCN_WARPS;
    return (int)(want < (long)sms * 32 ? want : (long)sms * 32);
}

}  // namespace gssd

using namespace gssd;

extern "C" int gssd_dcn_columns(const void *x_bf16, const float *offset, const float *mask, int n_img, int c_in, int height,
                                int width, int deformable_groups, void *col_bf16, void *stream) {
    if (!x_bf16 || !offset || !mask || !col_bf16) return GSSD_ERR_ARG;
    int rc = dcn_check(n_img, c_in, height, width, deformable_groups);
    if (rc) return rc;
    const long rows = (long)n_img * (height + 2) * (width + 2);
    dcn_columns_kernel<<<dcn_grid(rows), DCN_WARPS * 32, DCN_WARPS * (27 + 9 * DCN_REC) * deformable_groups * sizeof(float), (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16 *>(x_bf16), offset, mask, c_in, height, width, deformable_groups,
        reinterpret_cast<__nv_bfloat16 *>(col_bf16), rows);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_dcn_columns_bwd(const void *x_bf16, const float *offset, const float *mask, const void *dcol_bf16, int n_img,
                                    int c_in, int height, int width, int deformable_groups, float *dx_pm, float *d_offset,
                                    float *d_mask, void *stream) {
    if (!x_bf16 || !offset || !mask || !dcol_bf16 || !dx_pm || !d_offset || !d_mask) return GSSD_ERR_ARG;
    int rc = dcn_check(n_img, c_in, height, width, deformable_groups);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t rows = (size_t)n_img * (height + 2) * (width + 2);
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(dx_pm, 0, rows * c_in * sizeof(float), st));
    const long pixels = (long)n_img * height * width;
    dcn_columns_bwd_kernel<<<dcn_grid(pixels), DCN_WARPS * 32, DCN_WARPS * (27 + 9 * DCN_REC) * deformable_groups * sizeof(float), st>>>(
        reinterpret_cast<const __nv_bfloat16 *>(x_bf16), offset, mask, reinterpret_cast<const __nv_bfloat16 *>(dcol_bf16), c_in, height,
        width, deformable_groups, dx_pm, d_offset, d_mask, pixels);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}
