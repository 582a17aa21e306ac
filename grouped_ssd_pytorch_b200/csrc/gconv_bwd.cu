// gconv_bwd.cu — backward of the source block (what autograd computes through
// models/ssd_multiphase_custom_group.py:258-380 when train_lesion_multiphase_v2.py:248 calls loss.backward()).
//
//   data gradient   : a stride-1 "same" convolution of dY with the rotated, channel-swapped filter — the FORWARD kernel
//                     (gssd_conv_igemm, gconv.cu) on re-packed weights; nothing new here
//   weight gradient : gssd_conv_wgrad — a tcgen05/TMEM GEMM over the pixels,
//                         dW[tap][co][ci] = sum_row dY[row][co] * X[row + (dy-1)*(W+2) + (dx-1)][ci],
//                     both operands MN-major straight from SWIZZLE_128B TMA boxes of the pixel-major tensors (channels are
//                     contiguous in memory, which is the M / N direction of this GEMM), split over the pixels (split-K),
//                     fp32 tiles added into dW by TMA reductions
//   BN / ReLU / L2Norm backward, head-gradient gather: bandwidth-bound row kernels on the PM layout
#include "common.cuh"
#include "tc.cuh"
#include "tmap.cuh"

namespace gssd {

// ---- weight gradient --------------------------------------------------------------------------------------------------------
constexpr int WG_KS = 32;                        // pixel rows per pipeline stage (two K = 16 instructions per tile)
constexpr int WG_A_BOX = WG_KS * 128;            // one [32 rows x 64 channels] box of dY
constexpr int WG_B_BOX1 = WG_KS * 128;           // 1x1: one [32 x 64] box of X
constexpr int WG_B_BOX9 = 5 * 1024;              // 3x3: one [34 x 64] box of X (rows for the three dx taps), whole swizzle atoms
constexpr int WG_STG_BYTES = 128 * 32 * 4;       // epilogue staging: [128 co][32 ci] fp32
constexpr int WG_MAX_STAGES = 8;

struct WgradParams {
    int rows, wp;                   // rows of the PM tensors, padded width
    int taps;                       // 1 or 9
    int ng, cg;                     // output / input channels per group
    int m_tiles, n_units, nsub;     // per group: 128-row co tiles, units of `nsub` 128-column ci tiles (3x3: nsub = 3 dx taps of one ci tile)
    int tap_rows;                   // 3 (3x3: one unit per filter row) or 1
    int c_out_pad;                  // rows per tap in dW
    int stages, stage_bytes, b_box; // pipeline depth, bytes per stage, bytes per B box
    int stages_per_chunk, n_chunks;
};

// MN-major SWIZZLE_128B operand: 8-row (here: 8-pixel) atoms of 64 channels = 1024 bytes, SBO = 1024 to the next 8 pixels
// (the K direction), LBO = one box to the next 64 channels (the M / N direction); confirmed on hardware by
// tools/umma_mnmajor_probe.cu, row offsets inside the box included
__device__ __forceinline__ uint64_t wg_desc_mn(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)(1024u >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void wg_tma_reduce_add_2d(const CUtensorMap *map, const void *smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}

// unit = (group, co tile, ci unit, filter row, pixel chunk); 256 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocation, warps 4-7 epilogue
__global__ void __launch_bounds__(256, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
             const __grid_constant__ CUtensorMap map_dw, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *stg = smem + p.stages * p.stage_bytes;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(stg + WG_STG_BYTES);
    uint64_t *bar_empty = bar_full + WG_MAX_STAGES;
    uint64_t *bar_done = bar_empty + WG_MAX_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int u = blockIdx.x;
    const int chunk = u % p.n_chunks; u /= p.n_chunks;
    const int tr = u % p.tap_rows; u /= p.tap_rows;
    const int nu = u % p.n_units; u /= p.n_units;
    const int mt = u % p.m_tiles;
    const int g = u / p.m_tiles;
    const int s0 = chunk * p.stages_per_chunk;
    const int s1 = min(s0 + p.stages_per_chunk, (p.rows + WG_KS - 1) / WG_KS);
    const int n_it = max(s1 - s0, 0);
    const int co0 = g * p.ng + mt * 128;                                   // first output channel of the tile (global)
    const bool conv3 = p.taps == 9;
    const int ci_unit0 = conv3 ? nu * 128 : nu * p.nsub * 128;           // first input channel (within the group) of the unit
    const int nsub = conv3 ? 3 : min(p.nsub, (p.cg - ci_unit0) / 128);

    if (warp == 0 && lane == 0) { tc::prefetch_tensormap(&map_dy); tc::prefetch_tensormap(&map_x); tc::prefetch_tensormap(&map_dw); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < WG_MAX_STAGES; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_empty[i], 1); }
        tc::mbar_init(bar_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int b_tiles = conv3 ? 1 : nsub;                                  // 128-channel X tiles loaded per stage

    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool leader = tc::elect_one();
        const uint32_t tx = 2 * WG_A_BOX + (uint32_t)b_tiles * 2u * (conv3 ? (WG_KS + 2) * 128 : WG_B_BOX1);
        for (int it = 0; it < n_it; ++it) {
            const int s = it % p.stages, ph = (it / p.stages) & 1;
            tc::mbar_wait(&bar_empty[s], ph ^ 1);
            if (leader) {
                uint8_t *a = smem + s * p.stage_bytes, *b = a + 2 * WG_A_BOX;
                const int row = (s0 + it) * WG_KS;
                const int xrow = conv3 ? row + (tr - 1) * p.wp - 1 : row;  // signed: rows outside the tensor arrive as zeros
                tc::mbar_arrive_expect_tx(&bar_full[s], tx);
                tc::tma_load_2d(a, &map_dy, &bar_full[s], co0, row);
                tc::tma_load_2d(a + WG_A_BOX, &map_dy, &bar_full[s], co0 + 64, row);
                for (int t = 0; t < b_tiles; ++t) {
                    const int ci = g * p.cg + ci_unit0 + t * 128;
                    tc::tma_load_2d(b + (2 * t) * p.b_box, &map_x, &bar_full[s], ci, xrow);
                    tc::tma_load_2d(b + (2 * t + 1) * p.b_box, &map_x, &bar_full[s], ci + 64, xrow);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = tc::elect_one();
        // kind::f16, bf16 x bf16 -> fp32, M = N = 128, both operands MN-major (bits 15 / 16)
        const uint32_t idesc = tc::idesc_bf16_f32(128, 128) | (1u << 15) | (1u << 16);
        for (int it = 0; it < n_it; ++it) {
            const int s = it % p.stages, ph = (it / p.stages) & 1;
            tc::mbar_wait(&bar_full[s], ph);
            tc::fence_after_thread_sync();
            if (leader) {
                const uint32_t a = tc::smem_u32(smem + s * p.stage_bytes), b = a + 2 * WG_A_BOX;
                for (int t = 0; t < nsub; ++t) {
                    // 3x3: tile t = tap dx = t of the one X slab (start address + t pixels); 1x1: tile t = ci tile t
                    const uint32_t bt = conv3 ? b + (uint32_t)t * 128u : b + (uint32_t)(2 * t) * (uint32_t)p.b_box;
#pragma unroll
                    for (int ks = 0; ks < WG_KS / 16; ++ks)
                        tc::umma_bf16(tmem + t * 128, wg_desc_mn(a + ks * 2048, WG_A_BOX), wg_desc_mn(bt + ks * 2048, (uint32_t)p.b_box),
                                      idesc, (it > 0 || ks > 0) ? 1u : 0u);
                }
                tc::umma_commit(&bar_empty[s]);                            // the stage is free once these MMAs have read it
                if (it == n_it - 1) tc::umma_commit(bar_done);
            }
            __syncwarp();
        }
    } else if (warp >= 4 && n_it > 0) {
        // ===================== epilogue: TMEM -> shared -> TMA reduce-add into dW =====================
        const int ew = warp - 4;                                           // TMEM lanes 32*ew .. 32*ew + 31 = co
        tc::mbar_wait(bar_done, 0);
        tc::fence_after_thread_sync();
        float *row = reinterpret_cast<float *>(stg) + (ew * 32 + lane) * 32;
        for (int t = 0; t < nsub; ++t) {
            const int tap = conv3 ? tr * 3 + t : 0;
            const int ci0 = conv3 ? ci_unit0 : ci_unit0 + t * 128;
            for (int c = 0; c < 128; c += 32) {
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem + ((uint32_t)(ew * 32) << 16) + t * 128 + c, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(row + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");             // the four epilogue warps
                if (warp == 4 && lane == 0) {
                    wg_tma_reduce_add_2d(&map_dw, stg, ci0 + c, tap * p.c_out_pad + co0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging may be overwritten
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        if (warp == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 2) { tc::fence_after_thread_sync(); tc::tmem_dealloc(tmem, 512); }
}

// ---- head gradient gather ---------------------------------------------------------------------------------------------------
// (d_loc[B,P,4], d_conf[B,P,C]) of one source -> dH PM bf16 [rows, c_pad] (channels [loc 4A | conf A*C | zeros]) + per-channel
// sums (the bias gradients).  One warp per pixel row.
__global__ void __launch_bounds__(256) head_grad_pm_kernel(const float *__restrict__ d_loc, const float *__restrict__ d_conf, int n_priors,
                                                           int prior_off, int n_anchor, int n_cls, int rows, int hp, int wp, int h, int w,
                                                           int c_pad, __nv_bfloat16 *__restrict__ out, float *__restrict__ bias_grad) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int n_loc = 4 * n_anchor, n_head = n_loc + n_anchor * n_cls;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};                                      // channels lane, lane + 32, ... (c_pad <= 128)
    for (int m = blockIdx.x * wpb + warp; m < rows; m += gridDim.x * wpb) {
        const int img = m / (hp * wp), rem = m - img * (hp * wp), py = rem / wp, px = rem - py * wp;
        const bool interior = py >= 1 && py <= h && px >= 1 && px <= w;
        const size_t prior = (size_t)img * n_priors + prior_off + (size_t)((py - 1) * w + (px - 1)) * n_anchor;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            if (c >= c_pad) break;
            float v = 0.f;
            if (interior && c < n_head) v = c < n_loc ? d_loc[prior * 4 + c] : d_conf[prior * n_cls + (c - n_loc)];
            const __nv_bfloat16 bv = __float2bfloat16_rn(v);
            out[(size_t)m * c_pad + c] = bv;
            acc[q] += v;
        }
    }
    if (bias_grad != nullptr) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            if (c < n_head && acc[q] != 0.f) atomicAdd(bias_grad + c, acc[q]);
        }
    }
}

// ---- BatchNorm / ReLU / L2Norm backward on PM tensors --------------------------------------------------------------------------------
// The gradient that reaches the ReLU output of one stage of the chain is, per pixel row,
//     g0 = dy (+ add)                                                  plain
//     g0 = (dy - y * (sum_c dy*y) * rs / sqrt(ss)) (+ add)             when the consumer applied L2Norm to y (l2norm.py:19-23, deferred
//                                                                      form: dy already carries the 1/(sqrt(ss)+eps) factor and the weight)
// then g = g0 * [y > 0] (ReLU), then BatchNorm backward with batch statistics (x_hat from the raw conv output):
//     dx = gamma * rstd * (g - mean(g) - x_hat * mean(g * x_hat)),   dgamma = sum g*x_hat,   dbeta = sum g
// or, without statistics (eval-mode BN folded, or no BN): dx = g * scale[c].   out = dx (* rs_out[row] for the producer's wgrad/dgrad
// when ITS input was L2-normalised).  One warp per pixel row, 8 channels per lane and trip.
struct BnBwdArgs {
    const __nv_bfloat16 *dy, *y, *yraw, *add;   // y: post-ReLU activations (mask, L2 term); yraw: raw conv output (x_hat) or null
    __nv_bfloat16 *out;
    int rows, c, hp, wp, h, w;
    const float *chan_sum;                      // forward statistics (sum, sum of squares) or null
    const float *gamma;                         // BN weight (null = 1) — or the folded scale when chan_sum == null
    const float *ebn_w, *ebn_b;                 // eval-mode BN (folded in the forward): its weight / bias, for THEIR gradients:
                                                // x_hat = (y - beta)/gamma wherever y > 0 (elsewhere g = 0)
    float bn_eps, inv_count;
    const float *row_ss_l2; float l2_eps;       // L2Norm on y by the consumer, or null
    const float *row_ss_out; float l2_eps_out;  // scale the output rows by 1/(sqrt(ss)+eps) (the producer's input was L2-normalised), or null
    float *sums;                                // [3*c]: sum g, sum g*x_hat, sum dx   (reduce pass writes 0..2c, apply pass adds into 2c..3c)
    int relu;
};

// SLOTS = c / 256 groups of 8 channels per lane: the accumulators are sized for it, so that the 256- and 512-channel layers of the
// backbone keep several CTAs per SM (sized for 1024 channels the APPLY pass needs 254 registers: one CTA of 8 warps per SM)
template <bool APPLY, int SLOTS>
__global__ void __launch_bounds__(256, SLOTS == 1 ? 3 : (SLOTS == 2 ? 2 : 1)) bn_bwd_pm_kernel(BnBwdArgs a) {
    // one float4 of coefficients per channel, read with ONE shared-memory load per element (separate mean / rstd / mean(g) /
    // mean(g*x_hat) arrays plus gamma from global memory made the row loop issue-bound: 400 warp instructions per 256-channel row,
    // 198 us for 32 x 75 x 75 x 256, ncu):   reduce pass (mean, rstd, -, -)      x_hat = (raw - mean) * rstd
    //                                         apply pass  (mean, p, mg, k)        dx = k * (g - mg - (raw - mean) * p),
    //                                         p = rstd * mean(g*x_hat), mg = mean(g), k = gamma * rstd   (no statistics: (0, 0, 0, gamma))
    // afterwards the same memory holds the partial channel sums
    extern __shared__ __align__(16) float sm[];
    const int c = a.c;
    float4 *s_coef = reinterpret_cast<float4 *>(sm);
    const bool stats = a.chan_sum != nullptr;
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        float mean = 0.f, rstd = 1.f;
        if (stats) {
            mean = a.chan_sum[i] * a.inv_count;
            const float var = fmaxf(a.chan_sum[c + i] * a.inv_count - mean * mean, 0.f);
            rstd = rsqrtf(var + a.bn_eps);
        }
        // channel i = (slot * 32 + lane) * 8 + k is kept at [(slot * 8 + k) * 32 + lane]: the 32 lanes of a warp, which read the same
        // (slot, k) together, hit 32 consecutive float4 (in channel order they are 128 bytes apart: every lane on the same four banks)
        const int at = ((i >> 8) * 8 + (i & 7)) * 32 + ((i >> 3) & 31);
        if (!APPLY) {
            s_coef[at] = make_float4(mean, rstd, 0.f, 0.f);
        } else {
            const float gam = a.gamma ? a.gamma[i] : 1.f;
            s_coef[at] = stats ? make_float4(mean, rstd * (a.sums[c + i] * a.inv_count), a.sums[i] * a.inv_count, gam * rstd)
                               : make_float4(0.f, 0.f, 0.f, gam);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    // per-lane channel-sum accumulators live in shared memory after the row loop (c can be 1024: 32 per lane)
    const bool ebn = APPLY && !stats && a.ebn_w != nullptr;
    float acc0[8 * SLOTS], acc1[8 * SLOTS], acc2[8 * SLOTS];
#pragma unroll
    for (int j = 0; j < 8 * SLOTS; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; acc2[j] = 0.f; }
    for (int m = blockIdx.x * wpb + warp; m < a.rows; m += gridDim.x * wpb) {
        const int rem = m % (a.hp * a.wp), py = rem / a.wp, px = rem - py * a.wp;
        const bool interior = py >= 1 && py <= a.h && px >= 1 && px <= a.w;
        if (!interior) {
            if (APPLY) for (int i = lane; i < c / 8; i += 32) reinterpret_cast<uint4 *>(a.out + (size_t)m * c)[i] = make_uint4(0, 0, 0, 0);
            continue;
        }
        const uint4 *dyr = reinterpret_cast<const uint4 *>(a.dy + (size_t)m * c);
        const uint4 *yr = reinterpret_cast<const uint4 *>(a.y + (size_t)m * c);
        const uint4 *rawr = a.yraw ? reinterpret_cast<const uint4 *>(a.yraw + (size_t)m * c) : nullptr;
        const uint4 *addr = a.add ? reinterpret_cast<const uint4 *>(a.add + (size_t)m * c) : nullptr;
        float l2_coef = 0.f;
        if (a.row_ss_l2 != nullptr) {                                       // sum_c dy*y over the row
            float dot = 0.f;
            for (int i = lane; i < c / 8; i += 32) {
                const uint4 q = dyr[i], r = yr[i];
                const __nv_bfloat162 *qa = reinterpret_cast<const __nv_bfloat162 *>(&q), *ra = reinterpret_cast<const __nv_bfloat162 *>(&r);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(qa[j]), y2 = __bfloat1622float2(ra[j]);
                    dot = fmaf(f.x, y2.x, fmaf(f.y, y2.y, dot));
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
            const float n = sqrtf(a.row_ss_l2[m]);
            l2_coef = n > 0.f ? dot / ((n + a.l2_eps) * n) : 0.f;
        }
        const float rs_out = a.row_ss_out ? 1.f / (sqrtf(a.row_ss_out[m]) + a.l2_eps_out) : 1.f;
#pragma unroll
        for (int slot = 0; slot < SLOTS; ++slot) {                          // c <= 1024: at most 4 groups of 8 channels per lane
            const int i = lane + 32 * slot;
            if (i >= c / 8) break;
            const uint4 q = dyr[i], r = yr[i];
            uint4 w4 = make_uint4(0, 0, 0, 0), ad = make_uint4(0, 0, 0, 0);
            if (rawr) w4 = rawr[i];
            if (addr) ad = addr[i];
            const __nv_bfloat162 *qa = reinterpret_cast<const __nv_bfloat162 *>(&q), *ra = reinterpret_cast<const __nv_bfloat162 *>(&r);
            const __nv_bfloat162 *wa = reinterpret_cast<const __nv_bfloat162 *>(&w4), *aa = reinterpret_cast<const __nv_bfloat162 *>(&ad);
            uint4 o4;
            uint32_t *ow = reinterpret_cast<uint32_t *>(&o4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 dv = __bfloat1622float2(qa[j]), yv = __bfloat1622float2(ra[j]), rv = __bfloat1622float2(wa[j]), av = __bfloat1622float2(aa[j]);
                float res[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ch = i * 8 + 2 * j + e;
                    const float d = e ? dv.y : dv.x, yy = e ? yv.y : yv.x, raw = e ? rv.y : rv.x, ad1 = e ? av.y : av.x;
                    float g = d - yy * l2_coef + ad1;
                    if (a.relu && !(yy > 0.f)) g = 0.f;
                    const float4 cf = s_coef[(slot * 8 + 2 * j + e) * 32 + lane];
                    if (!APPLY) {
                        const float xh = stats ? (raw - cf.x) * cf.y : 0.f;
                        acc0[slot * 8 + 2 * j + e] += g;
                        acc1[slot * 8 + 2 * j + e] += g * xh;
                        res[e] = 0.f;
                    } else {
                        const float dx = cf.w * (g - cf.z - (raw - cf.x) * cf.y);
                        acc0[slot * 8 + 2 * j + e] += dx;
                        if (ebn) {
                            const float bw = a.ebn_w[ch];
                            acc1[slot * 8 + 2 * j + e] += bw != 0.f ? g * (yy - a.ebn_b[ch]) / bw : 0.f;
                            acc2[slot * 8 + 2 * j + e] += g;
                        }
                        res[e] = dx * rs_out;
                    }
                }
                ow[j] = tc::pack_bf16x2(res[0], res[1]);
            }
            if (APPLY) reinterpret_cast<uint4 *>(a.out + (size_t)m * c)[i] = o4;
        }
    }
    // channel sums: registers -> shared -> global atomics
    __syncthreads();
    float *s_a = sm, *s_b = sm + c, *s_c = sm + 2 * c;                      // reuse (the coefficients are dead)
    for (int i = threadIdx.x; i < 3 * c; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    {
#pragma unroll
        for (int slot = 0; slot < SLOTS; ++slot) {
            const int i = lane + 32 * slot;
            if (i >= c / 8) break;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&s_a[i * 8 + j], acc0[slot * 8 + j]);
                if (!APPLY || ebn) atomicAdd(&s_b[i * 8 + j], acc1[slot * 8 + j]);
                if (ebn) atomicAdd(&s_c[i * 8 + j], acc2[slot * 8 + j]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        if (!APPLY) { atomicAdd(a.sums + i, s_a[i]); atomicAdd(a.sums + c + i, s_b[i]); }
        else {
            atomicAdd(a.sums + 2 * c + i, s_a[i]);
            if (ebn) { atomicAdd(a.sums + i, s_c[i]); atomicAdd(a.sums + c + i, s_b[i]); }
        }
    }
}

// one CTA of 8 warps per 8 pixel rows, at most as many CTAs as are resident at once (the row loop strides over the rest)
template <int SLOTS>
static int launch_bn_bwd_pm(const BnBwdArgs &a, long rows, size_t smem, bool reduce_first, cudaStream_t st) {
    const long want = (rows + 7) / 8;
    if (reduce_first) {
        const int cap = resident_ctas(reinterpret_cast<const void *>(bn_bwd_pm_kernel<false, SLOTS>), 256, smem);
        bn_bwd_pm_kernel<false, SLOTS><<<(int)(want < cap ? want : cap), 256, smem, st>>>(a);
        GSSD_AFTER_LAUNCH();
    }
    const int cap = resident_ctas(reinterpret_cast<const void *>(bn_bwd_pm_kernel<true, SLOTS>), 256, smem);
    bn_bwd_pm_kernel<true, SLOTS><<<(int)(want < cap ? want : cap), 256, smem, st>>>(a);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

}  // namespace gssd

using namespace gssd;

extern "C" int gssd_conv_wgrad(const void *dy_bf16, const void *x_bf16, int n_img, int height, int width, int c_in, int c_out, int dy_channels,
                               int groups, int taps, float *dw, void *stream) {
    if (!dy_bf16 || !x_bf16 || !dw) return GSSD_ERR_ARG;
    if (n_img <= 0 || height <= 0 || width <= 0 || c_in <= 0 || c_out <= 0 || groups <= 0 || (taps != 1 && taps != 9)) return GSSD_ERR_ARG;
    if (c_in % groups || c_out % groups || dy_channels < c_out || dy_channels % 64) return GSSD_ERR_ARG;
    const int cg = c_in / groups, ng = c_out / groups;
    if (cg % 128) return GSSD_ERR_LIMIT;
    if (groups > 1 && ng % 128) return GSSD_ERR_LIMIT;                         // a co tile must not straddle two groups
    const long rows_l = (long)n_img * (height + 2) * (width + 2);
    if (rows_l > (1l << 30)) return GSSD_ERR_LIMIT;
    WgradParams p;
    p.rows = (int)rows_l; p.wp = width + 2; p.taps = taps; p.ng = ng; p.cg = cg;
    p.m_tiles = ceil_div(ng, 128);
    p.tap_rows = taps == 9 ? 3 : 1;
    p.nsub = taps == 9 ? 3 : (cg / 128 < 4 ? cg / 128 : 4);
    p.n_units = taps == 9 ? cg / 128 : ceil_div(cg / 128, p.nsub);
    p.c_out_pad = ceil_div(c_out, 128) * 128;
    p.b_box = taps == 9 ? WG_B_BOX9 : WG_B_BOX1;
    p.stage_bytes = 2 * WG_A_BOX + (taps == 9 ? 2 : 2 * p.nsub) * p.b_box;
    const int budget = 232448 - 1024 - WG_STG_BYTES - (2 * WG_MAX_STAGES + 1) * 8 - 16;
    p.stages = budget / p.stage_bytes;
    if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
    if (p.stages < 2) return GSSD_ERR_LIMIT;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int units = groups * p.m_tiles * p.n_units * p.tap_rows;
    const int total_stages = ceil_div(p.rows, WG_KS);
    p.n_chunks = sms / units > 0 ? sms / units : 1;
    if (p.n_chunks > total_stages) p.n_chunks = total_stages;
    p.stages_per_chunk = ceil_div(total_stages, p.n_chunks);
    p.n_chunks = ceil_div(total_stages, p.stages_per_chunk);
    CUtensorMap m_dy, m_x, m_dw;
    int rc = make_tmap_2d(&m_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy_bf16, (uint64_t)p.rows, (uint64_t)dy_channels, WG_KS, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_2d(&m_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x_bf16, (uint64_t)p.rows, (uint64_t)c_in, taps == 9 ? WG_KS + 2 : WG_KS, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_2d(&m_dw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dw, (uint64_t)taps * p.c_out_pad, (uint64_t)cg, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * p.c_out_pad * cg, st));
    const size_t smem = (size_t)p.stages * p.stage_bytes + WG_STG_BYTES + (2 * WG_MAX_STAGES + 1) * 8 + 16 + 1024;
    GSSD_RETURN_IF_CUDA(allow_max_smem(reinterpret_cast<const void *>(wgrad_kernel)));
    wgrad_kernel<<<units * p.n_chunks, 256, smem, st>>>(m_dy, m_x, m_dw, p);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" size_t gssd_conv_wgrad_bytes(int c_in, int c_out, int groups, int taps) {
    if (c_in <= 0 || c_out <= 0 || groups <= 0) return 0;
    return sizeof(float) * (size_t)taps * ceil_div(c_out, 128) * 128 * (c_in / groups);
}

extern "C" int gssd_head_grad_pm(const float *d_loc, const float *d_conf, int n_img, int height, int width, int n_priors, int prior_off,
                                 int n_anchor, int n_cls, int c_pad, void *out_bf16, float *bias_grad, void *stream) {
    if (!d_loc || !d_conf || !out_bf16) return GSSD_ERR_ARG;
    if (n_img <= 0 || height <= 0 || width <= 0 || n_anchor <= 0 || n_cls <= 0 || prior_off < 0 || n_priors <= 0) return GSSD_ERR_ARG;
    if (c_pad % 64 || c_pad > 128 || n_anchor * (4 + n_cls) > c_pad) return GSSD_ERR_LIMIT;
    const long rows = (long)n_img * (height + 2) * (width + 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (bias_grad) GSSD_RETURN_IF_CUDA(cudaMemsetAsync(bias_grad, 0, sizeof(float) * (size_t)n_anchor * (4 + n_cls), st));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = (int)((rows + 7) / 8 < (long)sms * 4 ? (rows + 7) / 8 : (long)sms * 4);
    head_grad_pm_kernel<<<blocks, 256, 0, st>>>(d_loc, d_conf, n_priors, prior_off, n_anchor, n_cls, (int)rows, height + 2, width + 2,
                                                height, width, c_pad, reinterpret_cast<__nv_bfloat16 *>(out_bf16), bias_grad);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_bn_relu_bwd_pm(const void *dy_bf16, const void *y_bf16, const void *yraw_bf16, const void *add_bf16, int n_img, int c,
                                   int height, int width, const float *chan_sum, const float *gamma, float bn_eps, int relu,
                                   const float *eval_bn_weight, const float *eval_bn_bias,
                                   const float *row_ss_l2, float l2_eps, const float *row_ss_out, float l2_eps_out,
                                   void *out_bf16, float *sums /* [3*c] */, void *stream) {
    if (!dy_bf16 || !y_bf16 || !out_bf16 || !sums) return GSSD_ERR_ARG;
    if (n_img <= 0 || c <= 0 || height <= 0 || width <= 0) return GSSD_ERR_ARG;
    if (c % 256 || c > 1024) return GSSD_ERR_LIMIT;
    if (chan_sum != nullptr && yraw_bf16 == nullptr) return GSSD_ERR_ARG;
    const long rows = (long)n_img * (height + 2) * (width + 2);
    BnBwdArgs a = {};
    a.dy = reinterpret_cast<const __nv_bfloat16 *>(dy_bf16); a.y = reinterpret_cast<const __nv_bfloat16 *>(y_bf16);
    a.yraw = reinterpret_cast<const __nv_bfloat16 *>(yraw_bf16); a.add = reinterpret_cast<const __nv_bfloat16 *>(add_bf16);
    a.out = reinterpret_cast<__nv_bfloat16 *>(out_bf16);
    a.rows = (int)rows; a.c = c; a.hp = height + 2; a.wp = width + 2; a.h = height; a.w = width;
    a.chan_sum = chan_sum; a.gamma = gamma; a.ebn_w = eval_bn_weight; a.ebn_b = eval_bn_bias; a.bn_eps = bn_eps; a.inv_count = (float)(1.0 / ((double)n_img * height * width));
    a.row_ss_l2 = row_ss_l2; a.l2_eps = l2_eps; a.row_ss_out = row_ss_out; a.l2_eps_out = l2_eps_out;
    a.sums = sums; a.relu = relu;
    cudaStream_t st = (cudaStream_t)stream;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 3 * (size_t)c, st));
    const size_t smem = 4 * (size_t)c * sizeof(float);
    const bool reduce_first = chan_sum != nullptr;                           // batch statistics: the sums come first
    switch (c / 256) {
        case 1: return launch_bn_bwd_pm<1>(a, rows, smem, reduce_first, st);
        case 2: return launch_bn_bwd_pm<2>(a, rows, smem, reduce_first, st);
        default: return launch_bn_bwd_pm<4>(a, rows, smem, reduce_first, st);
    }
}
