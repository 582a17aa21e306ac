// gconv.cu — the source block of GSSD: grouped conv -> BN -> ReLU -> [L2Norm] -> 1x1 fuse ->
// BN -> ReLU -> loc/conf 3x3 heads, models/ssd_multiphase_custom_group.py:258-380, as calls of ONE persistent
// implicit-GEMM convolution kernel on the 5th-generation tensor cores.
//
//   activations : bf16 "pixel-major padded" X[rows, C], rows = n_img*(H+2)*(W+2), zero 1-pixel border
//   weights     : bf16 W[c_out, taps*c_in/groups] (K-major)
//   GEMM        : D[m, n] = sum_tap sum_c X[m + dy*(W+2) + dx, g*cg + c] * W[n, tap*cg + c]
//
// Kernel anatomy (384 threads, 1 CTA per SM, persistent over work units = (row tile, group, n-tile)):
//   warp 0      TMA producer: A ring of activation "slabs" (tile rows + halo, 64 channels; loaded once per channel
//               chunk, every 3x3 tap reads it at a row offset) and B ring of weight-tile groups
//   warp 1      MMA issuer: one elected lane, tcgen05.mma (M=128, N=BN, K=16) accumulating in TMEM; one mbarrier wait
//               and one tcgen05.commit per B stage
//   warp 2      TMEM allocation (two accumulator stages)
//   warps 4-11  epilogue: tcgen05.ld (thread = row, 32 columns at a time) -> scale/shift/ReLU/L2Norm/BN-statistics ->
//               bf16 PM tile through a TMA store, or fp32 scatter into loc/conf (head mode)
// DESIGN.md section 4b has the measurements behind each choice.
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"
#include "tmap.cuh"

namespace gssd {

struct ConvParams {
    int rows, hp, wp, h, w;          // padded and interior extents
    int c_out, cg, ng;               // cg = c_in/groups, ng = c_out/groups
    int taps, k_per_tap;             // k_per_tap = cg/64 channel chunks
    int nt_per_group, units;         // column units per row tile = groups * nt_per_group
    int n_mtiles, total_units;
    int halo;                        // rows in front of the tile held by an A slab: W+3 for 3x3, 0 for 1x1
    int a_box_rows, a_nbox, a_stages, b_stages;
    int kg, tg;                      // channel chunks per A stage, weight tiles per B stage (one barrier round trip each)
    int relu;
    float l2_eps;
    const float *scale, *shift, *row_ss_in;
    __nv_bfloat16 *y;
    float *row_ss_out, *chan_sum;
    float *loc, *conf;
    int n_anchor, n_cls, prior_off, n_priors;
    long long *dbg;                  // optional per-CTA wait-cycle counters (gssd_debug_conv_timing)
    int dbg_flags;                   // development: 1 = epilogue skips its work, 2 = producer skips the TMA loads
};

#define TWAIT(acc, stmt) do { if (p.dbg) { long long _t = clock64(); stmt; acc += clock64() - _t; } else { stmt; } } while (0)

constexpr int CONV_STG_BYTES = 8 * 2 * 32 * 64;       // per epilogue warp: 2 x [32 rows][32 bf16] store staging
constexpr int CONV_MAX_STAGES = 8;
constexpr int CONV_BAR_BYTES = (4 * CONV_MAX_STAGES + 4) * 8 + 16;
constexpr int CONV_SMEM_LIMIT = 232448;               // 227 KB

// column sums over the 32 lanes of a warp: on return lane j holds sum_lanes v[j] in v[0]
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, o);
        }
    }
    return v[0];
}

// BN = accumulator columns of one unit, MSUB = 128-row sub-tiles that share every B tile (tile = 128*MSUB rows).
// Work unit = (row tile, group, n-tile); units are dealt round-robin to the persistent CTAs.
//   A ring: "slabs" = (128*MSUB + 2*halo) rows x 64 channels, one per channel chunk.  A 3x3 tap (dy,dx) is the slab
//           read from row (dy+1)*(W+2) + (dx+1) on: the UMMA descriptor simply starts there (the 128-byte swizzle is a
//           function of the absolute shared-memory address, so any row offset inside a 1024-byte-aligned slab is
//           legal), which loads every activation once per unit instead of once per tap.
//   B ring: BN x 64 weight tiles, one per (chunk, tap).
template <int BN, int MSUB>
__global__ void __launch_bounds__(384, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_y, const ConvParams p) {
    constexpr int B_BYTES = BN * 128;
    // accumulator stages: two units in flight (the epilogue of one overlaps the MMAs of the next) unless a unit alone fills
    // the 512 TMEM columns (BN = 256 with two sub-tiles: half the weight traffic per output row, epilogue not overlapped)
    constexpr int ACC = (2 * MSUB * BN <= 512) ? 2 : 1;
    constexpr int TMEM_COLS = (ACC * MSUB * BN) < 32 ? 32 : ACC * MSUB * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int slab_bytes = p.a_nbox * p.a_box_rows * 128;
    const int a_stage_bytes = p.kg * slab_bytes, b_stage_bytes = p.tg * B_BYTES;
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem_a + p.a_stages * a_stage_bytes;
    uint8_t *smem_stg = smem_b + p.b_stages * b_stage_bytes;
    uint64_t *bar_afull = reinterpret_cast<uint64_t *>(smem_stg + CONV_STG_BYTES);
    uint64_t *bar_aempty = bar_afull + CONV_MAX_STAGES;
    uint64_t *bar_bfull = bar_aempty + CONV_MAX_STAGES;
    uint64_t *bar_bempty = bar_bfull + CONV_MAX_STAGES;
    uint64_t *bar_tfull = bar_bempty + CONV_MAX_STAGES;
    uint64_t *bar_tempty = bar_tfull + 2;
    uint32_t *tmem_base_smem = reinterpret_cast<uint32_t *>(bar_tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tensormap(&map_x);
        tc::prefetch_tensormap(&map_w);
        if (p.y != nullptr) tc::prefetch_tensormap(&map_y);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < CONV_MAX_STAGES; ++i) {
            tc::mbar_init(&bar_afull[i], 1); tc::mbar_init(&bar_aempty[i], 1);
            tc::mbar_init(&bar_bfull[i], 1); tc::mbar_init(&bar_bempty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&bar_tfull[i], 1); tc::mbar_init(&bar_tempty[i], 256); }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_base_smem, TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = __shfl_sync(FULL, *tmem_base_smem, 0);          // warp-uniform by construction

    if (warp == 0) {
        // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
        // An A stage = kg channel-chunk slabs, a B stage = tg weight tiles (the 3 taps of one filter row, or the tiles
        // of kg chunks of a 1x1): one mbarrier round trip per stage on either side.
        const bool leader = tc::elect_one();
        {
            const int total_units = p.total_units, units = p.units, nt_per_group = p.nt_per_group, ng = p.ng, cg = p.cg;
            const int k_per_tap = p.k_per_tap, a_stages = p.a_stages, b_stages = p.b_stages, kg = p.kg, tg = p.tg;
            const int taps = p.taps;
            const int a_nbox = p.a_nbox, a_box_rows = p.a_box_rows, halo = p.halo;
            const bool skip = (p.dbg_flags & 2) != 0;
            int as = 0, aph = 0, bs = 0, bph = 0;
            long long w_ae = 0, w_be = 0, t_start = clock64();
            for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
                const int mt = unit / units, u = unit - mt * units;
                const int g = u / nt_per_group, nt = u - g * nt_per_group;
                const int n_row0 = g * ng + nt * BN;
                const int row0 = mt * (128 * MSUB) - halo;
                for (int cc = 0; cc < k_per_tap; cc += kg) {
                    TWAIT(w_ae, tc::mbar_wait(&bar_aempty[as], aph ^ 1));
                    if (leader) {
                        if (skip) { tc::mbar_arrive(&bar_afull[as]); } else {
                            tc::mbar_arrive_expect_tx(&bar_afull[as], a_stage_bytes);
                            for (int j = 0; j < kg; ++j)
                                for (int i = 0; i < a_nbox; ++i)
                                    tc::tma_load_2d(smem_a + as * a_stage_bytes + j * slab_bytes + i * a_box_rows * 128, &map_x,
                                                    &bar_afull[as], g * cg + (cc + j) * 64, row0 + i * a_box_rows);
                        }
                    }
                    __syncwarp();
                    if (++as == a_stages) { as = 0; aph ^= 1; }
                    for (int t0 = 0; t0 < taps; t0 += tg) {                    // 3x3: tg taps per stage; 1x1: one stage of kg chunks
                        TWAIT(w_be, tc::mbar_wait(&bar_bempty[bs], bph ^ 1));
                        if (leader) {
                            if (skip) { tc::mbar_arrive(&bar_bfull[bs]); } else {
                                tc::mbar_arrive_expect_tx(&bar_bfull[bs], b_stage_bytes);
                                for (int j = 0; j < tg; ++j) {
                                    // 3x3: tile j = tap t0 + j of chunk cc; 1x1: tile j = chunk cc + j
                                    const int kcol = taps == 9 ? (t0 + j) * cg + cc * 64 : (cc + j) * 64;
                                    tc::tma_load_2d(smem_b + bs * b_stage_bytes + j * B_BYTES, &map_w, &bar_bfull[bs], kcol, n_row0);
                                }
                            }
                        }
                        __syncwarp();
                        if (++bs == b_stages) { bs = 0; bph ^= 1; }
                        if (taps == 1) break;
                    }
                }
            }
            if (p.dbg && leader) { p.dbg[blockIdx.x * 16 + 0] = w_ae; p.dbg[blockIdx.x * 16 + 1] = w_be; p.dbg[blockIdx.x * 16 + 2] = clock64() - t_start; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
        // Every loop bound and address ingredient is copied into registers first: kernel parameters live in the
        // constant bank and the "memory" clobbers of the PTX wrappers would otherwise force a re-load per iteration.
        const bool leader = tc::elect_one();
        {
            constexpr uint32_t idesc = tc::idesc_bf16_f32(128, BN);
            const int total_units = p.total_units, k_per_tap = p.k_per_tap, a_stages = p.a_stages, b_stages = p.b_stages;
            const int tap_rows = p.taps == 9 ? 3 : 1;                               // 3x3 or 1x1
            const uint32_t row_pitch = (uint32_t)p.wp * 128u;                       // bytes between tap rows inside a slab
            const bool dbg = p.dbg != nullptr;
            const uint32_t a_ring = tc::smem_u32(smem_a), b_ring = tc::smem_u32(smem_b);
            const uint32_t bar_a = tc::smem_u32(bar_afull), bar_b = tc::smem_u32(bar_bfull);
            int as = 0, aph = 0, bs = 0, bph = 0;
            uint32_t it = 0;
            long long w_te = 0, w_af = 0, w_bf = 0, w_issue = 0, w_commit = 0, t_start = clock64();
            const int kg = p.kg, tg = p.tg, trows = p.taps == 9 ? p.tg / 3 : 1;
            // The tensor pipe's instruction queue is shallow (a few MMAs), so every barrier round trip and address
            // computation between two runs of tcgen05.mma shows up as idle pipe (tools/umma_rate.cu: 61% of the MMA
            // rate with 8 MMAs per wait/commit, 80% with 24).  One wait + one commit therefore cover a whole B stage:
            // the 3 taps of a filter row (24 MMAs at MSUB = 2) or kg chunks of a 1x1.
            for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++it) {
                const uint32_t acc = it % ACC;
                TWAIT(w_te, tc::mbar_wait(&bar_tempty[acc], ((it / ACC) & 1) ^ 1));   // the epilogue has drained this accumulator
                tc::fence_after_thread_sync();
                const uint32_t tmem_d0 = tmem_base + acc * MSUB * BN;
                for (int cc = 0; cc < k_per_tap; cc += kg) {
                    TWAIT(w_af, tc::mbar_wait(&bar_afull[as], aph));
                    const uint32_t a_base = a_ring + as * a_stage_bytes;
                    const bool last_chunk = cc + kg >= k_per_tap;
                    for (int ty0 = 0; ty0 < tap_rows; ty0 += trows) {
                        TWAIT(w_bf, tc::mbar_wait(&bar_bfull[bs], bph));
                        tc::fence_after_thread_sync();
                        const uint32_t b_base = b_ring + bs * b_stage_bytes;
                        const uint32_t a_row = a_base + ty0 * row_pitch;
                        const bool last_b = ty0 + trows >= tap_rows;
                        if (leader) {
                            const long long t_i0 = dbg ? clock64() : 0;
#pragma unroll
                            for (int q = 0; q < 9; ++q) {
                                if (q < tg) {
                                    // tile q: 3x3 -> tap (ty0 + q/3, q%3) of the slab; 1x1 -> the slab of chunk cc + q
                                    const uint32_t a_off = tap_rows == 3 ? (uint32_t)(q / 3) * row_pitch + (uint32_t)(q % 3) * 128u
                                                                         : (uint32_t)q * (uint32_t)slab_bytes;
                                    const uint64_t bdesc = tc::smem_desc_k128(b_base + q * B_BYTES);
#pragma unroll
                                    for (int sub = 0; sub < MSUB; ++sub) {
                                        const uint64_t adesc = tc::smem_desc_k128(a_row + a_off + (uint32_t)sub * (128u * 128u));
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            tc::umma_bf16(tmem_d0 + sub * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (cc | ty0 | q | k) != 0);
                                    }
                                }
                            }
                            tc::umma_commit_addr(bar_b + (CONV_MAX_STAGES + bs) * 8);              // bar_bempty: weight tiles free
                            if (last_b) {
                                tc::umma_commit_addr(bar_a + (CONV_MAX_STAGES + as) * 8);          // bar_aempty: slabs free
                                if (last_chunk) tc::umma_commit(&bar_tfull[acc]);
                            }
                            if (dbg) w_issue += clock64() - t_i0;
                        }
                        __syncwarp();
                        if (++bs == b_stages) { bs = 0; bph ^= 1; }
                    }
                    if (++as == a_stages) { as = 0; aph ^= 1; }
                }
            }
            if (p.dbg && leader) { p.dbg[blockIdx.x * 16 + 3] = w_te; p.dbg[blockIdx.x * 16 + 4] = w_af; p.dbg[blockIdx.x * 16 + 5] = w_bf; p.dbg[blockIdx.x * 16 + 6] = clock64() - t_start; p.dbg[blockIdx.x * 16 + 9] = w_issue; p.dbg[blockIdx.x * 16 + 10] = w_commit; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps =====================
        // warp % 4 selects the TMEM lane quarter (a hardware rule); the two warps of a quarter split each unit's
        // 32-column chunks between them.  One warp per scheduler cannot hide its own instruction latencies, so the
        // chunk body is kept short: flags are tested once per chunk, never per element.
        const int ew = warp & 3, eset = (warp - 4) >> 2;
        const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
        constexpr int NCH = BN / 32;
        const bool head = p.loc != nullptr, has_rs = p.row_ss_in != nullptr, has_scale = p.scale != nullptr, has_shift = p.shift != nullptr;
        const bool do_stats = p.chan_sum != nullptr, do_relu = p.relu != 0, do_y = p.y != nullptr, do_ss = p.row_ss_out != nullptr;
        const int loc_cols = 4 * p.n_anchor, head_cols = loc_cols + p.n_anchor * p.n_cls, c_out = p.c_out;
        const int hpwp = p.hp * p.wp, wp = p.wp, hh = p.h, ww = p.w, rows = p.rows, units = p.units, nt_per_group = p.nt_per_group, ng = p.ng;
        const int total_units = p.total_units;
        const float *scale = p.scale, *shift = p.shift;
        uint8_t *stg = smem_stg + (warp - 4) * (2 * 32 * 64);
        uint32_t it = 0, n_store = 0;
        long long w_tf = 0, w_ld = 0, w_sg = 0, t_start = clock64();
        for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++it) {
            const uint32_t acc = it % ACC;
            const int mt = unit / units, u = unit - mt * units;
            const int g = u / nt_per_group, nt = u - g * nt_per_group;
            const int col_base = g * ng + nt * BN;
            TWAIT(w_tf, tc::mbar_wait(&bar_tfull[acc], (it / ACC) & 1));
            tc::fence_after_thread_sync();
            int cur_sub = -1, m = 0, m_warp = 0;
            bool interior = false;
            float rs = 1.f, ss = 0.f;
            float *loc_row = nullptr, *conf_row = nullptr;
#pragma unroll 1
            for (int item = ((p.dbg_flags & 1) ? MSUB * NCH : eset); item < MSUB * NCH; item += 2) {
                const int sub = item / NCH, c0 = (item - sub * NCH) * 32;
                if (sub != cur_sub) {
                    if (do_ss && cur_sub >= 0 && interior) atomicAdd(p.row_ss_out + m, ss);
                    cur_sub = sub; ss = 0.f;
                    m_warp = mt * (128 * MSUB) + sub * 128 + ew * 32;
                    m = m_warp + lane;
                    const int img = m / hpwp, rem = m - img * hpwp;
                    const int py = rem / wp, px = rem - py * wp;
                    interior = m < rows && py >= 1 && py <= hh && px >= 1 && px <= ww;
                    rs = (has_rs && interior) ? 1.f / (sqrtf(__ldg(p.row_ss_in + m)) + p.l2_eps) : 1.f;
                    if (head && interior) {
                        const size_t prior = (size_t)img * p.n_priors + p.prior_off + (size_t)((py - 1) * ww + (px - 1)) * p.n_anchor;
                        loc_row = p.loc + prior * 4;
                        conf_row = p.conf + prior * p.n_cls;
                    }
                }
                const int col0 = col_base + c0;
                float sc[32], sh[32];
                if (!head) {                                                  // issue the coefficient loads before the TMEM wait
                    if (has_scale) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 t = __ldg(reinterpret_cast<const float4 *>(scale + col0 + j));
                            sc[j] = t.x; sc[j + 1] = t.y; sc[j + 2] = t.z; sc[j + 3] = t.w;
                        }
                    }
                    if (has_shift) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 t = __ldg(reinterpret_cast<const float4 *>(shift + col0 + j));
                            sh[j] = t.x; sh[j + 1] = t.y; sh[j + 2] = t.z; sh[j + 3] = t.w;
                        }
                    }
                }
                uint32_t r[32];
                TWAIT(w_ld, { tc::tmem_ld_32x32(tmem_base + lane_base + (acc * MSUB + sub) * BN + c0, r); tc::tmem_ld_wait(); });
                float v[32];
                if (head) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int c = col0 + j;
                        v[j] = __uint_as_float(r[j]) + ((has_shift && c < c_out) ? __ldg(shift + c) : 0.f);
                    }
                    if (interior) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int c = col0 + j;
                            if (c + 4 <= loc_cols) {
                                *reinterpret_cast<float4 *>(loc_row + c) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            } else {
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    if (c + q < head_cols) conf_row[c + q - loc_cols] = v[j + q];
                            }
                        }
                    }
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (has_rs) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= rs;
                }
                if (has_scale) {
                    if (has_shift) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= sc[j];
                    }
                } else if (has_shift) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += sh[j];
                }
                if (do_stats) {                                               // train-mode BN statistics of the raw output
                    float s1[32], s2[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { s1[j] = interior ? v[j] : 0.f; s2[j] = s1[j] * s1[j]; }
                    const float a = warp_transpose_sum(s1, lane), b2 = warp_transpose_sum(s2, lane);
                    atomicAdd(p.chan_sum + col0 + lane, a);
                    atomicAdd(p.chan_sum + c_out + col0 + lane, b2);
                }
                if (do_relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (do_ss) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) ss = fmaf(v[j], v[j], ss);
                }
                if (do_y) {
                    // 32 output columns = one 32x32 bf16 box (64-byte rows, TMA's 64-byte swizzle) per store
                    uint8_t *buf = stg + (n_store & 1) * (32 * 64);
                    TWAIT(w_sg, { if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); __syncwarp(); });   // buffer free again
                    const uint32_t keep = interior ? 0xffffffffu : 0u;        // border rows are written as zeros
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 o;
                        o.x = tc::pack_bf16x2(v[8 * q + 0], v[8 * q + 1]) & keep;
                        o.y = tc::pack_bf16x2(v[8 * q + 2], v[8 * q + 3]) & keep;
                        o.z = tc::pack_bf16x2(v[8 * q + 4], v[8 * q + 5]) & keep;
                        o.w = tc::pack_bf16x2(v[8 * q + 6], v[8 * q + 7]) & keep;
                        const int chunk = q ^ ((lane >> 1) & 3);
                        *reinterpret_cast<uint4 *>(buf + lane * 64 + chunk * 16) = o;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && m_warp < rows) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(&map_y), "r"(tc::smem_u32(buf)), "r"(col0), "r"(m_warp) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++n_store;
                }
            }
            if (do_ss && cur_sub >= 0 && interior) atomicAdd(p.row_ss_out + m, ss);
            tc::fence_before_thread_sync();
            tc::mbar_arrive(&bar_tempty[acc]);                                // 256 arrivals free the accumulator
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (p.dbg && warp == 4 && lane == 0) { p.dbg[blockIdx.x * 16 + 7] = w_tf; p.dbg[blockIdx.x * 16 + 8] = clock64() - t_start; p.dbg[blockIdx.x * 16 + 11] = w_ld; p.dbg[blockIdx.x * 16 + 12] = w_sg; }
    }

    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 2) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- host side ---------------------------------------------------------------------------------------
// bf16 matrix [rows, cols] (cols innermost), box = box_rows x 64 columns, 128-byte swizzle, zero fill outside
static int make_map_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols = 64) {
    return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, box_rows, box_cols,
                        box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

template <int BN, int MSUB>
static int launch_conv(const CUtensorMap &mx, const CUtensorMap &mw, const CUtensorMap &my, const ConvParams &p, size_t smem,
                       cudaStream_t st) {
    const void *kern = reinterpret_cast<const void *>(&conv_igemm_kernel<BN, MSUB>);
    GSSD_RETURN_IF_CUDA(allow_max_smem(kern));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // persistent: one CTA per SM and no more CTAs than units; equalise the units per CTA
    const int waves = ceil_div(p.total_units, sms);
    const int grid = ceil_div(p.total_units, waves);
    conv_igemm_kernel<BN, MSUB><<<grid, 384, smem, st>>>(mx, mw, my, p);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

// ---- packing / layout kernels ------------------------------------------------------------------------
// w[c_out, cg, kh, kw] fp32 -> out[c_out(+pad), taps*cg] bf16 with k = tap*cg + c
__global__ void pack_weights_kernel(const float *__restrict__ w, int c_out, int cg, int taps, int c_in_total,
                                    const float *__restrict__ in_scale, int ng, __nv_bfloat16 *__restrict__ out, long total) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % (taps * cg));
        const int co = (int)(i / (taps * cg));
        const int t = k / cg, c = k - t * cg;
        float v = w[((size_t)co * cg + c) * taps + t];
        if (in_scale != nullptr) v *= in_scale[(co / ng) * cg + c];
        (void)c_in_total;
        out[i] = __float2bfloat16_rn(v);
    }
}

// x[n, c, h, w] fp32 -> y PM bf16 [(n, h+2, w+2), c]; one CTA per (padded row, image, 64-channel slab), transposed through smem
// (one CTA per row and image looping over the slabs left most SMs idle: 103 us for 4 x 1024 x 38 x 38, 36 MB of traffic)
__global__ void __launch_bounds__(256) nchw_to_pm_kernel(const float *__restrict__ x, int c, int h, int w, __nv_bfloat16 *__restrict__ y) {
    extern __shared__ float tile[];                                          // [64][w + 1]
    const int py = blockIdx.x, img = blockIdx.y, c0 = blockIdx.z * 64, hp = h + 2, wp = w + 2;
    const int nc = min(64, c - c0);
    __nv_bfloat16 *yrow = y + ((size_t)img * hp + py) * wp * c + c0;
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
    if (py == 0 || py == hp - 1) {
        for (int i = threadIdx.x; i < wp * nc; i += blockDim.x) yrow[(size_t)(i / nc) * c + i % nc] = zero;
        return;
    }
    for (int i = threadIdx.x; i < nc; i += blockDim.x) { yrow[i] = zero; yrow[(size_t)(wp - 1) * c + i] = zero; }
    const int pitch = w | 1;                                                 // odd: the transposed reads below are at most 2-way conflicts
    if (nc == 64 && (c & 1) == 0) {                                          // (even c: 4-byte aligned channel pairs)
        // no index arithmetic per element (the generic loop below spends its time on two integer divisions per element:
        // 149 us for 32 x 256 x 75 x 75 = 1.9 TB/s, ncu): warps over channels / lanes over pixels in, warps over pixels / lanes
        // over channel pairs out — 128-byte loads and 128-byte stores per warp
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int cc = warp; cc < 64; cc += nw) {
            const float *src = x + (((size_t)img * c + c0 + cc) * h + (py - 1)) * w;
            for (int xx = lane; xx < w; xx += 32) tile[cc * pitch + xx] = src[xx];
        }
        __syncthreads();
        for (int xx = warp; xx < w; xx += nw) {
            const float v0 = tile[(2 * lane) * pitch + xx], v1 = tile[(2 * lane + 1) * pitch + xx];
            *reinterpret_cast<__nv_bfloat162 *>(yrow + (size_t)(xx + 1) * c + 2 * lane) = __floats2bfloat162_rn(v0, v1);
        }
        return;
    }
    for (int i = threadIdx.x; i < nc * w; i += blockDim.x) {
        const int cc = i / w, xx = i - cc * w;
        tile[cc * pitch + xx] = x[(((size_t)img * c + c0 + cc) * h + (py - 1)) * w + xx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nc * w; i += blockDim.x) {
        const int xx = i / nc, cc = i - xx * nc;
        yrow[(size_t)(xx + 1) * c + cc] = __float2bfloat16_rn(tile[cc * pitch + xx]);
    }
}

// PM (bf16 or fp32) -> NCHW fp32, the interior pixels; same decomposition
template <typename T>
__global__ void __launch_bounds__(256) pm_to_nchw_kernel(const T *__restrict__ x, int c, int h, int w, float *__restrict__ y) {
    extern __shared__ float tile[];                                          // [64][w + 1]
    const int yy = blockIdx.x, img = blockIdx.y, c0 = blockIdx.z * 64, hp = h + 2, wp = w + 2;
    const int nc = min(64, c - c0);
    const T *xrow = x + (((size_t)img * hp + yy + 1) * wp + 1) * c + c0;
    const int pitch = w | 1;
    if (nc == 64 && (c & 1) == 0 && sizeof(T) == 2) {                        // see nchw_to_pm_kernel
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int xx = warp; xx < w; xx += nw) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162 *>(reinterpret_cast<const __nv_bfloat16 *>(xrow) + (size_t)xx * c + 2 * lane);
            tile[(2 * lane) * pitch + xx] = __low2float(v);
            tile[(2 * lane + 1) * pitch + xx] = __high2float(v);
        }
        __syncthreads();
        for (int cc = warp; cc < 64; cc += nw) {
            float *dst = y + (((size_t)img * c + c0 + cc) * h + yy) * w;
            for (int xx = lane; xx < w; xx += 32) dst[xx] = tile[cc * pitch + xx];
        }
        return;
    }
    for (int i = threadIdx.x; i < nc * w; i += blockDim.x) {
        const int xx = i / nc, cc = i - xx * nc;
        tile[cc * pitch + xx] = (float)xrow[(size_t)xx * c + cc];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nc * w; i += blockDim.x) {
        const int cc = i / w, xx = i - cc * w;
        y[(((size_t)img * c + c0 + cc) * h + yy) * w + xx] = tile[cc * pitch + xx];
    }
}

// train-mode BN (+ReLU) in place on a PM tensor; one warp per pixel row of c channels
__global__ void __launch_bounds__(256) bn_act_pm_kernel(__nv_bfloat16 *y, __nv_bfloat16 *y_out, int rows, int c, int hp, int wp, int h, int w,
                                                        const float *__restrict__ chan_sum, const float *__restrict__ gamma,
                                                        const float *__restrict__ beta, float bn_eps, float inv_count, int relu,
                                                        float *__restrict__ row_ss_out) {
    // per channel (a, b) = (rstd*gamma, beta - mean*a) as one float2.  With whole groups of 256 channels, channel
    // (slot*32 + lane)*8 + k is kept at [(slot*8 + k)*32 + lane]: the lanes of a warp, which read the same (slot, k) together and
    // are 8 channels apart, then hit consecutive words instead of the same banks (see bn_bwd_pm_kernel, gconv_bwd.cu)
    extern __shared__ __align__(8) float coef[];
    float2 *coef2 = reinterpret_cast<float2 *>(coef);
    const bool lane_major = (c & 255) == 0;
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const float mean = chan_sum[i] * inv_count;
        const float var = fmaxf(chan_sum[c + i] * inv_count - mean * mean, 0.f);
        const float a = rsqrtf(var + bn_eps) * (gamma ? gamma[i] : 1.f);
        const int at = lane_major ? ((i >> 8) * 8 + (i & 7)) * 32 + ((i >> 3) & 31) : i;
        coef2[at] = make_float2(a, (beta ? beta[i] : 0.f) - mean * a);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int m = blockIdx.x * wpb + warp; m < rows; m += gridDim.x * wpb) {
        const int rem = m % (hp * wp), py = rem / wp, px = rem - py * wp;
        const bool interior = py >= 1 && py <= h && px >= 1 && px <= w;
        float ss = 0.f;
        if (!interior && y_out != y) {                                       // out of place: the border of the output is zero too
            uint4 *orow = reinterpret_cast<uint4 *>(y_out + (size_t)m * c);
            for (int i = lane; i < c / 8; i += 32) orow[i] = make_uint4(0, 0, 0, 0);
        }
        if (interior) {
            const uint4 *row = reinterpret_cast<const uint4 *>(y + (size_t)m * c);
            uint4 *orow = reinterpret_cast<uint4 *>(y_out + (size_t)m * c);
            for (int i = lane; i < c / 8; i += 32) {
                uint4 q = row[i];
                uint32_t *qw = reinterpret_cast<uint32_t *>(&q);
                const int base = lane_major ? (i >> 5) * 256 + lane : i * 8;  // + k*32 / + k for channel i*8 + k
                const int kstep = lane_major ? 32 : 1;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __nv_bfloat162 b2 = *reinterpret_cast<__nv_bfloat162 *>(&qw[j]);
                    const float2 c0 = coef2[base + (2 * j) * kstep], c1 = coef2[base + (2 * j + 1) * kstep];
                    float lo = __low2float(b2) * c0.x + c0.y;
                    float hi = __high2float(b2) * c1.x + c1.y;
                    if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                    const uint32_t packed = tc::pack_bf16x2(lo, hi);
                    __nv_bfloat162 rb = *reinterpret_cast<const __nv_bfloat162 *>(&packed);
                    lo = __low2float(rb); hi = __high2float(rb);                // the values the consumer will read
                    ss += lo * lo + hi * hi;
                    qw[j] = packed;
                }
                orow[i] = q;
            }
        }
        if (row_ss_out != nullptr) {
#pragma unroll
            for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
            if (lane == 0) row_ss_out[m] = ss;
        }
    }
}

// max pooling on PM tensors (nn.MaxPool2d semantics: windows clipped to the image, padding never wins); 8 channels / thread
__global__ void __launch_bounds__(256) maxpool_pm_kernel(const __nv_bfloat16 *__restrict__ x, int c, int h, int w, int oh, int ow,
                                                         int k, int stride, int pad, __nv_bfloat16 *__restrict__ y, long total) {
    const int c8 = c / 8;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c8);
        long r = i / c8;
        const int px = (int)(r % (ow + 2)); r /= (ow + 2);
        const int py = (int)(r % (oh + 2));
        const int img = (int)(r / (oh + 2));
        uint4 out = make_uint4(0, 0, 0, 0);
        if (py >= 1 && py <= oh && px >= 1 && px <= ow) {
            float m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
            const int y0 = (py - 1) * stride - pad, x0 = (px - 1) * stride - pad;
            for (int ky = 0; ky < k; ++ky) {
                const int iy = y0 + ky;
                if (iy < 0 || iy >= h) continue;
                for (int kx = 0; kx < k; ++kx) {
                    const int ix = x0 + kx;
                    if (ix < 0 || ix >= w) continue;
                    const uint4 v = *reinterpret_cast<const uint4 *>(x + (((size_t)img * (h + 2) + iy + 1) * (w + 2) + ix + 1) * c + cc * 8);
                    const uint32_t *vw = reinterpret_cast<const uint32_t *>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162 *>(&vw[j]);
                        m[2 * j] = fmaxf(m[2 * j], __low2float(b2));
                        m[2 * j + 1] = fmaxf(m[2 * j + 1], __high2float(b2));
                    }
                }
            }
            out.x = tc::pack_bf16x2(m[0], m[1]); out.y = tc::pack_bf16x2(m[2], m[3]);
            out.z = tc::pack_bf16x2(m[4], m[5]); out.w = tc::pack_bf16x2(m[6], m[7]);
        }
        *reinterpret_cast<uint4 *>(y + i * 8) = out;
    }
}

__global__ void bn_mean_var_kernel(const float *__restrict__ chan_sum, int c, float inv_count, float unbias, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c) {
        const float mean = chan_sum[i] * inv_count;
        out[i] = mean;
        out[c + i] = fmaxf(chan_sum[c + i] * inv_count - mean * mean, 0.f) * unbias;
    }
}

}  // namespace gssd

using namespace gssd;

static long long *g_conv_dbg = nullptr;
static int g_conv_dbg_flags = 0;
extern "C" __attribute__((visibility("default"))) void gssd_debug_conv_flags(int f) { g_conv_dbg_flags = f; }
/* development aid (not in gssd.h): per-CTA wait-cycle counters [grid][16] of the following gssd_conv_igemm launches */
extern "C" __attribute__((visibility("default"))) void gssd_debug_conv_timing(long long *buf) { g_conv_dbg = buf; }

extern "C" int gssd_conv_igemm(const gssd_conv_desc *d, void *stream) {
    if (d == nullptr || d->x == nullptr || d->w == nullptr) return GSSD_ERR_ARG;
    if (d->n_img <= 0 || d->height <= 0 || d->width <= 0 || d->c_in <= 0 || d->c_out <= 0 || d->groups <= 0) return GSSD_ERR_ARG;
    if (d->taps != 1 && d->taps != 9) return GSSD_ERR_ARG;
    if (d->c_in % d->groups || d->c_out % d->groups) return GSSD_ERR_ARG;
    const bool head = d->loc != nullptr;
    if (head && (d->conf == nullptr || d->y != nullptr || d->n_anchor <= 0 || d->n_cls <= 0 || d->n_priors <= 0 || d->prior_off < 0)) return GSSD_ERR_ARG;
    if (head && (d->groups != 1 || d->c_out != d->n_anchor * (4 + d->n_cls))) return GSSD_ERR_ARG;
    if (!head && d->y == nullptr && d->chan_sum == nullptr && d->row_ss_out == nullptr) return GSSD_ERR_ARG;
    const int cg = d->c_in / d->groups, ng = d->c_out / d->groups;
    if (cg % 64) return GSSD_ERR_LIMIT;
    int bn;
    if (head) {
        if (d->c_out > 64) return GSSD_ERR_LIMIT;
        bn = d->c_out <= 32 ? 32 : 64;
    } else {
        bn = ng % 256 == 0 ? 256 : (ng % 128 == 0 ? 128 : (ng % 64 == 0 ? 64 : 0));
        if (bn == 0) return GSSD_ERR_LIMIT;
        if (bn == 256 && d->taps == 9) bn = 128;                 // a filter row of three 256-wide weight tiles leaves no room for the slabs
    }
    const long rows_l = (long)d->n_img * (d->height + 2) * (d->width + 2);
    if (rows_l > (1l << 30)) return GSSD_ERR_LIMIT;

    ConvParams p;
    p.rows = (int)rows_l; p.hp = d->height + 2; p.wp = d->width + 2; p.h = d->height; p.w = d->width;
    p.c_out = d->c_out; p.cg = cg; p.ng = ng; p.taps = d->taps; p.k_per_tap = cg / 64;
    p.nt_per_group = head ? 1 : ng / bn; p.units = d->groups * p.nt_per_group;
    p.relu = d->relu; p.l2_eps = d->l2_eps;
    p.scale = d->scale; p.shift = d->shift; p.row_ss_in = d->row_ss_in;
    p.y = reinterpret_cast<__nv_bfloat16 *>(d->y); p.row_ss_out = d->row_ss_out; p.chan_sum = d->chan_sum;
    p.dbg = g_conv_dbg; p.dbg_flags = g_conv_dbg_flags;
    p.loc = d->loc; p.conf = d->conf; p.n_anchor = d->n_anchor; p.n_cls = d->n_cls; p.prior_off = d->prior_off; p.n_priors = d->n_priors;

    // ---- tile geometry ----
    // two 128-row sub-tiles share every weight tile (heads: tiny weight tiles, nothing to share).  At BN = 256 a pair would
    // fill all 512 TMEM columns and serialise the epilogue behind the MMAs: measured slower (fuse_11 50 us against 38 us),
    // although it halves the weight bytes per output row
    const int msub = (bn <= 128 && !head) ? 2 : 1;
    const int tile_rows = 128 * msub;
    p.halo = d->taps == 9 ? p.wp + 1 : 0;
    const int slab_rows = tile_rows + 2 * p.halo;
    p.a_nbox = ceil_div(slab_rows, 256);
    p.a_box_rows = ceil_div(ceil_div(slab_rows, p.a_nbox), 8) * 8;
    const int slab_bytes = p.a_nbox * p.a_box_rows * 128, b_bytes = bn * 128;
    const int budget = CONV_SMEM_LIMIT - 1024 - CONV_BAR_BYTES - CONV_STG_BYTES;
    int a_stage_bytes, b_stage_bytes;
    if (d->taps == 9) {
        p.kg = 1;
        p.tg = 9 * b_bytes <= 40 * 1024 ? 9 : 3;                              // B stage = the whole 3x3 filter when small, else one filter row
        a_stage_bytes = slab_bytes; b_stage_bytes = p.tg * b_bytes;
        p.a_stages = (budget - 2 * b_stage_bytes) / a_stage_bytes;
        if (p.a_stages > 4) p.a_stages = 4;
        if (p.a_stages < 2) return GSSD_ERR_LIMIT;                            // feature map too wide for the slab scheme
        p.b_stages = (budget - p.a_stages * a_stage_bytes) / b_stage_bytes;
    } else {
        p.kg = p.k_per_tap % 2 == 0 ? 2 : 1;                                  // A/B stage = two channel chunks when they fit twice
        if (2 * p.kg * (slab_bytes + b_bytes) > budget) p.kg = 1;
        p.tg = p.kg;
        a_stage_bytes = p.kg * slab_bytes; b_stage_bytes = p.tg * b_bytes;
        p.a_stages = p.b_stages = budget / (a_stage_bytes + b_stage_bytes);   // the two rings advance in lockstep
    }
    if (p.a_stages > CONV_MAX_STAGES) p.a_stages = CONV_MAX_STAGES;
    if (p.b_stages > CONV_MAX_STAGES) p.b_stages = CONV_MAX_STAGES;
    if (p.a_stages < 2 || p.b_stages < 2) return GSSD_ERR_LIMIT;              // feature map too wide for the slab scheme
    const size_t smem = (size_t)p.a_stages * a_stage_bytes + (size_t)p.b_stages * b_stage_bytes + CONV_STG_BYTES + CONV_BAR_BYTES + 1024;
    p.n_mtiles = ceil_div(p.rows, tile_rows);
    p.total_units = p.n_mtiles * p.units;

    CUtensorMap mx, mw, my;
    int rc = make_map_2d(&mx, d->x, (uint64_t)p.rows, (uint64_t)d->c_in, (uint32_t)p.a_box_rows);
    if (rc) return rc;
    // head weights are stored zero-padded to a whole number of 32-row boxes (gssd_conv_pack_weights)
    const uint64_t w_rows = head ? (uint64_t)ceil_div(d->c_out, 32) * 32 : (uint64_t)d->c_out;
    rc = make_map_2d(&mw, d->w, w_rows, (uint64_t)d->taps * cg, (uint32_t)bn);
    if (rc) return rc;
    if (d->y != nullptr) {
        rc = make_map_2d(&my, d->y, (uint64_t)p.rows, (uint64_t)d->c_out, 32, 32);
        if (rc) return rc;
    } else {
        my = mx;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (d->row_ss_out != nullptr) GSSD_RETURN_IF_CUDA(cudaMemsetAsync(d->row_ss_out, 0, sizeof(float) * (size_t)p.rows, st));
    switch (bn) {
        case 32: return launch_conv<32, 1>(mx, mw, my, p, smem, st);
        case 64: return msub == 2 ? launch_conv<64, 2>(mx, mw, my, p, smem, st) : launch_conv<64, 1>(mx, mw, my, p, smem, st);
        case 128: return launch_conv<128, 2>(mx, mw, my, p, smem, st);
        default: return launch_conv<256, 1>(mx, mw, my, p, smem, st);
    }
}

extern "C" int gssd_conv_pack_weights(const float *w, int c_out, int c_in_per_group, int groups, int taps,
                                      const float *in_scale, void *out_bf16, void *stream) {
    if (w == nullptr || out_bf16 == nullptr || c_out <= 0 || c_in_per_group <= 0 || groups <= 0 || (taps != 1 && taps != 9)) return GSSD_ERR_ARG;
    if (c_out % groups) return GSSD_ERR_ARG;
    const long total = (long)c_out * taps * c_in_per_group;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, c_out, c_in_per_group, taps, c_in_per_group * groups, in_scale,
                                                                  c_out / groups, reinterpret_cast<__nv_bfloat16 *>(out_bf16), total);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_nchw_to_pm(const float *x, int n_img, int c, int h, int w, void *y_bf16, void *stream) {
    if (x == nullptr || y_bf16 == nullptr || n_img <= 0 || c <= 0 || h <= 0 || w <= 0) return GSSD_ERR_ARG;
    if (w > 512) return GSSD_ERR_LIMIT;
    if (n_img > 65535 || (c + 63) / 64 > 65535) return GSSD_ERR_LIMIT;
    nchw_to_pm_kernel<<<dim3(h + 2, n_img, (c + 63) / 64), 256, 64 * (w + 1) * sizeof(float), (cudaStream_t)stream>>>(
        x, c, h, w, reinterpret_cast<__nv_bfloat16 *>(y_bf16));
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_pm_to_nchw(const void *x_bf16, int n_img, int c, int h, int w, float *y, void *stream) {
    if (x_bf16 == nullptr || y == nullptr || n_img <= 0 || c <= 0 || h <= 0 || w <= 0) return GSSD_ERR_ARG;
    if (w > 512) return GSSD_ERR_LIMIT;
    if (n_img > 65535 || (c + 63) / 64 > 65535) return GSSD_ERR_LIMIT;
    pm_to_nchw_kernel<__nv_bfloat16><<<dim3(h, n_img, (c + 63) / 64), 256, 64 * (w + 1) * sizeof(float), (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16 *>(x_bf16), c, h, w, y);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

extern "C" int gssd_pmf32_to_nchw(const float *x_pm, int n_img, int c, int h, int w, float *y, void *stream) {
    if (x_pm == nullptr || y == nullptr || n_img <= 0 || c <= 0 || h <= 0 || w <= 0) return GSSD_ERR_ARG;
    if (w > 512 || n_img > 65535 || (c + 63) / 64 > 65535) return GSSD_ERR_LIMIT;
    pm_to_nchw_kernel<float><<<dim3(h, n_img, (c + 63) / 64), 256, 64 * (w + 1) * sizeof(float), (cudaStream_t)stream>>>(x_pm, c, h, w, y);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

static int bn_act_pm_impl(void *y_bf16, void *y_out_bf16, int n_img, int c, int h, int w, const float *chan_sum, const float *gamma,
                          const float *beta, float bn_eps, int relu, float *row_ss_out, float *mean_var_out, void *stream) {
    if (y_bf16 == nullptr || chan_sum == nullptr || n_img <= 0 || c <= 0 || h <= 0 || w <= 0) return GSSD_ERR_ARG;
    if (c % 8 || c > 8192) return GSSD_ERR_LIMIT;
    const long rows = (long)n_img * (h + 2) * (w + 2);
    const double count = (double)n_img * h * w;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = (int)((rows + 7) / 8 < (long)sms * 8 ? (rows + 7) / 8 : (long)sms * 8);
    bn_act_pm_kernel<<<blocks, 256, 2 * c * sizeof(float), (cudaStream_t)stream>>>(
        reinterpret_cast<__nv_bfloat16 *>(y_bf16), reinterpret_cast<__nv_bfloat16 *>(y_out_bf16 ? y_out_bf16 : y_bf16), (int)rows, c, h + 2, w + 2, h, w, chan_sum, gamma, beta, bn_eps,
        (float)(1.0 / count), relu, row_ss_out);
    GSSD_AFTER_LAUNCH();
    if (mean_var_out != nullptr) {
        bn_mean_var_kernel<<<ceil_div(c, 256), 256, 0, (cudaStream_t)stream>>>(chan_sum, c, (float)(1.0 / count),
                                                                              (float)(count > 1 ? count / (count - 1) : 1.0), mean_var_out);
        GSSD_AFTER_LAUNCH();
    }
    return GSSD_OK;
}

extern "C" int gssd_bn_act_pm(void *y_bf16, int n_img, int c, int h, int w, const float *chan_sum, const float *gamma,
                              const float *beta, float bn_eps, int relu, float *row_ss_out, float *mean_var_out, void *stream) {
    return bn_act_pm_impl(y_bf16, nullptr, n_img, c, h, w, chan_sum, gamma, beta, bn_eps, relu, row_ss_out, mean_var_out, stream);
}

extern "C" int gssd_bn_act_pm_to(const void *y_bf16, void *y_out_bf16, int n_img, int c, int h, int w, const float *chan_sum,
                                 const float *gamma, const float *beta, float bn_eps, int relu, float *row_ss_out,
                                 float *mean_var_out, void *stream) {
    if (y_out_bf16 == nullptr) return GSSD_ERR_ARG;
    return bn_act_pm_impl(const_cast<void *>(y_bf16), y_out_bf16, n_img, c, h, w, chan_sum, gamma, beta, bn_eps, relu, row_ss_out,
                          mean_var_out, stream);
}

extern "C" int gssd_maxpool_pm(const void *x_bf16, int n_img, int c, int h, int w, int kernel, int stride, int pad, int ceil_mode,
                               void *y_bf16, int *out_h, int *out_w, void *stream) {
    if (n_img <= 0 || c <= 0 || h <= 0 || w <= 0 || kernel <= 0 || stride <= 0 || pad < 0 || 2 * pad > kernel) return GSSD_ERR_ARG;
    if (c % 8) return GSSD_ERR_LIMIT;
    auto osz = [&](int in) {                                                   // nn.MaxPool2d output size
        const int num = in + 2 * pad - kernel;
        int o = (ceil_mode ? (num + stride - 1) / stride : num / stride) + 1;
        if (ceil_mode && (o - 1) * stride >= in + pad) --o;                    // the last window must start inside the input
        return o;
    };
    const int oh = osz(h), ow = osz(w);
    if (out_h) *out_h = oh;
    if (out_w) *out_w = ow;
    if (oh <= 0 || ow <= 0) return GSSD_ERR_ARG;
    if (x_bf16 == nullptr || y_bf16 == nullptr) return (x_bf16 == nullptr && y_bf16 == nullptr) ? GSSD_OK : GSSD_ERR_ARG;   // size query
    const long total = (long)n_img * (oh + 2) * (ow + 2) * (c / 8);
    const int blocks = (int)((total + 255) / 256 < 148l * 16 ? (total + 255) / 256 : 148l * 16);
    maxpool_pm_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16 *>(x_bf16), c, h, w, oh, ow, kernel,
                                                                 stride, pad, reinterpret_cast<__nv_bfloat16 *>(y_bf16), total);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}
