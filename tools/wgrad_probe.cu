// wgrad_probe.cu — the stand-alone prototype of the weight-gradient kernel (grouped 3x3, 128 channels per group) with a host check
// at small sizes and a timing at the configs[1] size.  Ran green on a B200 at the start of round 2 (profiles/r2_staged_probes.txt:
// exact, 68.6 us); the library kernel that grew out of it is gssd_conv_wgrad (csrc/gconv_bwd.cu).
//
//   dW[g][tap][co][ci] = sum_row dY[row][g*128 + co] * X[row + (dy-1)*(W+2) + (dx-1)][g*128 + ci]      tap = dy*3 + dx
//
// X and dY are the bf16 "pixel-major padded" tensors of gconv.cu ([rows, C], rows = n*(H+2)*(W+2), zero border), so the
// sum may run over ALL rows: border rows of dY are zero, rows outside the tensor are zero-filled by TMA.
// As a GEMM: M = co (128 per group), N = ci (128 per group), K = rows.  Both operands are MN-major for it (channels are
// contiguous), which tcgen05 takes through a_major / b_major; a TMA box of [rows x 64 channels] with SWIZZLE_128B is the
// canonical MN-major atom sequence (SBO = 1024 bytes to the next 8 rows, LBO = one box to the next 64 channels).
//
// One CTA per (group, filter row dy, K-chunk): the three taps dx = 0..2 of a filter row read ONE slab of X (34 rows per 32
// rows of dY) at row offsets 0 / 1 / 2 — the descriptor start address moves by 128 bytes, as for the forward kernel's taps —
// and accumulate into 3 x 128 TMEM columns.  4 groups x 3 rows x 12 chunks = 144 CTAs.  Per 32-row stage: 16.9 KB of TMA
// loads against 6 MMAs (M = N = 128, K = 16: 64 cycles each), so the loop is bound by the tensor pipe, not by L2.
// The epilogue adds the fp32 tiles into dW with TMA reductions (cp.reduce.async.bulk.tensor .add): 12 CTAs add into each tile.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/wgrad_probe tools/wgrad_probe.cu && tools/wgrad_probe
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../grouped_ssd_pytorch_b200/csrc/tc.cuh"

using namespace gssd;

constexpr int KS = 32;                      // rows of dY per stage (two K = 16 instructions per tap)
constexpr int XROWS = KS + 2;               // slab rows of X per stage
constexpr int A_BOX = KS * 128;             // one [32 x 64] box of dY
constexpr int B_BOX = 5 * 1024;             // one [34 x 64] box of X, padded to whole 1024-byte swizzle atoms
constexpr int STAGE_BYTES = 2 * A_BOX + 2 * B_BOX;
constexpr int STAGES = 8;
constexpr int STG_BYTES = 128 * 32 * 4;     // epilogue staging: [128 co][32 ci] fp32
constexpr uint32_t TX_BYTES = 2 * A_BOX + 2 * XROWS * 128;

struct WgradParams {
    int rows, wp;                           // rows of the PM tensors, padded width
    int stages_per_chunk, n_chunks;
};

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)(1024u >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, const void *smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(256, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
             const __grid_constant__ CUtensorMap map_dw, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *stg = smem + STAGES * STAGE_BYTES;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(stg + STG_BYTES);
    uint64_t *bar_empty = bar_full + STAGES;
    uint64_t *bar_done = bar_empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // unit = (group, filter row, K-chunk)
    const int chunk = blockIdx.x % p.n_chunks;
    const int dy = (blockIdx.x / p.n_chunks) % 3;
    const int g = blockIdx.x / (3 * p.n_chunks);
    const int s0 = chunk * p.stages_per_chunk;
    const int s1 = min(s0 + p.stages_per_chunk, (p.rows + KS - 1) / KS);
    const int n_it = max(s1 - s0, 0);

    if (warp == 0 && lane == 0) { tc::prefetch_tensormap(&map_dy); tc::prefetch_tensormap(&map_x); tc::prefetch_tensormap(&map_dw); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_empty[i], 1); }
        tc::mbar_init(bar_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool leader = tc::elect_one();
        for (int it = 0; it < n_it; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            tc::mbar_wait(&bar_empty[s], ph ^ 1);
            if (leader) {
                uint8_t *a = smem + s * STAGE_BYTES, *b = a + 2 * A_BOX;
                const int row = (s0 + it) * KS;
                const int xrow = row + (dy - 1) * p.wp - 1;           // signed: rows outside the tensor arrive as zeros
                tc::mbar_arrive_expect_tx(&bar_full[s], TX_BYTES);
                tc::tma_load_2d(a, &map_dy, &bar_full[s], g * 128, row);
                tc::tma_load_2d(a + A_BOX, &map_dy, &bar_full[s], g * 128 + 64, row);
                tc::tma_load_2d(b, &map_x, &bar_full[s], g * 128, xrow);
                tc::tma_load_2d(b + B_BOX, &map_x, &bar_full[s], g * 128 + 64, xrow);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = tc::elect_one();
        // kind::f16, bf16 x bf16 -> fp32, M = N = 128, both operands MN-major (bits 15 / 16)
        const uint32_t idesc = tc::idesc_bf16_f32(128, 128) | (1u << 15) | (1u << 16);
        for (int it = 0; it < n_it; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            tc::mbar_wait(&bar_full[s], ph);
            tc::fence_after_thread_sync();
            if (leader) {
                const uint32_t a = tc::smem_u32(smem + s * STAGE_BYTES), b = a + 2 * A_BOX;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int ks = 0; ks < KS / 16; ++ks)
                        tc::umma_bf16(tmem + dx * 128, desc_mn_sw128(a + ks * 2048, A_BOX), desc_mn_sw128(b + dx * 128 + ks * 2048, B_BOX),
                                      idesc, (it > 0 || ks > 0) ? 1u : 0u);
                tc::umma_commit(&bar_empty[s]);                       // the stage is free once these MMAs have read it
                if (it == n_it - 1) tc::umma_commit(bar_done);
            }
            __syncwarp();
        }
    } else if (warp >= 4 && n_it > 0) {
        // ===================== epilogue: TMEM -> shared -> TMA reduce-add into dW =====================
        const int ew = warp - 4;                                       // TMEM lanes 32*ew .. 32*ew + 31 = co
        tc::mbar_wait(bar_done, 0);
        tc::fence_after_thread_sync();
        float *row = reinterpret_cast<float *>(stg) + (ew * 32 + lane) * 32;
        for (int dx = 0; dx < 3; ++dx) {
            for (int c = 0; c < 128; c += 32) {
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem + ((uint32_t)(ew * 32) << 16) + dx * 128 + c, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(row + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");         // the four epilogue warps
                if (warp == 4 && lane == 0) {
                    tma_reduce_add_2d(&map_dw, stg, c, (g * 9 + dy * 3 + dx) * 128);
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

// ---- host ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(sym);
}
static bool make_map(CUtensorMap *m, CUtensorMapDataType dt, int esize, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                     uint32_t box_cols, CUtensorMapSwizzle sw) {
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * (uint64_t)esize};
    cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    return encode_fn()(m, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int run(int n_img, int h, int w, int groups, bool check, int reps, int force_chunks = 0) {
    const int hp = h + 2, wp = w + 2, rows = n_img * hp * wp, C = groups * 128;
    std::vector<__nv_bfloat16> x((size_t)rows * C), dy((size_t)rows * C);
    std::vector<float> xf(x.size()), dyf(dy.size());
    srand(7);
    for (int r = 0; r < rows; ++r) {
        const int py = (r / wp) % hp, px = r % wp;
        const bool border = py == 0 || py == hp - 1 || px == 0 || px == wp - 1;
        for (int c = 0; c < C; ++c) {
            const float a = border ? 0.f : (float)(rand() % 7 - 3), b = border ? 0.f : (float)(rand() % 5 - 2);
            x[(size_t)r * C + c] = __float2bfloat16(a); xf[(size_t)r * C + c] = a;
            dy[(size_t)r * C + c] = __float2bfloat16(b); dyf[(size_t)r * C + c] = b;
        }
    }
    __nv_bfloat16 *d_x, *d_dy; float *d_dw;
    const size_t dw_elems = (size_t)groups * 9 * 128 * 128;
    cudaMalloc(&d_x, x.size() * 2); cudaMalloc(&d_dy, dy.size() * 2); cudaMalloc(&d_dw, dw_elems * 4);
    cudaMemcpy(d_x, x.data(), x.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d_dy, dy.data(), dy.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap m_dy, m_x, m_dw;
    if (!make_map(&m_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_dy, rows, C, KS, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_map(&m_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_x, rows, C, XROWS, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_map(&m_dw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d_dw, (uint64_t)groups * 9 * 128, 128, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE)) {
        printf("cuTensorMapEncodeTiled failed\n"); return 1;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    WgradParams p;
    p.rows = rows; p.wp = wp;
    const int total_stages = (rows + KS - 1) / KS;
    p.n_chunks = std::max(1, std::min(total_stages, sms / (groups * 3)));
    if (force_chunks > 0) p.n_chunks = std::min(force_chunks, total_stages);
    p.stages_per_chunk = (total_stages + p.n_chunks - 1) / p.n_chunks;
    const size_t smem = STAGES * STAGE_BYTES + STG_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
    cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = groups * 3 * p.n_chunks;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0.f;
    for (int rep = 0; rep < reps; ++rep) {
        cudaMemset(d_dw, 0, dw_elems * 4);
        cudaEventRecord(e0);
        wgrad_kernel<<<grid, 256, smem>>>(m_dy, m_x, m_dw, p);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double flop = 2.0 * rows * 128.0 * 128.0 * 9 * groups;
    printf("n=%d %dx%d groups=%d: grid %d (%d chunks x %d stages), %.1f us, %.0f TFLOP/s\n", n_img, h, w, groups, grid, p.n_chunks,
           p.stages_per_chunk, ms * 1e3, flop / (ms * 1e-3) / 1e12);
    if (check) {
        std::vector<float> got(dw_elems);
        cudaMemcpy(got.data(), d_dw, dw_elems * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int g = 0; g < groups; ++g)
            for (int t = 0; t < 9; ++t) {
                const int off = (t / 3 - 1) * wp + (t % 3 - 1);
                for (int co = 0; co < 128; ++co)
                    for (int ci = 0; ci < 128; ++ci) {
                        double acc = 0;
                        for (int r = 0; r < rows; ++r) {
                            const int rx = r + off;
                            if (rx < 0 || rx >= rows) continue;
                            acc += (double)dyf[(size_t)r * C + g * 128 + co] * xf[(size_t)rx * C + g * 128 + ci];
                        }
                        worst = fmax(worst, fabs(acc - got[(((size_t)g * 9 + t) * 128 + co) * 128 + ci]));
                    }
            }
        printf("  max |dW - host| = %g %s\n", worst, worst == 0 ? "(exact: small integers)" : "<- WRONG");
    }
    cudaFree(d_x); cudaFree(d_dy); cudaFree(d_dw);
    return 0;
}

int main() {
    if (run(2, 6, 6, 2, true, 1)) return 1;          // 128 rows = 4 stages, one per chunk: no accumulation over stages yet
    if (run(2, 6, 6, 2, true, 1, 1)) return 1;       // the same in ONE chunk: accumulation over 4 stages
    if (run(3, 10, 7, 1, true, 1, 1)) return 1;      // 324 rows (not a multiple of 32) = 11 stages in one chunk: the ring of 8 wraps
    if (run(3, 10, 7, 1, true, 1, 3)) return 1;      // three chunks adding into the same tiles
    return run(32, 38, 38, 4, false, 5);             // configs[1] source 1: forward of the same conv takes 54-60 us
}
