// umma_mnmajor_probe.cu — probe for the wgrad kernel planned in DESIGN.md §7: does tcgen05.mma (kind::f16, bf16 -> fp32) take
// BOTH operands MN-major (a_major = b_major = 1) from shared memory laid out the way a SWIZZLE_128B TMA box of a
// pixel-major tensor lands there?
//
//   D[m][n] = sum_k A[k][m] * B[k][n]        A: [K][M] bf16, m contiguous   (dY: [pixel][co])
//                                            B: [K][N] bf16, n contiguous   (X : [pixel][ci])
//
// Shared-memory image of an operand with MN = 128, K = 64 (what two TMA boxes of [64 rows x 64 channels] produce):
//   byte(k, mn) = (mn / 64) * LBO + (k / 8) * SBO + (k % 8) * 128 + (((mn % 64) / 8) ^ (k % 8)) * 16 + (mn % 8) * 2
//   with SBO = 1024 (8 rows of 128 bytes = one swizzle atom) and LBO = (K / 8) * 1024 (one whole box).
// One K = 16 instruction covers two atoms along K: descriptor start address + 2048 bytes per step.
// The probe tries the (LBO, SBO) assignment above and the swapped one (variant 0 is exact on a B200, profiles/r1_mnmajor_probe.txt)
// and, for the wgrad kernel's taps, B descriptors that start 1 / 2 rows into the box; it prints the max |error| against the host:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_mnmajor_probe tools/umma_mnmajor_probe.cu && tools/umma_mnmajor_probe
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../grouped_ssd_pytorch_b200/csrc/tc.cuh"

using namespace gssd;

constexpr int M = 128, N = 128, K = 64;
constexpr uint32_t ATOM = 1024, BOX = (K / 8) * ATOM;      // one [K x 64] box
constexpr int KB = K + 8;                                   // rows of B kept in shared memory (row-offset variants)
constexpr uint32_t BOXB = (KB / 8) * ATOM;

__device__ __forceinline__ uint32_t mn_major_offset(int k, int mn, uint32_t box = BOX) {
    return (uint32_t)(mn / 64) * box + (uint32_t)(k / 8) * ATOM + (uint32_t)(k % 8) * 128 +
           (uint32_t)((((mn % 64) / 8) ^ (k % 8)) * 16) + (uint32_t)(mn % 8) * 2;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
           (1ull << 46) | (2ull << 61);
}

// variant 0: LBO = box (next 64 of MN), SBO = atom (next 8 of K);  variant 1: the two swapped;
// variants 2 / 3: as 0, with the B descriptor starting 1 / 2 rows (128 / 256 bytes) into its box — what the three taps of a
// filter row do with one slab of X in tools/wgrad_probe.cu:  D = A^T B[r : r + K]
__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int variant) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sa = smem, *sb = smem + 2 * BOX;
    const int row_off = variant >= 2 ? variant - 1 : 0;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < K * M; i += blockDim.x) {
        const int k = i / M, mn = i % M;
        *reinterpret_cast<__nv_bfloat16 *>(sa + mn_major_offset(k, mn)) = A[i];
    }
    for (int i = threadIdx.x; i < KB * N; i += blockDim.x) {      // B has KB rows in global memory too
        const int k = i / N, mn = i % N;
        *reinterpret_cast<__nv_bfloat16 *>(sb + mn_major_offset(k, mn, BOXB)) = B[i];
    }
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic writes -> reads by the tensor core
    if (threadIdx.x < 32) tc::tmem_alloc(&tmem_slot, 128);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        // instruction descriptor of tc::idesc_bf16_f32 plus a_major (bit 15) and b_major (bit 16) = MN-major
        const uint32_t idesc = tc::idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);
        const bool swapped = variant == 1;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t ad = desc_sw128(tc::smem_u32(sa) + ks * 2 * ATOM, swapped ? ATOM : BOX, swapped ? BOX : ATOM);
            const uint64_t bd = desc_sw128(tc::smem_u32(sb) + row_off * 128 + ks * 2 * ATOM, swapped ? ATOM : BOXB, swapped ? BOXB : ATOM);
            tc::umma_bf16(tmem, ad, bd, idesc, ks > 0);
        }
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_thread_sync();
    // 4 warps x 32 lanes = the 128 rows of D; 4 x 32 columns each
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < N; c += 32) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
        tc::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (threadIdx.x < 32) { tc::fence_after_thread_sync(); tc::tmem_dealloc(tmem, 128); }
}

int main() {
    std::vector<__nv_bfloat16> a(K * M), b(KB * N);
    std::vector<float> af(K * M), bf(KB * N), ref(M * N, 0.f), got(M * N);
    srand(1);
    for (int i = 0; i < K * M; ++i) { a[i] = __float2bfloat16((rand() % 17 - 8) * 0.125f); af[i] = __bfloat162float(a[i]); }
    for (int i = 0; i < KB * N; ++i) { b[i] = __float2bfloat16((rand() % 13 - 6) * 0.25f); bf[i] = __bfloat162float(b[i]); }
    __nv_bfloat16 *da, *db; float *dd;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, got.size() * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = 2 * BOX + 2 * BOXB + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char *names[4] = {"LBO = box, SBO = atom", "LBO = atom, SBO = box", "as 0, B starts 1 row into its box", "as 0, B starts 2 rows into its box"};
    for (int variant = 0; variant < 4; ++variant) {
        const int r0 = variant >= 2 ? variant - 1 : 0;
        std::fill(ref.begin(), ref.end(), 0.f);
        for (int k = 0; k < K; ++k)
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) ref[m * N + n] += af[k * M + m] * bf[(k + r0) * N + n];
        cudaMemset(dd, 0, got.size() * 4);
        probe_kernel<<<1, 128, smem>>>(da, db, dd, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(got.data(), dd, got.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int i = 0; i < M * N; ++i) worst = fmax(worst, fabs((double)got[i] - ref[i]));
        printf("variant %d (%s): max |D - ref| = %g %s\n", variant, names[variant], worst,
               worst == 0 ? "<- exact (small integers / 8: every product is exact in fp32)" : "");
    }
    return 0;
}
