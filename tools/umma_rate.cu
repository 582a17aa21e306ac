// umma_rate.cu — micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 -> fp32) issued back to back by one
// thread, for cta_group::1 (M=128) and cta_group::2 (M=256 over a CTA pair), N in {64,128,256}, with the A operand
// descriptor starting at an aligned row or at an odd row offset inside a 128B-swizzled slab.
// Development aid behind DESIGN.md's tile-shape choices:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../grouped_ssd_pytorch_b200/csrc/tc.cuh"

using namespace gssd;

template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int reps, int a_row_off, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (threadIdx.x < 32) {
        if (CG == 1) {
            tc::tmem_alloc(&tmem_slot, 512);
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    const uint32_t tmem = tmem_slot;
    long long dt = 0;
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = tc::idesc_bf16_f32(128 * CG, n);
        const uint64_t adesc = tc::smem_desc_k128(tc::smem_u32(smem) + a_row_off * 128);
        const uint64_t bdesc = tc::smem_desc_k128(tc::smem_u32(smem) + 64 * 1024);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tmem + (r & 1) * 256;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (CG == 1) {
                    tc::umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
                } else {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(1) : "memory");
                }
            }
        }
        if (CG == 1) {
            tc::umma_commit(&bar);
        } else {
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(tc::smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        }
        tc::mbar_wait(&bar, 0);
        dt = clock64() - t0;
        out[blockIdx.x / CG] = dt;
    } else if (CG == 2 && threadIdx.x == 0) {
        tc::mbar_wait(&bar, 0);                       // the multicast commit arrives here too
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (threadIdx.x < 32) {
        tc::fence_after_thread_sync();
        if (CG == 1) tc::tmem_dealloc(tmem, 512);
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// mode bits: 1 = tcgen05.commit to a scratch barrier after every group of 8 MMAs, 2 = mbarrier.try_wait on an already
// completed barrier before every group, 4 = tcgen05.fence::after_thread_sync per group, 8 = whole warp walks the loop
__global__ void __launch_bounds__(128, 1) loop_kernel(int n, int groups, int mode, long long *out, int vary, int taps_per_group) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, scratch[8], ready;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar, 1); tc::mbar_init(&ready, 1);
        for (int i = 0; i < 8; ++i) tc::mbar_init(&scratch[i], 1);
        tc::fence_barrier_init();
        tc::mbar_arrive(&ready);                       // phase 0 of `ready` is complete: try_wait(parity 0) succeeds at once
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (threadIdx.x < 32) {
        const bool whole = (mode & 8) != 0;
        const bool leader = tc::elect_one();
        if (whole || leader) {
            const uint32_t idesc = tc::idesc_bf16_f32(128, n);
            const uint32_t a0 = tc::smem_u32(smem), b0 = tc::smem_u32(smem) + 64 * 1024;
            const long long t0 = clock64();
            for (int g = 0; g < groups; ++g) {
                if (mode & 2) tc::mbar_wait(&ready, 0);
                if (mode & 4) tc::fence_after_thread_sync();
                if (leader) {
                    for (int t = 0; t < taps_per_group; ++t) {
                        const uint64_t adesc = tc::smem_desc_k128(a0 + ((vary & 1) ? ((g + t) % 9) * 128 : 0));
                        const uint64_t bdesc = tc::smem_desc_k128(b0 + ((vary & 2) ? ((g + t) & 1) * 16384 : 0));
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc::umma_bf16(tmem + ((vary & 4) ? sub * 128 : 0), adesc + 2 * k + ((vary & 8) ? sub * 1024 : 0), bdesc + 2 * k, idesc, 1);
                    }
                    if (mode & 1) tc::umma_commit(&scratch[g & 7]);
                }
                if (whole) __syncwarp();
            }
            if (leader) {
                tc::umma_commit(&bar);
                tc::mbar_wait(&bar, 0);
                out[blockIdx.x] = clock64() - t0;
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (threadIdx.x < 32) { tc::fence_after_thread_sync(); tc::tmem_dealloc(tmem, 512); }
}

static void run_loop(int n, int mode, int vary = 15, int tpg = 1) {
    long long *out;
    cudaMalloc(&out, sizeof(long long) * 256);
    const int groups = 256, smem = 97 * 1024 + 1024, grid = 74;
    cudaFuncSetAttribute(loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    loop_kernel<<<grid, 128, smem>>>(n, groups, mode, out, vary, tpg);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i];
    avg /= grid;
    printf("loop N=%3d mode=%2d (commit=%d try_wait=%d fence=%d whole_warp=%d): %s  vary=%2d taps/group=%d: %.1f cycles per 8 MMAs (nominal %.0f)\n", n, mode,
           mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, cudaGetErrorString(e), vary, tpg, avg / groups / tpg, 8.0 * 128 * n * 16 / 4096.0);
    cudaFree(out);
}

template <int CG>
static void run(int n, int a_off, int grid) {
    long long *out;
    cudaMalloc(&out, sizeof(long long) * 256);
    cudaMemset(out, 0, sizeof(long long) * 256);
    const int reps = 512, smem = 97 * 1024 + 1024;
    cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid * CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG>, n, reps, a_off, out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i];
    avg /= grid;
    const double per = avg / (reps * 4), macs = 128.0 * CG * n * 16;
    printf("cta_group::%d M=%3d N=%3d a_row_off=%2d grid=%3d : %s  %.1f cycles/MMA  %.0f MAC/cycle/SM  (nominal %.0f cycles)\n", CG, 128 * CG, n, a_off,
           grid, cudaGetErrorString(e), per, macs / per / CG, macs / CG / 4096.0);
    cudaFree(out);
}

int main() {
    for (int vary : {0, 1, 2, 4, 8, 12, 15}) run_loop(128, 0, vary, 1);
    for (int tpg : {1, 3, 9}) run_loop(128, 15, 15, tpg);
    for (int tpg : {1, 3, 9}) run_loop(128, 15, 7, tpg);
    return 0;
    for (int grid : {1, 74}) {
        for (int n : {64, 128, 256}) {
            run<1>(n, 0, grid);
            run<1>(n, 41, grid);
        }
        for (int n : {64, 128, 256}) {
            run<2>(n, 0, grid);
            run<2>(n, 41, grid);
        }
    }
    return 0;
}
