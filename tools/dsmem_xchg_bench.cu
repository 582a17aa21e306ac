// dsmem_xchg_bench.cu — micro-benchmark: what one all-to-all exchange of a 256-bin histogram between the CTAs of a
// thread-block cluster costs, per round, with
//   A  the scheme of csrc/select.cuh today: remote atomicAdd of the non-empty bins into every CTA's totals + cluster.sync()
//      (barrier.cluster.arrive.release = MEMBAR.ALL.GPU ... UCGABAR_ARV, wait = UCGABAR_WAIT ; CCTL.IVALL);
//   B  the same pushes + a RELAXED arrive / wait pair (no MEMBAR) — timing only: without the release the pushes are not ordered
//      against the barrier, so the sums of this variant are NOT checked;
//   C  bulk copies: every CTA sends its whole 1 KB histogram to a per-sender row of every peer with
//      cp.async.bulk.shared::cluster.shared::cta, completion (complete_tx) on an mbarrier of the RECEIVING CTA; the
//      receiver waits on its own mbarrier and adds up the rows — no cluster barrier, no fence, no L1 invalidate.
// Rows and mbarriers are double-buffered by round parity; a CTA can only be one round ahead of a peer (it needs the
// peer's data of round r to finish round r), so two buffers are enough (same argument as the two histogram buffers of
// radix_select).  Every variant checks its sums (except B) and reports cycles per round as seen by CTA 0.
// Development aid for DESIGN.md §7 (batch-32 latency of the loss kernel):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_xchg_bench tools/dsmem_xchg_bench.cu && tools/dsmem_xchg_bench
#include <cstdio>
#include <cstdint>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include "../grouped_ssd_pytorch_b200/csrc/tc.cuh"

namespace cg = cooperative_groups;
using namespace gssd;

constexpr int NT = 256;
constexpr int BINS = 256;
constexpr int MAX_S = 8;

__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
// 16-byte multiples, both addresses 16-byte aligned; completes `bytes` on the mbarrier at remote_bar (shared::cluster address)
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar) : "memory");
}

// the bins a CTA contributes in round `round`: deterministic, a few of them empty
__device__ __forceinline__ uint32_t contribution(uint32_t rank, int bin, int round) {
    const uint32_t v = (rank * 131u + (uint32_t)bin * 7u + (uint32_t)round * 13u) % 11u;
    return v < 3 ? 0u : v;
}

struct Shared {
    alignas(16) uint32_t hist[2][BINS];            // what this CTA sends (C: source of the bulk copies)
    alignas(16) uint32_t total[2][BINS];           // A / B: added to by every CTA of the cluster
    alignas(16) uint32_t rows[2][MAX_S][BINS];     // C: one row per sender
    alignas(8) uint64_t bar[2];                    // C: completion of a round's incoming rows
};

template <int MODE>   // 0 = A, 1 = B, 2 = C
__global__ void __launch_bounds__(NT, 1) xchg_kernel(int rounds, long long *cycles, int *errors) {
    __shared__ Shared sh;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned S = cluster.num_blocks(), rank = cluster.block_rank();
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * BINS; i += NT) (&sh.total[0][0])[i] = 0;
    if (MODE == 2 && tid == 0) { tc::mbar_init(&sh.bar[0], 1); tc::mbar_init(&sh.bar[1], 1); tc::fence_barrier_init(); }
    __syncthreads();
    cluster.sync();                                   // everybody runs, buffers cleared, mbarriers initialised
    int bad = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const int b = r & 1;
        for (int i = tid; i < BINS; i += NT) {
            sh.hist[b][i] = contribution(rank, i, r);
            if (MODE != 2) sh.total[b ^ 1][i] = 0;        // the next round adds into it, after this round's barrier
        }
        if (MODE == 2) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes of hist -> async-proxy reads
            __syncthreads();
            if (tid == 0) tc::mbar_arrive_expect_tx(&sh.bar[b], (S - 1) * BINS * 4);
            if (tid < (int)S && tid != (int)rank) {       // one thread per peer
                const uint32_t dst = mapa(tc::smem_u32(&sh.rows[b][rank][0]), tid);
                const uint32_t bar = mapa(tc::smem_u32(&sh.bar[b]), tid);
                bulk_copy_to_peer(dst, tc::smem_u32(&sh.hist[b][0]), BINS * 4, bar);
            }
            tc::mbar_wait(&sh.bar[b], (r >> 1) & 1);
            for (int i = tid; i < BINS; i += NT) {
                uint32_t t = sh.hist[b][i];
                for (unsigned q = 0; q < S; ++q) if (q != rank) t += sh.rows[b][q][i];
                uint32_t want = 0;
                for (unsigned q = 0; q < S; ++q) want += contribution(q, i, r);
                bad += t != want;
            }
            __syncthreads();                              // rows[b] / hist[b] are free again two rounds from now
        } else {
            __syncthreads();
            for (int i = tid; i < BINS; i += NT) {
                const uint32_t c = sh.hist[b][i];
                if (c) for (unsigned q = 0; q < S; ++q) atomicAdd(&cluster.map_shared_rank(&sh.total[b][0], q)[i], c);
            }
            if (MODE == 0) {
                cluster.sync();
            } else {
                asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
            }
            if (MODE == 0) {
                for (int i = tid; i < BINS; i += NT) {
                    uint32_t want = 0;
                    for (unsigned q = 0; q < S; ++q) want += contribution(q, i, r);
                    bad += sh.total[b][i] != want;
                }
            }
            __syncthreads();
        }
    }
    const long long t1 = clock64();
    if (bad) atomicAdd(errors, bad);
    if (blockIdx.x == 0 && tid == 0) *cycles = t1 - t0;
    cluster.sync();                                   // nobody leaves while a peer may still write into it
}

template <int MODE>
static void run(const char *name, int S, int rounds) {
    long long *cyc; int *err;
    cudaMalloc(&cyc, 8); cudaMalloc(&err, 4);
    cudaMemset(cyc, 0, 8); cudaMemset(err, 0, 4);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(S * 16, 1, 1);                 // 16 clusters, as many as a batch-32 launch keeps busy per two SMs
    cfg.blockDim = dim3(NT, 1, 1);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, xchg_kernel<MODE>, rounds, cyc, err);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long c = 0; int bad = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&bad, err, 4, cudaMemcpyDeviceToHost);
    printf("%-46s S=%d  %8.0f cycles/round  %s%s\n", name, S, (double)c / rounds,
           e != cudaSuccess ? cudaGetErrorString(e) : (bad ? "WRONG SUMS" : "ok"), MODE == 1 ? " (sums not checked)" : "");
    cudaFree(cyc); cudaFree(err);
}

int main() {
    const int rounds = 2000;
    for (int S : {2, 4, 8}) {
        run<0>("A  remote atomics + cluster.sync()", S, rounds);
        run<1>("B  remote atomics + relaxed arrive / wait", S, rounds);
        run<2>("C  bulk copies + mbarrier of the receiver", S, rounds);
    }
    return 0;
}
