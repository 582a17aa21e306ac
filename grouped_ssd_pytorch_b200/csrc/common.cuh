// common.cuh — shared device helpers for libgssd_b200.so (sm_100a).
//
// Arithmetic contract: every float operation that feeds a discrete decision (IoU, thresholds, mining
// keys, decoded boxes) is an IEEE round-to-nearest add/sub/mul/div in the reference's association
// order.  The library is compiled with -fmad=false and the hot expressions additionally use the
// __f*_rn intrinsics, which ptxas never contracts into FMA.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/gssd.h"

namespace cg = cooperative_groups;

namespace gssd {

constexpr unsigned FULL = 0xffffffffu;

// host-side launch accounting (gssd_launch_count)
void note_launch(int n = 1);

#define GSSD_RETURN_IF_CUDA(expr)                          \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return (int)_e;             \
    } while (0)

#define GSSD_AFTER_LAUNCH()                                \
    do {                                                   \
        gssd::note_launch();                               \
        cudaError_t _e = cudaPeekAtLastError();            \
        if (_e != cudaSuccess) return (int)_e;             \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Opt a kernel in to the full 227 KB of shared memory (static + dynamic) once per kernel and device.
// (Keyed on the function address: all instantiations of one kernel template share a pointer TYPE.)
static inline cudaError_t allow_max_smem(const void *kern) {
    struct Entry { const void *fn; unsigned dev_mask; };
    static Entry table[64];
    static int n_entries = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    Entry *ent = nullptr;
    for (int i = 0; i < n_entries; ++i) if (table[i].fn == kern) { ent = &table[i]; break; }
    if (ent && dev < 32 && ((ent->dev_mask >> dev) & 1u)) return cudaSuccess;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    int optin = 0;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return e;
    if (!ent && n_entries < 64) { ent = &table[n_entries++]; ent->fn = kern; ent->dev_mask = 0; }
    if (ent && dev < 32) ent->dev_mask |= 1u << dev;
    return cudaSuccess;
}

// Cluster size (CTAs per image, 1/2/4/8) for the per-image kernels.  Measured on B200 (tools/quickperf.py with every size
// forced, profiles/r1_cluster_sizes.txt): while the batch alone cannot fill the machine (8 CTAs per image still fit 4 CTAs per
// SM) the widest cluster wins by a factor of two; beyond that the cluster barriers cost more than the parallelism returns and
// the size only has to keep enough warps resident — the loss kernel's per-image key array (4 bytes per prior of the slice)
// limits the CTAs per SM at SSD512 prior counts, so it keeps 4 CTAs per image there.
enum { GSSD_KERNEL_MATCH = 0, GSSD_KERNEL_LOSS = 1 };
static inline int pick_cluster_size(int kernel, int B, int P) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int S;
    if ((long)B * 8 <= 4l * sms) S = 8;                            // B <= 74 on B200
    else if (kernel == GSSD_KERNEL_LOSS) S = P >= 16384 ? 4 : (B <= 512 ? 2 : 1);
    else S = 2;
    while (S > 1 && P / S < 512) S >>= 1;                          // tiny prior sets: nothing to split
    return S;
}

static inline int resident_ctas(const void *kern, int threads, size_t smem) {
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
        per_sm = 1;
        cudaGetLastError();
    }
    return sms * per_sm;
}

// ---- peer exchange of the loss statistics (gssd_xchg, include/gssd.h) ---------------------------------------
// One 64-bit word per (value, step parity, source rank): (epoch << 32) | value.  Value and tag travel in ONE naturally
// aligned 8-byte store, so a reader that sees the epoch of the current step also sees the value — no fence between them.
constexpr int GSSD_FUSED_MAX_CTAS = 1024;    // CTAs of one rank's one-launch loss kernel (fused.cu) that have a slot
struct XBuf {
    unsigned long long xmax[2][GSSD_XCHG_MAX_RANKS];   // max of conf (ordered uint32) of every rank   (two-launch path)
    unsigned long long npos[2][GSSD_XCHG_MAX_RANKS];   // number of positives of every rank           (two-launch path)
    uint32_t epoch;                          // completed steps (advanced by the LAST CTA of a step's last kernel)
    uint32_t match_done;                     // CTA counter of the running stage-1 kernel (two-launch path)
    // one-launch path: every CTA of every rank owns a word per value; hdr = how many CTAs that rank runs this step
    unsigned long long hdr[2][GSSD_XCHG_MAX_RANKS];
    unsigned long long slot[2][2][GSSD_XCHG_MAX_RANKS][GSSD_FUSED_MAX_CTAS];   // [parity][0: conf max, 1: positives][rank][cta]
};
struct XDev {                                // kernel-argument image of gssd_xchg
    XBuf *peers[GSSD_XCHG_MAX_RANKS];
    int rank, world;                         // world == 0: exchange disabled
    unsigned long long timeout_ns;           // a wait for a peer's word that lasts longer traps (0 = wait for ever)
};
static inline XDev xdev_from(const gssd_xchg *x) {
    XDev d = {};
    if (x != nullptr) {
        for (int r = 0; r < GSSD_XCHG_MAX_RANKS; ++r) d.peers[r] = reinterpret_cast<XBuf *>(x->peers[r]);
        d.rank = x->rank; d.world = x->world;
        d.timeout_ns = (unsigned long long)x->timeout_ms * 1000000ull;
    }
    return d;
}

// rendezvous state of the one-launch MultiBoxLoss (fused.cu) on ONE GPU: zero-initialised once.  Every CTA owns a word per
// value, (epoch << 32) | value, written with one plain store and polled by the readers: no atomics, no fences, nothing to
// reset (the epoch, advanced by the last CTA to leave, tells the launches apart).
struct FusedState {
    uint32_t epoch; uint32_t pad[3];
    unsigned long long slot[4][GSSD_FUSED_MAX_CTAS];    // [0: conf max, 1: positives, 2: loss_l partial, 3: loss_c partial][cta]
};

// ---- optional phase timing (debug build only: -DGSSD_PHASE_TIMING, see tools/phase_times.py) ----------
#ifdef GSSD_PHASE_TIMING
// one array + accessor per translation unit (no relocatable device code in this build)
#define GSSD_PHASE_DECL(name)                                                                              \
    namespace gssd { __device__ long long g_phase_clock_##name[32]; }                                      \
    extern "C" __attribute__((visibility("default"))) int gssd_debug_phase_clocks_##name(long long *out) { \
        return (int)cudaMemcpyFromSymbol(out, gssd::g_phase_clock_##name, sizeof(long long) * 32);         \
    }
#define GSSD_PHASE(name, i, cond)                                                       \
    do { if ((cond) && threadIdx.x == 0) g_phase_clock_##name[i] = clock64(); } while (0)
#else
#define GSSD_PHASE_DECL(name)
#define GSSD_PHASE(name, i, cond) do { } while (0)
#endif

// ---- order-preserving float <-> uint32 (larger float <=> larger unsigned) -----------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// ---- box arithmetic -------------------------------------------------------------------------------
// point_form, box_utils.py:4-13 : c -/+ wh/2 (the /2 is exact)
__device__ __forceinline__ float4 point_form(float4 p) {
    float hw = __fmul_rn(p.z, 0.5f), hh = __fmul_rn(p.w, 0.5f);
    return make_float4(__fsub_rn(p.x, hw), __fsub_rn(p.y, hh), __fadd_rn(p.x, hw), __fadd_rn(p.y, hh));
}
__device__ __forceinline__ float box_area(float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
// intersect, box_utils.py:28-46
__device__ __forceinline__ float box_inter(float4 a, float4 b) {
    float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    w = w < 0.f ? 0.f : w;          // torch.clamp(min=0)
    h = h < 0.f ? 0.f : h;
    return __fmul_rn(w, h);
}
// jaccard, box_utils.py:49-67 : inter / ((area_a + area_b) - inter)
__device__ __forceinline__ float box_iou_exact(float4 a, float area_a, float4 b, float area_b) {
    float inter = box_inter(a, b);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}
// Same value whenever the union is positive (always, for priors of positive area): the IEEE divide is
// skipped for the disjoint pairs, which are the vast majority of (GT, prior) pairs.
__device__ __forceinline__ float box_iou_fast(float4 a, float area_a, float4 b, float area_b) {
    float inter = box_inter(a, b);
    if (!(inter > 0.f)) return 0.f;
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}
// encode, box_utils.py:114-135 (true divisions, as the reference's CPU path)
__device__ __forceinline__ float4 encode_box(float4 m, float4 p, float v0, float v1) {
    float4 o;
    o.x = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(m.x, m.z), 0.5f), p.x), __fmul_rn(v0, p.z));
    o.y = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(m.y, m.w), 0.5f), p.y), __fmul_rn(v0, p.w));
    o.z = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(m.z, m.x), p.z)), v1);
    o.w = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(m.w, m.y), p.w)), v1);
    return o;
}
// decode, box_utils.py:139-157
__device__ __forceinline__ float4 decode_box(float4 l, float4 p, float v0, float v1) {
    float cx = __fadd_rn(p.x, __fmul_rn(__fmul_rn(l.x, v0), p.z));
    float cy = __fadd_rn(p.y, __fmul_rn(__fmul_rn(l.y, v0), p.w));
    float w = __fmul_rn(p.z, expf(__fmul_rn(l.z, v1)));
    float h = __fmul_rn(p.w, expf(__fmul_rn(l.w, v1)));
    cx = __fsub_rn(cx, __fmul_rn(w, 0.5f));
    cy = __fsub_rn(cy, __fmul_rn(h, 0.5f));
    return make_float4(cx, cy, __fadd_rn(w, cx), __fadd_rn(h, cy));
}

// ---- warp / block reductions ---------------------------------------------------------------------
// all 32 lanes must be converged; one CREDUX instruction each on sm_100a (float min / max: redux.sync.f32, new with Blackwell)
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(FULL, v); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float warp_min(float v) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// split cluster barrier: arrive early (e.g. once the buffers other CTAs will write into are initialised), wait right before
// the first access to another CTA's shared memory — which also guarantees that every CTA of the cluster is running
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ float4 ldg_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ float2 ldg_stream(const float2 *p) { return __ldcs(p); }
__device__ __forceinline__ float ldg_stream(const float *p) { return __ldcs(p); }

}  // namespace gssd
