// select.cuh — per-image radix select in shared memory (replaces the reference's double argsort,
// multibox_loss.py:102-106, and the full sort + slice of nms, box_utils.py:194-196).
//
// The keys of one image live in shared memory as order-preserving uint32 (f2ord), either in one CTA
// or split in contiguous slices across the CTAs of a thread-block cluster; every CTA then PUSHES its
// non-empty histogram bins into the totals of all CTAs of the image (distributed-shared-memory reductions
// that need no answer), so that one cluster barrier per pass is all the waiting there is and nothing is
// read remotely afterwards.  The result is the exact "k largest keys" set with a
// deterministic rule for equal keys (lower index first or higher index first), expressed as one
// 64-bit cut: an element is selected iff  composite(key, index) >= cut,  where
//   composite = key << 32 | (low_first ? ~index : index).
// MSB-first 8-bit digits; as soon as the bin that holds the k-th key has <= 32 members (typically
// after two passes) those members are gathered and ranked by one warp instead of running the
// remaining passes.
#pragma once
#include "common.cuh"

namespace gssd {

struct SelectShared {
    uint32_t hist[2][256];   // this CTA's bins (double-buffered: the last pass is read again when equal keys straddle the cut)
    uint32_t total[2][256];  // cluster: bins of the whole image, added to by every CTA of the cluster (double-buffered per pass)
    uint32_t warp_tmp[32];
    uint32_t digit, k_rem, eq_total, eq_local_before;
    uint32_t fin_count;
    int      tie_idx;
    unsigned long long fin_list[32];
    unsigned long long cut;
};

__device__ __forceinline__ unsigned long long sel_composite(uint32_t key, uint32_t index, bool low_first) {
    return ((unsigned long long)key << 32) | (low_first ? ~index : index);
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    unsigned lo = __shfl_xor_sync(FULL, (unsigned)v, m), hi = __shfl_xor_sync(FULL, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}

// cluster variant of radix_select: call at kernel start, all threads; then cluster_arrive() (any time later) and cluster_wait()
// before radix_select, so that no CTA adds into totals that their owner has not cleared yet
template <int NT>
__device__ __forceinline__ void radix_select_prepare(SelectShared *s) {
    for (int i = threadIdx.x; i < 512; i += NT) (&s->total[0][0])[i] = 0;
    if (threadIdx.x == 0) s->fin_count = 0;
}

// compare-exchange half: keep the larger (keep_max) or the smaller of (mine, other)
__device__ __forceinline__ void cmpx(unsigned long long &mine, unsigned long long other, bool keep_max) {
    if ((other > mine) == keep_max) mine = other;
}

// NT threads per CTA (multiple of 32, >= 256).  keys[0..n_local) are this CTA's slice, whose first
// element has image-wide index `index_base`.  1 <= k <= number of keys in the image.
// CLUSTER: slices are ordered by cluster rank; radix_select_prepare + a cluster barrier must have happened.  Returns the
// cut (may differ between the CTAs of an image, each value classifies that CTA's own elements correctly).
template <int NT, bool CLUSTER>
__device__ unsigned long long radix_select(const uint32_t *keys, int n_local, uint32_t index_base, uint32_t k,
                                           bool low_first, SelectShared *s, long long *dbg_clk = nullptr) {
    // debug build only (tools/phase_times.py): clock stamps of thread 0, four per pass + two for the final ranking
#define SEL_STAMP(i) do { if (dbg_clk && threadIdx.x == 0) dbg_clk[i] = clock64(); } while (0)
    int stamp = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned nranks = CLUSTER ? cluster.num_blocks() : 1;
    const unsigned rank = CLUSTER ? cluster.block_rank() : 0;
    const int n_round = (n_local + NT - 1) / NT * NT;

    if (!CLUSTER && tid == 0) s->fin_count = 0;
    uint32_t prefix = 0, mask = 0, k_rem = k, eq_total = 0;
    int buf = 0;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8, buf ^= 1) {
        for (int i = tid; i < 256; i += NT) {
            s->hist[buf][i] = 0;
            if (CLUSTER) s->total[buf ^ 1][i] = 0;               // the next pass adds into it, after this pass's barrier
        }
        __syncthreads();
        SEL_STAMP(stamp++);
        for (int i = tid; i < n_round; i += NT) {
            bool in = i < n_local;
            const uint32_t key = in ? keys[i] : 0;
            in = in && ((key & mask) == prefix);
            const unsigned act = __ballot_sync(FULL, in);
            if (in) {
                const uint32_t d = (key >> shift) & 255u;
                const unsigned peers = __match_any_sync(act, d);     // one atomic per distinct digit
                if (lane == __ffs(peers) - 1) atomicAdd(&s->hist[buf][d], __popc(peers));
            }
        }
        __syncthreads();
        SEL_STAMP(stamp++);
        const uint32_t *tot = s->hist[buf];
        if (CLUSTER) {
            // push the non-empty bins into every CTA's totals (mine included): reductions without a return value, in flight
            // while the other CTAs of the image are still counting
            for (int i = tid; i < 256; i += NT) {
                const uint32_t c = s->hist[buf][i];
                if (c)
                    for (unsigned r = 0; r < nranks; ++r) atomicAdd(&cluster.map_shared_rank(&s->total[buf][0], r)[i], c);
            }
            cluster.sync();
            tot = s->total[buf];
        }
        SEL_STAMP(stamp++);
        // suffix sums over the 256 bins (bin 255 first): the first 8 warps own one bin per thread
        if (tid < 256) {
            const uint32_t c = tot[tid];
            uint32_t incl = c;                                   // inclusive suffix within the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_down_sync(FULL, incl, o);
                if (lane + o < 32) incl += t;
            }
            if (lane == 0) s->warp_tmp[warp] = incl;             // warp total
            asm volatile("bar.sync 1, 256;");                    // the first 256 threads only
            uint32_t above = 0;
            for (int w = warp + 1; w < 8; ++w) above += s->warp_tmp[w];
            const uint32_t excl = above + incl - c;              // keys in bins > tid
            if (excl < k_rem && k_rem <= excl + c) {
                s->digit = tid; s->k_rem = k_rem - excl; s->eq_total = c;
            }
        }
        __syncthreads();
        SEL_STAMP(stamp++);
        prefix |= s->digit << shift;
        mask |= 255u << shift;
        k_rem = s->k_rem;
        eq_total = s->eq_total;
        if (eq_total <= 32u) break;                              // few enough to rank directly
    }

    if (eq_total <= 32u) {
        // ---- gather the members of the bin and rank them with one warp; in a cluster every CTA receives every member
        // (<= 32 of them), so that nothing is read remotely after the barrier and no CTA has to outlive another ---------
        for (int i = tid; i < n_local; i += NT) {
            const uint32_t key = keys[i];
            if ((key & mask) == prefix) {
                const unsigned long long comp = sel_composite(key, index_base + (uint32_t)i, low_first);
                if (CLUSTER) {
                    for (unsigned r = 0; r < nranks; ++r) {
                        SelectShared *sr = cluster.map_shared_rank(s, r);
                        sr->fin_list[atomicAdd(&sr->fin_count, 1u)] = comp;
                    }
                } else {
                    s->fin_list[atomicAdd(&s->fin_count, 1u)] = comp;
                }
            }
        }
        if (CLUSTER) cluster.sync(); else __syncthreads();
        SEL_STAMP(stamp++);
        if (warp == 0) {
            unsigned long long c = lane < (int)eq_total ? s->fin_list[lane] : 0ull;
#pragma unroll
            for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
                for (int st = size >> 1; st >= 1; st >>= 1) {
                    const bool lower = (lane & st) == 0, desc = (lane & size) == 0;
                    cmpx(c, shfl_xor_u64(c, st), desc == lower);
                }
            }
            c = ((unsigned long long)__shfl_sync(FULL, (unsigned)(c >> 32), k_rem - 1) << 32) |
                __shfl_sync(FULL, (unsigned)c, k_rem - 1);
            if (lane == 0) s->cut = c;
        }
        __syncthreads();
        SEL_STAMP(stamp++);
        return s->cut;
    }

    // ---- all 32 bits resolved and more than 32 keys are exactly equal to v = prefix -----------------------
    buf ^= 1;                                // buffer of the last pass: hist[buf][digit] = local count of keys == v
    const uint32_t v = prefix, need = k_rem;
    const unsigned long long take_all = (unsigned long long)v << 32;
    const unsigned long long take_none = ((unsigned long long)v << 32) + 0x100000000ull;
    if (need == eq_total) {
        if (CLUSTER) cluster.sync();
        return take_all;
    }
    const uint32_t eq_local = s->hist[buf][s->digit];
    uint32_t before = 0;                     // equal keys in slices of lower rank
    if (CLUSTER) {
        if (tid == 0) {
            uint32_t t = 0;
            for (unsigned q = 0; q < rank; ++q) t += cluster.map_shared_rank(&s->hist[buf][0], q)[s->digit];
            s->eq_local_before = t;
        }
        __syncthreads();
        before = s->eq_local_before;
        cluster.sync();
    }
    // number of local equal keys to take, counted from the preferred end
    long long want;
    if (low_first) want = (long long)need - before;
    else           want = (long long)need - ((long long)eq_total - before - eq_local);
    if (want <= 0) return take_none;
    if (want >= (long long)eq_local) return take_all;
    // local index of the want-th equal key from the preferred end (block scan in visiting order)
    if (tid == 0) s->tie_idx = 0;
    __syncthreads();
    uint32_t running = 0;
    const int n_chunks = (n_local + NT - 1) / NT;
    for (int c = 0; c < n_chunks; ++c) {
        const int chunk = low_first ? c : n_chunks - 1 - c;
        const int i = chunk * NT + (low_first ? tid : NT - 1 - tid);   // thread order = visiting order
        const bool eqk = i < n_local && keys[i] == v;
        const unsigned b = __ballot_sync(FULL, eqk);
        if (lane == 0) s->warp_tmp[warp] = __popc(b);
        __syncthreads();
        uint32_t off = running;
        for (int w = 0; w < warp; ++w) off += s->warp_tmp[w];
        const uint32_t my = off + __popc(b & ((1u << lane) - 1)) + 1;  // 1-based position among equal keys
        if (eqk && my == (uint32_t)want) s->tie_idx = i;
        for (int w = 0; w < NT / 32; ++w) running += s->warp_tmp[w];
        __syncthreads();
        if (running >= (uint32_t)want) break;                          // uniform: running is block-wide
    }
    __syncthreads();
    return sel_composite(v, index_base + (uint32_t)s->tie_idx, low_first);
#undef SEL_STAMP
}

}  // namespace gssd
