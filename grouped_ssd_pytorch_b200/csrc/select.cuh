// select.cuh — per-image radix select in shared memory (replaces the reference's double argsort,
// multibox_loss.py:102-106, and the full sort + slice of nms, box_utils.py:194-196).
//
// The keys of one image live in shared memory as order-preserving uint32 (f2ord), either in one CTA
// or split in contiguous slices across the CTAs of a thread-block cluster; histograms are then
// summed over distributed shared memory.  The result is the exact "k largest keys" set with a
// deterministic rule for equal keys at the cut (lower index first or higher index first).
#pragma once
#include "common.cuh"

namespace gssd {

struct SelectShared {
    uint32_t hist[2][256];   // double-buffered so that one cluster.sync per pass is enough
    uint32_t total[256];
    uint32_t warp_tmp[32];
    uint32_t digit, k_rem, eq_total, eq_local_before;
    int      tie_cut;
};

struct SelectResult {
    uint32_t v;        // value of the k-th largest key
    uint32_t need;     // how many keys equal to v belong to the selection (>= 1)
    uint32_t eq;       // how many keys equal to v exist (whole image)
    int      tie_cut;  // local index bound for equal keys, see selected()
    bool     low_first;
    __device__ __forceinline__ bool selected(uint32_t key, int local_idx) const {
        if (key > v) return true;
        if (key != v) return false;
        return low_first ? local_idx < tie_cut : local_idx >= tie_cut;
    }
};

// NT threads per CTA (multiple of 32, >= 256).  keys[0..n_local) are this CTA's slice.  k >= 1 and
// k <= total number of keys in the image.  CLUSTER: slices are ordered by cluster rank.
template <int NT, bool CLUSTER>
__device__ SelectResult radix_select(const uint32_t *keys, int n_local, uint32_t k, bool low_first,
                                     SelectShared *s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned nranks = CLUSTER ? cluster.num_blocks() : 1;
    const unsigned rank = CLUSTER ? cluster.block_rank() : 0;

    uint32_t prefix = 0, mask = 0, k_rem = k;
    int buf = 0;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8, buf ^= 1) {
        for (int i = tid; i < 256; i += NT) s->hist[buf][i] = 0;
        __syncthreads();
        const int n_round = (n_local + NT - 1) / NT * NT;
        for (int i = tid; i < n_round; i += NT) {
            bool in = i < n_local;
            uint32_t key = in ? keys[i] : 0;
            in = in && ((key & mask) == prefix);
            unsigned act = __ballot_sync(FULL, in);
            if (in) {
                uint32_t d = (key >> shift) & 255u;
                unsigned peers = __match_any_sync(act, d);       // one atomic per distinct digit
                if (lane == __ffs(peers) - 1) atomicAdd(&s->hist[buf][d], __popc(peers));
            }
        }
        __syncthreads();
        if (CLUSTER) {
            cluster.sync();
            for (int i = tid; i < 256; i += NT) {
                uint32_t t = 0;
                for (unsigned r = 0; r < nranks; ++r) t += cluster.map_shared_rank(&s->hist[buf][0], r)[i];
                s->total[i] = t;
            }
        } else {
            for (int i = tid; i < 256; i += NT) s->total[i] = s->hist[buf][i];
        }
        __syncthreads();
        // suffix sums over the 256 bins (bin 255 first): the first 8 warps own one bin per thread
        if (tid < 256) {
            uint32_t c = s->total[tid];
            uint32_t incl = c;                                   // inclusive suffix within the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_down_sync(FULL, incl, o);
                if (lane + o < 32) incl += t;
            }
            if (lane == 0) s->warp_tmp[warp] = incl;             // warp total
            __syncwarp();
            // named barrier over the first 256 threads only
            asm volatile("bar.sync 1, 256;");
            uint32_t above = 0;
            for (int w = warp + 1; w < 8; ++w) above += s->warp_tmp[w];
            uint32_t excl = above + incl - c;                    // keys in bins > tid
            if (excl < k_rem && k_rem <= excl + c) {
                s->digit = tid; s->k_rem = k_rem - excl; s->eq_total = c;
            }
        }
        __syncthreads();
        prefix |= s->digit << shift;
        mask |= 255u << shift;
        k_rem = s->k_rem;
    }
    buf ^= 1;   // the buffer of the last pass: hist[buf][digit] = local count of keys == prefix

    SelectResult r;
    r.v = prefix; r.need = k_rem; r.eq = s->eq_total; r.low_first = low_first;
    // All ties selected: no index rule needed.
    if (r.need == r.eq) {
        r.tie_cut = low_first ? 0x7fffffff : 0;
        if (CLUSTER) cluster.sync();        // nobody leaves while its histograms may still be read
        return r;
    }
    // Partial tie: `need` of the `eq` equal keys, in index order (from the low or the high end).
    uint32_t eq_local = s->hist[buf][s->digit];
    uint32_t before = 0;                    // equal keys in slices of lower rank
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
    if (low_first) want = (long long)r.need - before;
    else           want = (long long)r.need - ((long long)r.eq - before - eq_local);
    if (want <= 0) { r.tie_cut = low_first ? 0 : 0x7fffffff; return r; }
    if (want >= (long long)eq_local) { r.tie_cut = low_first ? 0x7fffffff : 0; return r; }
    // find the local index of the want-th equal key from the preferred end (block scan in index order)
    if (tid == 0) s->tie_cut = -1;
    __syncthreads();
    uint32_t running = 0;
    const int n_chunks = (n_local + NT - 1) / NT;
    for (int c = 0; c < n_chunks; ++c) {
        int chunk = low_first ? c : n_chunks - 1 - c;
        int i = chunk * NT + (low_first ? tid : NT - 1 - tid);   // thread order = visiting order
        bool eqk = i < n_local && keys[i] == r.v;
        unsigned b = __ballot_sync(FULL, eqk);
        if (lane == 0) s->warp_tmp[warp] = __popc(b);
        __syncthreads();
        uint32_t off = running;
        for (int w = 0; w < warp; ++w) off += s->warp_tmp[w];
        uint32_t my = off + __popc(b & ((1u << lane) - 1)) + 1; // 1-based position among equal keys
        if (eqk && my == (uint32_t)want) s->tie_cut = i;
        for (int w = 0; w < NT / 32; ++w) running += s->warp_tmp[w];
        __syncthreads();
        if (running >= (uint32_t)want) break;                    // uniform: running is block-wide
    }
    __syncthreads();
    r.tie_cut = low_first ? s->tie_cut + 1 : s->tie_cut;          // low: idx < cut ; high: idx >= cut
    return r;
}

}  // namespace gssd
