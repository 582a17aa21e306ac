// match.cu — prior <-> ground-truth matching (box_utils.py:70-111, batched over the loop at
// multibox_loss.py:67-72).
//
// One thread-block cluster per image (1..8 CTAs, contiguous prior slices).  GT boxes are staged in
// shared memory; each thread sweeps its priors (one coalesced float4 each), keeps the best GT per
// prior in registers and feeds a per-GT running best prior through a warp redux + ballot into a
// shared-memory atomicMax on a packed (IoU bits, ~prior index) key, so that ties resolve to the lowest
// prior index exactly like torch.max.  The per-GT winners are combined over distributed shared
// memory, then the sequential "force match" (box_utils.py:101-105, last GT wins) is applied by one
// thread, and the result is emitted either as 2-byte tags for the fused loss or as the reference's
// materialised loc_t / conf_t.
#include "common.cuh"

namespace gssd {

constexpr int MATCH_NT = 256;

struct MatchArgs {
    const float4 *priors; int P;
    const float *conf; int C;                 // optional: batch max of conf (stage 1 of the loss)
    const float *gt; const int32_t *gt_off;
    float threshold, var0, var1;
    uint16_t *tags;                           // [B,P] or null
    uint32_t *stats;                          // stats header (conf_max_ord, num_pos_total, ...) or null
    int32_t *num_pos;                         // [B] or null
    float4 *loc_t; int64_t *conf_t; int32_t *bti_out;   // materialised outputs or null
    int slice;                                // priors per CTA
};

// dynamic shared memory layout
//   u64    sbest[G]   packed (IoU bits, ~prior) best prior per GT, this CTA's slice
//   float  sgt[G][6]  (x1,y1,x2,y2,area,label)
//   int    sbp[G]     best prior per GT over the whole image
//   u16    stag[slice]
static size_t match_smem_bytes(int g_max, int slice) {
    return (size_t)g_max * (8 + 6 * 4 + 4) + (size_t)slice * 2 + 16;
}

template <bool MATERIALISE, bool CONF_MAX>
__global__ void __launch_bounds__(MATCH_NT) match_kernel(MatchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned nranks = cluster.num_blocks();
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int g0 = a.gt_off[b];
    const int G = a.gt_off[b + 1] - g0;

    unsigned long long *sbest = reinterpret_cast<unsigned long long *>(smem_raw);     // 8-byte aligned first
    float *sgt = reinterpret_cast<float *>(sbest + G);
    int *sbp = reinterpret_cast<int *>(sgt + 6 * G);
    uint16_t *stag = reinterpret_cast<uint16_t *>(sbp + G);
    __shared__ int s_warp_cnt[MATCH_NT / 32];
    __shared__ float s_warp_max[MATCH_NT / 32];

    for (int g = tid; g < G; g += MATCH_NT) {
        const float *row = a.gt + 5 * (size_t)(g0 + g);
        float4 t = make_float4(row[0], row[1], row[2], row[3]);
        sgt[6 * g + 0] = t.x; sgt[6 * g + 1] = t.y; sgt[6 * g + 2] = t.z; sgt[6 * g + 3] = t.w;
        sgt[6 * g + 4] = box_area(t);
        sgt[6 * g + 5] = row[4];
        sbest[g] = 0ull;
    }
    __syncthreads();

    const int p0 = rank * a.slice;
    const int p1 = min(a.P, p0 + a.slice);
    float cmax = -INFINITY;

    // ---- IoU sweep -----------------------------------------------------------------------------
    for (int base = p0 + (tid & ~31); base < p1; base += MATCH_NT) {    // warp-uniform trip count
        const int p = base + lane;
        const bool valid = p < p1;
        float4 pr = valid ? a.priors[p] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 pb = point_form(pr);
        float area_b = box_area(pb);
        if (CONF_MAX && valid) {
            const float *row = a.conf + ((size_t)b * a.P + p) * a.C;
            if (a.C == 2) {
                float2 v = *reinterpret_cast<const float2 *>(row);
                cmax = fmaxf(cmax, fmaxf(v.x, v.y));
            } else {
                for (int c = 0; c < a.C; ++c) cmax = fmaxf(cmax, row[c]);
            }
        }
        float best = -1.f;
        int bidx = 0;
        for (int g = 0; g < G; ++g) {
            float4 t = make_float4(sgt[6 * g], sgt[6 * g + 1], sgt[6 * g + 2], sgt[6 * g + 3]);
            float iou = valid ? box_iou_fast(t, sgt[6 * g + 4], pb, area_b) : 0.f;
            if (iou > best) { best = iou; bidx = g; }            // first max over GT (torch.max dim 0)
            // best prior for this GT: warp max of the IoU bits (IoU >= +0, so uint order == float order),
            // lowest lane among the maxima == lowest prior index
            unsigned bits = __float_as_uint(iou);
            unsigned m = __reduce_max_sync(FULL, bits);
            unsigned who = __ballot_sync(FULL, valid && bits == m);
            if (lane == 0) {
                unsigned long long key = ((unsigned long long)m << 32) | (0xffffffffu - (unsigned)(base + __ffs(who) - 1));
                if (key > sbest[g]) atomicMax(&sbest[g], key);
            }
        }
        if (valid) stag[p - p0] = (uint16_t)(bidx | (!(best < a.threshold) ? 0x8000 : 0));   // box_utils.py:108
    }
    __syncthreads();

    // ---- best prior per GT over the whole image, then the sequential force match -------------------
    if (nranks > 1) cluster.sync();
    for (int g = tid; g < G; g += MATCH_NT) {
        unsigned long long m = sbest[g];
        for (unsigned r = 0; r < nranks; ++r) {
            if (r == rank) continue;
            unsigned long long o = cluster.map_shared_rank(sbest, r)[g];
            m = o > m ? o : m;
        }
        sbp[g] = (int)(0xffffffffu - (unsigned)(m & 0xffffffffu));
    }
    __syncthreads();
    if (tid == 0) {
        for (int g = 0; g < G; ++g) {                            // box_utils.py:101-105, in GT order
            int bp = sbp[g];
            if (bp >= p0 && bp < p1) stag[bp - p0] = (uint16_t)(0x8000 | g);
        }
    }
    __syncthreads();

    // ---- emit --------------------------------------------------------------------------------------
    int npos = 0;
    for (int p = p0 + tid; p < p1; p += MATCH_NT) {
        uint16_t tag = stag[p - p0];
        const bool pos = tag & 0x8000;
        const int g = tag & 0x7fff;
        npos += pos;
        const size_t o = (size_t)b * a.P + p;
        if (a.tags) a.tags[o] = tag;
        if (MATERIALISE) {
            float4 t = make_float4(sgt[6 * g], sgt[6 * g + 1], sgt[6 * g + 2], sgt[6 * g + 3]);
            a.loc_t[o] = encode_box(t, a.priors[p], a.var0, a.var1);                  // box_utils.py:109-110
            float c = pos ? __fadd_rn(sgt[6 * g + 5], 1.f) : 0.f;                     // 107-108
            a.conf_t[o] = (int64_t)c;                                                 // 111
            if (a.bti_out) a.bti_out[o] = g;
        }
    }
    if (a.num_pos || CONF_MAX) {
        npos = warp_sum(npos);
        cmax = warp_max(cmax);
        if (lane == 0) { s_warp_cnt[tid >> 5] = npos; s_warp_max[tid >> 5] = cmax; }
        __syncthreads();
        if (tid == 0) {
            int tot = 0; float mx = -INFINITY;
            for (int w = 0; w < MATCH_NT / 32; ++w) { tot += s_warp_cnt[w]; mx = fmaxf(mx, s_warp_max[w]); }
            if (a.num_pos) atomicAdd(&a.num_pos[b], tot);
            if (a.stats) {
                atomicAdd(reinterpret_cast<int *>(&a.stats[1]), tot);
                if (CONF_MAX && p1 > p0) atomicMax(&a.stats[0], f2ord(mx));
            }
        }
    }
    if (nranks > 1) cluster.sync();     // keep sbest alive until every CTA of the image has read it
}

static int pick_cluster(int B, int P) {
    // enough CTAs to cover the 148 SMs about twice, at most 8 per image, at least ~512 priors per CTA
    int s = 1;
    while (s < 8 && B * s < 296 && P / (s * 2) >= 512) s *= 2;
    return s;
}

template <bool MAT, bool CMAX>
static int launch_match(const MatchArgs &a_in, int B, int g_max, cudaStream_t stream) {
    MatchArgs a = a_in;
    int S = pick_cluster(B, a.P);
    a.slice = ceil_div(a.P, S);
    size_t smem = match_smem_bytes(g_max, a.slice);
    auto kern = match_kernel<MAT, CMAX>;
    if (smem > 48 * 1024)
        GSSD_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(S, B, 1);
    cfg.blockDim = dim3(MATCH_NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GSSD_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

static int check_match_sizes(int P, int B, int sum_G, int g_max) {
    if (P <= 0 || B <= 0 || sum_G <= 0 || g_max <= 0) return sum_G <= 0 && B > 0 ? GSSD_ERR_EMPTY : GSSD_ERR_ARG;
    if (g_max > GSSD_MAX_GT_PER_IMAGE || P > GSSD_MAX_PRIORS) return GSSD_ERR_LIMIT;
    return GSSD_OK;
}

}  // namespace gssd

using namespace gssd;

extern "C" int gssd_match(const float *priors, int P, const float *gt, const int32_t *gt_off, int B,
                          int sum_G, int g_max, float threshold, float var0, float var1,
                          float *loc_t, int64_t *conf_t, int32_t *best_truth_idx,
                          void *ws, size_t ws_bytes, void *stream) {
    (void)ws; (void)ws_bytes;
    if (!priors || !gt || !gt_off || !loc_t || !conf_t) return GSSD_ERR_ARG;
    int rc = check_match_sizes(P, B, sum_G, g_max);
    if (rc) return rc;
    MatchArgs a = {};
    a.priors = reinterpret_cast<const float4 *>(priors); a.P = P;
    a.gt = gt; a.gt_off = gt_off; a.threshold = threshold; a.var0 = var0; a.var1 = var1;
    a.loc_t = reinterpret_cast<float4 *>(loc_t); a.conf_t = conf_t; a.bti_out = best_truth_idx;
    return launch_match<true, false>(a, B, g_max, (cudaStream_t)stream);
}

extern "C" int gssd_mbox_match(const float *priors, int P, const float *conf, int C,
                               const float *gt, const int32_t *gt_off, int B, int sum_G, int g_max,
                               float threshold, uint16_t *tags, void *stats_buf, void *stream) {
    if (!priors || !gt || !gt_off || !tags || !stats_buf) return GSSD_ERR_ARG;
    int rc = check_match_sizes(P, B, sum_G, g_max);
    if (rc) return rc;
    if (conf && (C < 2 || C > GSSD_MAX_CLASSES)) return GSSD_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    GSSD_RETURN_IF_CUDA(cudaMemsetAsync(stats_buf, 0, sizeof(gssd_loss_stats) + sizeof(int32_t) * (size_t)B, st));
    MatchArgs a = {};
    a.priors = reinterpret_cast<const float4 *>(priors); a.P = P;
    a.conf = conf; a.C = C;
    a.gt = gt; a.gt_off = gt_off; a.threshold = threshold;
    a.tags = tags;
    a.stats = reinterpret_cast<uint32_t *>(stats_buf);
    a.num_pos = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(stats_buf) + sizeof(gssd_loss_stats));
    return conf ? launch_match<false, true>(a, B, g_max, st) : launch_match<false, false>(a, B, g_max, st);
}
