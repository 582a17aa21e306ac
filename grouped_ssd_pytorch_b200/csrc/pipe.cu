// pipe.cu — host-buffer pipeline over the multibox kernels (include/gssd.h: gssd_pipe_*): the hot path as one native
// call per step with HOST inputs and outputs, `depth` steps in flight on three streams (H2D copies | match + loss |
// Detect), so that PCIe traffic, kernels and the device->host results of neighbouring steps overlap.
#include <initializer_list>
#include <new>

#include "common.cuh"

struct gssd_pipe {
    gssd_pipe_cfg cfg;
    const float *priors;
    gssd_pipe_slot slot[8];
    cudaStream_t s_copy, s_main, s_side;
    cudaEvent_t ev_in[8], ev_free[8], ev_side[8], ev_done[8];
    bool busy[8], begun[8], loss_on[8], fused[8];
    int g_sum[8], g_max[8];
    int64_t next;
    bool use_x;
    gssd_xchg x;
    bool det_logits;
    float class_bias[GSSD_MAX_CLASSES];
};

namespace {
size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// walks the arena layout of one slot; with base == nullptr it only measures
size_t layout_slot(const gssd_pipe_cfg &c, uint8_t *base, gssd_pipe_slot *s) {
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t *ptr = base ? base + off : nullptr; off += align_up(bytes); return ptr; };
    const size_t BP = (size_t)c.B * c.P;
    float *loc = (float *)take(BP * 4 * 4);
    float *conf = (float *)take(BP * c.C * 4);                               // loc | conf | gt | gt_off stay adjacent: ONE H2D per step
    float *gt = (float *)take((size_t)c.max_gt_rows * 5 * 4);                // when the host buffers are laid out the same way
    int32_t *gt_off = (int32_t *)take((size_t)(c.B + 1) * 4);
    float *scores = (float *)take(BP * c.C * 4);
    uint16_t *tags = (uint16_t *)take(BP * 2);
    void *stats = take(gssd_stats_bytes(c.B));
    float *losses = (float *)take(16);
    float *grad_loc = (float *)take(BP * 4 * 4);
    float *grad_conf = (float *)take(BP * c.C * 4);
    float *out = (float *)take((size_t)c.B * c.C * c.top_k * 5 * 4);
    const size_t ws_bytes = gssd_workspace_bytes(GSSD_WS_LOSS, c.B, c.P, c.C, c.max_gt_rows, 0);
    void *ws = take(ws_bytes);
    void *fstate = take(gssd_fused_state_bytes());
    if (s) *s = gssd_pipe_slot{loc, conf, scores, gt, gt_off, tags, stats, losses, grad_loc, grad_conf, out, ws, ws_bytes, fstate};
    return off;
}

bool cfg_ok(const gssd_pipe_cfg *c) {
    return c && c->B > 0 && c->P > 0 && c->C >= 2 && c->top_k > 0 && c->max_gt_rows >= c->B && c->depth >= 1 && c->depth <= 8;
}
int64_t cuda_err(cudaError_t e) { return -(1000 + (int64_t)e); }
#define PIPE_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return cuda_err(_e); } while (0)
#define PIPE_RC(expr) do { int _rc = (expr); if (_rc != 0) return _rc < 0 ? (int64_t)_rc : cuda_err((cudaError_t)_rc); } while (0)

// H2D of one step's inputs + matching + Detect (+ its D2H): everything that does not need the global statistics
// two_stage: matching is launched here and the loss in finish_step (the caller all-gathers the statistics in between);
// otherwise finish_step runs the whole loss as one launch when the batch has a one-launch form (fused.cu)
int64_t begin_step(gssd_pipe *p, const float *loc_h, const float *conf_h, const float *scores_h, const float *gt_h,
                   const int32_t *gt_off_h, int sum_g, int g_max, float *det_h, bool two_stage) {
    const gssd_pipe_cfg &c = p->cfg;
    // gt_h == NULL: a Detect-only step (inference, ssd_multiphase_custom_group.py:384-390) — no matching, no loss
    const bool do_loss = gt_h != nullptr;
    if (!loc_h || !conf_h) return GSSD_ERR_ARG;
    if (!do_loss && !det_h) return GSSD_ERR_ARG;
    if (do_loss) {
        if (!gt_off_h || sum_g <= 0 || g_max <= 0) return GSSD_ERR_ARG;
        if (sum_g > c.max_gt_rows) return GSSD_ERR_LIMIT;
        // the row offsets are host memory: check them here, the kernels size their shared memory from g_max and index gt by them
        if (gt_off_h[0] != 0 || gt_off_h[c.B] != sum_g) return GSSD_ERR_ARG;
        for (int b = 0; b < c.B; ++b) {
            const int rows = gt_off_h[b + 1] - gt_off_h[b];
            if (rows <= 0) return rows == 0 ? GSSD_ERR_EMPTY : GSSD_ERR_ARG;
            if (rows > g_max) return GSSD_ERR_ARG;
        }
    }
    const int64_t ticket = p->next;
    const int k = (int)(ticket % c.depth);
    if (p->begun[k]) return GSSD_ERR_ARG;                                    // gssd_pipe_begin `depth` steps ago was never finished
    if (p->busy[k]) { PIPE_CUDA(cudaEventSynchronize(p->ev_done[k])); p->busy[k] = false; }
    const gssd_pipe_slot &s = p->slot[k];
    const size_t BP = (size_t)c.B * c.P, n_loc = BP * 4 * 4, n_conf = BP * c.C * 4;
    PIPE_CUDA(cudaStreamWaitEvent(p->s_copy, p->ev_free[k], 0));             // the kernels that read this slot are done
    const bool do_detect = det_h != nullptr && (p->det_logits || scores_h != nullptr);
    const bool have_scores = do_detect && !p->det_logits;
    const uint8_t *l8 = (const uint8_t *)loc_h, *c8 = (const uint8_t *)conf_h, *g8 = (const uint8_t *)gt_h, *o8 = (const uint8_t *)gt_off_h;
    const size_t n_gt_max = align_up((size_t)c.max_gt_rows * 5 * 4);
    const bool conf_adj = c8 == l8 + align_up(n_loc);
    // every small copy costs the DMA engine microseconds of its own: with the host buffers in the slot's layout (HostBuffers in
    // pipeline.py) a step's inputs are ONE transfer — loc | conf | gt rows (up to max_gt_rows) | row offsets
    const bool all_adj = do_loss && conf_adj && g8 == c8 + align_up(n_conf) && o8 == g8 + n_gt_max;
    if (all_adj) {
        PIPE_CUDA(cudaMemcpyAsync(s.loc, loc_h, align_up(n_loc) + align_up(n_conf) + n_gt_max + (size_t)(c.B + 1) * 4, cudaMemcpyHostToDevice, p->s_copy));
    } else {
        if (conf_adj) {
            PIPE_CUDA(cudaMemcpyAsync(s.loc, loc_h, align_up(n_loc) + n_conf, cudaMemcpyHostToDevice, p->s_copy));
        } else {
            PIPE_CUDA(cudaMemcpyAsync(s.loc, loc_h, n_loc, cudaMemcpyHostToDevice, p->s_copy));
            PIPE_CUDA(cudaMemcpyAsync(s.conf, conf_h, n_conf, cudaMemcpyHostToDevice, p->s_copy));
        }
        if (do_loss) {
            PIPE_CUDA(cudaMemcpyAsync(s.gt, gt_h, (size_t)sum_g * 5 * 4, cudaMemcpyHostToDevice, p->s_copy));
            PIPE_CUDA(cudaMemcpyAsync(s.gt_off, gt_off_h, (size_t)(c.B + 1) * 4, cudaMemcpyHostToDevice, p->s_copy));
        }
    }
    if (have_scores) PIPE_CUDA(cudaMemcpyAsync(s.scores, scores_h, n_conf, cudaMemcpyHostToDevice, p->s_copy));
    PIPE_CUDA(cudaEventRecord(p->ev_in[k], p->s_copy));
    PIPE_CUDA(cudaStreamWaitEvent(p->s_main, p->ev_in[k], 0));
    const bool fused = do_loss && !two_stage && gssd_mbox_fused_supported(c.B, c.P, c.C, g_max) != 0;
    if (!do_loss || fused) {
    } else if (p->use_x)
        PIPE_RC(gssd_mbox_match_x(p->priors, c.P, s.conf, c.C, s.gt, s.gt_off, c.B, sum_g, g_max, c.match_thresh, s.tags, s.stats, &p->x, p->s_main));
    else
        PIPE_RC(gssd_mbox_match(p->priors, c.P, s.conf, c.C, s.gt, s.gt_off, c.B, sum_g, g_max, c.match_thresh, s.tags, s.stats, p->s_main));
    if (do_detect) {
        PIPE_CUDA(cudaStreamWaitEvent(p->s_side, p->ev_in[k], 0));
        if (p->det_logits)
            PIPE_RC(gssd_detect_logits(s.loc, s.conf, p->class_bias, p->priors, c.B, c.P, c.C, c.top_k, c.conf_thresh, c.nms_thresh,
                                       c.var0, c.var1, s.detect_out, nullptr, nullptr, p->s_side));
        else
            PIPE_RC(gssd_detect(s.loc, s.scores, p->priors, c.B, c.P, c.C, c.top_k, c.conf_thresh, c.nms_thresh, c.var0, c.var1,
                                s.detect_out, nullptr, nullptr, p->s_side));
        PIPE_CUDA(cudaMemcpyAsync(det_h, s.detect_out, (size_t)c.B * c.C * c.top_k * 5 * 4, cudaMemcpyDeviceToHost, p->s_side));
    }
    PIPE_CUDA(cudaEventRecord(p->ev_side[k], p->s_side));
    p->g_sum[k] = sum_g; p->g_max[k] = g_max; p->begun[k] = true; p->loss_on[k] = do_loss; p->fused[k] = fused;
    p->next = ticket + 1;
    return ticket;
}

int64_t finish_step(gssd_pipe *p, int64_t ticket, const gssd_loss_stats *global_stats, int n_global, float *losses_h) {
    const gssd_pipe_cfg &c = p->cfg;
    if (ticket < 0 || ticket >= p->next || ticket + c.depth < p->next) return GSSD_ERR_ARG;
    const int k = (int)(ticket % c.depth);
    if (!p->begun[k]) return GSSD_ERR_ARG;
    if (p->loss_on[k] && !losses_h) return GSSD_ERR_ARG;
    const gssd_pipe_slot &s = p->slot[k];
    if (!p->loss_on[k]) {
    } else if (p->fused[k])
        PIPE_RC(gssd_mbox_loss_fused(s.loc, s.conf, p->priors, c.B, c.P, c.C, s.gt, s.gt_off, p->g_sum[k], p->g_max[k], c.match_thresh,
                                     c.negpos_ratio, c.var0, c.var1, s.fused_state, p->use_x ? &p->x : nullptr, s.losses, s.grad_loc,
                                     s.grad_conf, nullptr, nullptr, nullptr, s.ws, s.ws_bytes, p->s_main));
    else if (p->use_x && global_stats == nullptr)
        PIPE_RC(gssd_mbox_loss_x(s.loc, s.conf, p->priors, c.B, c.P, c.C, s.gt, s.gt_off, p->g_sum[k], p->g_max[k], s.tags, s.stats,
                                 &p->x, c.negpos_ratio, c.var0, c.var1, s.losses, s.grad_loc, s.grad_conf, nullptr, nullptr,
                                 s.ws, s.ws_bytes, p->s_main));
    else
        PIPE_RC(gssd_mbox_loss(s.loc, s.conf, p->priors, c.B, c.P, c.C, s.gt, s.gt_off, p->g_sum[k], p->g_max[k], s.tags, s.stats,
                               global_stats, n_global, c.negpos_ratio, c.var0, c.var1, s.losses, s.grad_loc, s.grad_conf, nullptr, nullptr,
                               s.ws, s.ws_bytes, p->s_main));
    if (p->loss_on[k]) PIPE_CUDA(cudaMemcpyAsync(losses_h, s.losses, 8, cudaMemcpyDeviceToHost, p->s_main));
    PIPE_CUDA(cudaStreamWaitEvent(p->s_main, p->ev_side[k], 0));             // Detect and its D2H belong to the step
    PIPE_CUDA(cudaEventRecord(p->ev_free[k], p->s_main));
    PIPE_CUDA(cudaEventRecord(p->ev_done[k], p->s_main));
    p->busy[k] = true; p->begun[k] = false;
    return 0;
}
}  // namespace

extern "C" size_t gssd_pipe_arena_bytes(const gssd_pipe_cfg *cfg) {
    if (!cfg_ok(cfg)) return 0;
    return layout_slot(*cfg, nullptr, nullptr) * (size_t)cfg->depth + 256;
}

extern "C" void gssd_pipe_destroy(gssd_pipe *p);

extern "C" int gssd_pipe_create(gssd_pipe **out, const gssd_pipe_cfg *cfg, const float *priors, void *arena, size_t arena_bytes) {
    if (!out || !cfg_ok(cfg) || !priors || !arena) return GSSD_ERR_ARG;
    if (cfg->P > GSSD_MAX_PRIORS || cfg->top_k > GSSD_MAX_TOP_K || cfg->C > GSSD_MAX_CLASSES) return GSSD_ERR_LIMIT;
    if (!(cfg->var0 > 0.f) || !(cfg->var1 > 0.f) || !(cfg->nms_thresh > 0.f)) return GSSD_ERR_VALUE;
    if (arena_bytes < gssd_pipe_arena_bytes(cfg)) return GSSD_ERR_WS;
    gssd_pipe *p = new (std::nothrow) gssd_pipe();
    if (!p) return GSSD_ERR_ARG;
    p->cfg = *cfg; p->priors = priors; p->next = 0; p->use_x = false; p->det_logits = false;
    p->s_copy = p->s_main = p->s_side = nullptr;
    for (int k = 0; k < 8; ++k) p->ev_in[k] = p->ev_free[k] = p->ev_side[k] = p->ev_done[k] = nullptr;
    for (int i = 0; i < GSSD_MAX_CLASSES; ++i) p->class_bias[i] = 0.f;
    uint8_t *base = (uint8_t *)(((uintptr_t)arena + 255) & ~(uintptr_t)255);
    const size_t per = layout_slot(*cfg, nullptr, nullptr);
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < cfg->depth; ++k) {
        layout_slot(*cfg, base + per * k, &p->slot[k]);
        p->busy[k] = p->begun[k] = p->loss_on[k] = p->fused[k] = false;
        if (e == cudaSuccess) e = cudaMemset(p->slot[k].fused_state, 0, gssd_fused_state_bytes());
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_main, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_side, cudaStreamNonBlocking);
    for (int k = 0; k < cfg->depth && e == cudaSuccess; ++k) {
        e = cudaEventCreateWithFlags(&p->ev_in[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_free[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_side[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_done[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { gssd_pipe_destroy(p); return (int)e; }
    *out = p;
    return GSSD_OK;
}

extern "C" void gssd_pipe_destroy(gssd_pipe *p) {
    if (!p) return;
    // also the unwinding path of a gssd_pipe_create that failed half-way: handles that were never created are null
    for (cudaStream_t st : {p->s_copy, p->s_main, p->s_side}) if (st) cudaStreamSynchronize(st);
    for (int k = 0; k < 8; ++k)
        for (cudaEvent_t ev : {p->ev_in[k], p->ev_free[k], p->ev_side[k], p->ev_done[k]}) if (ev) cudaEventDestroy(ev);
    for (cudaStream_t st : {p->s_copy, p->s_main, p->s_side}) if (st) cudaStreamDestroy(st);
    delete p;
}

extern "C" int gssd_pipe_set_detect_logits(gssd_pipe *p, int enable, const float *class_bias) {
    if (!p) return GSSD_ERR_ARG;
    p->det_logits = enable != 0;
    for (int i = 0; i < p->cfg.C && i < GSSD_MAX_CLASSES; ++i) p->class_bias[i] = (enable && class_bias) ? class_bias[i] : 0.f;
    return GSSD_OK;
}

extern "C" int gssd_pipe_set_xchg(gssd_pipe *p, const gssd_xchg *x) {
    if (!p) return GSSD_ERR_ARG;
    if (!x) { p->use_x = false; return GSSD_OK; }
    if (x->world < 1 || x->world > GSSD_XCHG_MAX_RANKS || x->rank < 0 || x->rank >= x->world) return GSSD_ERR_ARG;
    p->x = *x; p->use_x = true;
    return GSSD_OK;
}

extern "C" int gssd_pipe_slot_info(const gssd_pipe *p, int slot, gssd_pipe_slot *out) {
    if (!p || !out || slot < 0 || slot >= p->cfg.depth) return GSSD_ERR_ARG;
    *out = p->slot[slot];
    return GSSD_OK;
}

extern "C" int64_t gssd_pipe_submit(gssd_pipe *p, const float *loc_h, const float *conf_h, const float *scores_h, const float *gt_h,
                                    const int32_t *gt_off_h, int sum_g, int g_max, float *losses_h, float *det_h) {
    if (!p) return GSSD_ERR_ARG;
    const int64_t t = begin_step(p, loc_h, conf_h, scores_h, gt_h, gt_off_h, sum_g, g_max, det_h, false);
    if (t < 0) return t;
    const int64_t rc = finish_step(p, t, nullptr, 0, losses_h);
    return rc < 0 ? rc : t;
}

extern "C" int64_t gssd_pipe_begin(gssd_pipe *p, const float *loc_h, const float *conf_h, const float *scores_h, const float *gt_h,
                                   const int32_t *gt_off_h, int sum_g, int g_max, float *det_h, void **stream_out) {
    if (!p) return GSSD_ERR_ARG;
    const int64_t t = begin_step(p, loc_h, conf_h, scores_h, gt_h, gt_off_h, sum_g, g_max, det_h, true);
    if (t >= 0 && stream_out) *stream_out = (void *)p->s_main;
    return t;
}

extern "C" int gssd_pipe_finish(gssd_pipe *p, int64_t ticket, const gssd_loss_stats *global_stats, int n_global, float *losses_h) {
    if (!p) return GSSD_ERR_ARG;
    const int64_t rc = finish_step(p, ticket, global_stats, n_global, losses_h);
    return rc < -1000 ? (int)(-rc - 1000) : (int)rc;
}

extern "C" int gssd_pipe_wait(gssd_pipe *p, int64_t ticket) {
    if (!p || ticket < 0 || ticket >= p->next) return GSSD_ERR_ARG;
    if (ticket + p->cfg.depth < p->next) return GSSD_OK;                     // long finished: its slot has been reused since
    const int k = (int)(ticket % p->cfg.depth);
    if (p->begun[k]) return GSSD_ERR_ARG;                                    // begin without finish
    if (p->busy[k]) {
        cudaError_t e = cudaEventSynchronize(p->ev_done[k]);
        if (e != cudaSuccess) return (int)e;
        p->busy[k] = false;
    }
    return GSSD_OK;
}
