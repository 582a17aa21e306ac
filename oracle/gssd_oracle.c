/*
 * gssd_oracle.c — CPU restatement of the GSSD multibox hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for libgssd_b200.so.  It restates, in scalar C, the algorithm of
 * the reference's `ssd_liverdet/layers` package, function by function, in the reference's own
 * operation order (every function cites the file:line it follows; paths are relative to
 * /root/reference/ssd_liverdet/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py imports the unmodified reference (torch CPU)
 * in the build container, runs it on seeded inputs and commits inputs' seeds + outputs under
 * tests/golden/ (npz files); tests/test_oracle_golden.py checks every function here against them
 * (integer outputs bit-exact, float outputs to 1e-6 relative: glibc expf/logf vs torch's Sleef).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, float stays float;
 *        -fopenmp: images are processed in parallel, OMP_NUM_THREADS selects the core count)
 *
 * Tie contracts where the reference's unstable sorts leave the order undefined:
 *   OHNM : descending key, equal keys -> lower prior index first      (multibox_loss.py:102-106)
 *   NMS  : ascending score, equal scores -> lower index first, the list is consumed from its end,
 *          so among equal scores the HIGHER index is visited first     (box_utils.py:194-207)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gssd.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* PriorBox — layers/functions/prior_box.py:14-172.  Python floats are doubles; the list is cast
 * to float32 once by torch.Tensor(mean) (line 168) and then clamped (170-171). */
static int prior_cfg_check(const gssd_prior_cfg *c) {
    if (!c || c->n_maps <= 0 || c->n_maps > GSSD_MAX_FEATURE_MAPS) return GSSD_ERR_ARG;
    for (int k = 0; k < c->n_maps; ++k)
        if (c->n_ar[k] < 0 || c->n_ar[k] > GSSD_MAX_ASPECT_RATIOS || c->feature_maps[k] <= 0)
            return GSSD_ERR_ARG;
    for (int i = 0; i < 2; ++i)                      /* prior_box.py:28-30 */
        if (c->variance[i] <= 0) return GSSD_ERR_VALUE;
    return GSSD_OK;
}

static int legacy_boxes_per_cell(const gssd_prior_cfg *c, int k) {
    int n = 1;                                       /* prior_box.py:150-151 */
    if (c->max_sizes[k] > 0) n += 1;                 /* 152-158 */
    for (int a = 0; a < c->n_ar[k]; ++a)
        if (!(fabs(c->aspect_ratios[k][a] - 1) < 1e-6)) n += 1;   /* 160-165 */
    return n;
}

EXPORT int gssd_oracle_priorbox_count(const gssd_prior_cfg *c) {
    int rc = prior_cfg_check(c);
    if (rc) return rc;
    long total = 0;
    for (int k = 0; k < c->n_maps; ++k) {
        int f = c->feature_maps[k];
        int per = (c->version == GSSD_PRIOR_LEGACY) ? legacy_boxes_per_cell(c, k)
                                                    : 2 + 2 * c->n_ar[k];
        total += (long)f * f * per;
    }
    return (int)total;
}

EXPORT int gssd_oracle_priorbox(const gssd_prior_cfg *c, float *out) {
    int rc = prior_cfg_check(c);
    if (rc) return rc;
    float *o = out;
#define EMIT(a, b, c_, d) do { o[0] = (float)(a); o[1] = (float)(b); o[2] = (float)(c_); o[3] = (float)(d); o += 4; } while (0)
    for (int k = 0; k < c->n_maps; ++k) {
        int f = c->feature_maps[k];
        for (int i = 0; i < f; ++i) for (int j = 0; j < f; ++j) {   /* product(range(f), repeat=2) */
            if (c->version != GSSD_PRIOR_LEGACY) {
                double f_k = c->min_dim / c->steps[k];               /* prior_box.py:38 */
                double cx = (j + 0.5) / f_k, cy = (i + 0.5) / f_k;   /* 40-41 */
                double s_k = c->min_sizes[k] / c->min_dim;           /* 45 */
                EMIT(cx, cy, s_k, s_k);
                double s_kp = sqrt(s_k * (c->max_sizes[k] / c->min_dim));  /* 50 */
                EMIT(cx, cy, s_kp, s_kp);
                for (int a = 0; a < c->n_ar[k]; ++a) {
                    double r = sqrt(c->aspect_ratios[k][a]);
                    if (c->version == GSSD_PRIOR_V2) {               /* 54-56 */
                        EMIT(cx, cy, s_k * r, s_k / r);
                        EMIT(cx, cy, s_k / r, s_k * r);
                    } else {                                         /* 84-85, 113-114: squares */
                        EMIT(cx, cy, s_k * r, s_k * r);
                        EMIT(cx, cy, s_k / r, s_k / r);
                    }
                }
            } else {                                                 /* 141-167, corner form */
                double step = c->min_dim / f;                        /* 144 */
                double c_x = (j + 0.5) * step, c_y = (i + 0.5) * step;
                double c_w = c->min_sizes[k] / 2, c_h = c_w;
                double s = c->min_dim;
                EMIT((c_x - c_w) / s, (c_y - c_h) / s, (c_x + c_w) / s, (c_y + c_h) / s);
                if (c->max_sizes[k] > 0) {
                    c_w = c_h = sqrt(c->min_sizes[k] * c->max_sizes[k]) / 2;
                    EMIT((c_x - c_w) / s, (c_y - c_h) / s, (c_x + c_w) / s, (c_y + c_h) / s);
                }
                for (int a = 0; a < c->n_ar[k]; ++a) {
                    double ar = c->aspect_ratios[k][a];
                    if (!(fabs(ar - 1) < 1e-6)) {
                        c_w = c->min_sizes[k] * sqrt(ar) / 2;
                        c_h = c->min_sizes[k] / sqrt(ar) / 2;
                        EMIT((c_x - c_w) / s, (c_y - c_h) / s, (c_x + c_w) / s, (c_y + c_h) / s);
                    }
                }
            }
        }
    }
#undef EMIT
    if (c->clip)                                                     /* 170-171 */
        for (float *p = out; p < o; ++p) *p = *p < 0.f ? 0.f : (*p > 1.f ? 1.f : *p);
    return GSSD_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* box_utils.py:4-13 */
EXPORT void gssd_oracle_point_form(const float *b, int n, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *p = b + 4 * i; float *o = out + 4 * i;
        o[0] = p[0] - p[2] / 2; o[1] = p[1] - p[3] / 2;
        o[2] = p[0] + p[2] / 2; o[3] = p[1] + p[3] / 2;
    }
}

/* box_utils.py:16-25, documented intent: ((xy1+xy2)/2, xy2-xy1) */
EXPORT void gssd_oracle_center_size(const float *b, int n, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *p = b + 4 * i; float *o = out + 4 * i;
        o[0] = (p[2] + p[0]) / 2; o[1] = (p[3] + p[1]) / 2;
        o[2] = p[2] - p[0];       o[3] = p[3] - p[1];
    }
}

static inline float clamp0(float x) { return x < 0.f ? 0.f : x; }   /* torch.clamp(min=0) */

/* box_utils.py:28-46 */
static inline float inter1(const float *a, const float *b) {
    float mx = fminf(a[2], b[2]) - fmaxf(a[0], b[0]);
    float my = fminf(a[3], b[3]) - fmaxf(a[1], b[1]);
    return clamp0(mx) * clamp0(my);
}

EXPORT void gssd_oracle_intersect(const float *a, int A, const float *b, int Bn, float *out) {
    for (int i = 0; i < A; ++i) for (int j = 0; j < Bn; ++j)
        out[(size_t)i * Bn + j] = inter1(a + 4 * i, b + 4 * j);
}

/* box_utils.py:49-67 */
static inline float iou1(const float *a, const float *b) {
    float inter = inter1(a, b);
    float area_a = (a[2] - a[0]) * (a[3] - a[1]);
    float area_b = (b[2] - b[0]) * (b[3] - b[1]);
    float uni = area_a + area_b - inter;
    return inter / uni;
}

EXPORT void gssd_oracle_jaccard(const float *a, int A, const float *b, int Bn, float *out) {
    for (int i = 0; i < A; ++i) for (int j = 0; j < Bn; ++j)
        out[(size_t)i * Bn + j] = iou1(a + 4 * i, b + 4 * j);
}

/* box_utils.py:114-135.  The Python scalars multiply/divide float32 tensors as float32. */
static inline void encode1(const float *m, const float *p, float v0, float v1, float *o) {
    o[0] = ((m[0] + m[2]) / 2 - p[0]) / (v0 * p[2]);
    o[1] = ((m[1] + m[3]) / 2 - p[1]) / (v0 * p[3]);
    o[2] = logf((m[2] - m[0]) / p[2]) / v1;
    o[3] = logf((m[3] - m[1]) / p[3]) / v1;
}

EXPORT void gssd_oracle_encode(const float *matched, const float *priors, int n, float v0, float v1,
                               float *out) {
    for (int i = 0; i < n; ++i) encode1(matched + 4 * i, priors + 4 * i, v0, v1, out + 4 * i);
}

/* box_utils.py:139-157 */
static inline void decode1(const float *l, const float *p, float v0, float v1, float *o) {
    float cx = p[0] + l[0] * v0 * p[2];
    float cy = p[1] + l[1] * v0 * p[3];
    float w = p[2] * expf(l[2] * v1);
    float h = p[3] * expf(l[3] * v1);
    cx -= w / 2; cy -= h / 2;          /* boxes[:, :2] -= boxes[:, 2:] / 2 */
    w += cx; h += cy;                  /* boxes[:, 2:] += boxes[:, :2]     */
    o[0] = cx; o[1] = cy; o[2] = w; o[3] = h;
}

EXPORT void gssd_oracle_decode(const float *loc, const float *priors, int n, float v0, float v1,
                               float *out) {
    for (int i = 0; i < n; ++i) decode1(loc + 4 * i, priors + 4 * i, v0, v1, out + 4 * i);
}

/* box_utils.py:160-168 */
static float tensor_max(const float *x, size_t n) {
    float m = x[0];
    for (size_t i = 1; i < n; ++i) if (x[i] > m) m = x[i];
    return m;
}

static inline float lse_row(const float *x, int C, float x_max) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(x[c] - x_max);
    return logf(s) + x_max;
}

EXPORT void gssd_oracle_log_sum_exp(const float *x, int rows, int C, float *out) {
    float m = tensor_max(x, (size_t)rows * C);
    for (int r = 0; r < rows; ++r) out[r] = lse_row(x + (size_t)r * C, C, m);
}

/* ------------------------------------------------------------------------------------------ */
/* match — box_utils.py:70-111, one image. */
EXPORT int gssd_oracle_match(float threshold, const float *truths, const float *labels, int G,
                             const float *priors, int P, float v0, float v1,
                             float *loc_t, int64_t *conf_t, int32_t *best_truth_idx_out,
                             float *best_truth_overlap_out) {
    if (G <= 0) return GSSD_ERR_EMPTY;                       /* reference: IndexError */
    float *pf = (float *)malloc(sizeof(float) * 4 * P);
    float *ov = (float *)malloc(sizeof(float) * (size_t)G * P);
    float *bto = (float *)malloc(sizeof(float) * P);
    int32_t *bti = (int32_t *)malloc(sizeof(int32_t) * P);
    int32_t *bpi = (int32_t *)malloc(sizeof(int32_t) * G);
    gssd_oracle_point_form(priors, P, pf);
    gssd_oracle_jaccard(truths, G, pf, P, ov);               /* 88-91 */
    for (int g = 0; g < G; ++g) {                            /* 94: max over priors, first max */
        int best = 0; float bv = ov[(size_t)g * P];
        for (int p = 1; p < P; ++p) if (ov[(size_t)g * P + p] > bv) { bv = ov[(size_t)g * P + p]; best = p; }
        bpi[g] = best;
    }
    for (int p = 0; p < P; ++p) {                            /* 96: max over truths, first max */
        int best = 0; float bv = ov[p];
        for (int g = 1; g < G; ++g) if (ov[(size_t)g * P + p] > bv) { bv = ov[(size_t)g * P + p]; best = g; }
        bti[p] = best; bto[p] = bv;
    }
    for (int g = 0; g < G; ++g) bto[bpi[g]] = 2.f;           /* 101 */
    for (int g = 0; g < G; ++g) bti[bpi[g]] = g;             /* 104-105: sequential, last wins */
    for (int p = 0; p < P; ++p) {
        const float *m = truths + 4 * bti[p];                /* 106 */
        float conf = labels[bti[p]] + 1.f;                   /* 107 */
        if (bto[p] < threshold) conf = 0.f;                  /* 108 */
        encode1(m, priors + 4 * p, v0, v1, loc_t + 4 * p);   /* 109-110 */
        conf_t[p] = (int64_t)conf;                           /* 111: float -> int64 on assignment */
    }
    if (best_truth_idx_out) memcpy(best_truth_idx_out, bti, sizeof(int32_t) * P);
    if (best_truth_overlap_out) memcpy(best_truth_overlap_out, bto, sizeof(float) * P);
    free(pf); free(ov); free(bto); free(bti); free(bpi);
    return GSSD_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* MultiBoxLoss.forward — layers/modules/multibox_loss.py:46-120 + autograd of its outputs. */
typedef struct { float key; int32_t idx; } key_idx;

static int cmp_key_desc(const void *a, const void *b) {       /* stable descending */
    const key_idx *x = (const key_idx *)a, *y = (const key_idx *)b;
    if (x->key > y->key) return -1;
    if (x->key < y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

EXPORT int gssd_oracle_multibox_loss(
        const float *loc, const float *conf, const float *priors, int B, int P, int C,
        const float *gt, const int32_t *gt_off, float threshold, int negpos_ratio, float v0, float v1,
        float *losses /* [2] */, int32_t *num_pos_out /* [B] */,
        float *loc_t_out /* [B,P,4] opt */, int64_t *conf_t_out /* [B,P] opt */,
        uint8_t *pos_out /* [B,P] opt */, uint8_t *neg_out /* [B,P] opt */,
        float *key_out /* [B,P] opt: mining key after loss_c[pos]=0 */,
        float *kth_gap_out /* [B] opt: key[k-1]-key[k] of the sorted keys (0 = tie at the cut) */,
        float *grad_loc /* [B,P,4] opt */, float *grad_conf /* [B,P,C] opt */,
        const float *x_max_override /* opt: batch-global max of conf when this is one shard of a batch */,
        const int32_t *n_override /* opt: batch-global number of positives */,
        float *stats_out /* [2] opt: this batch's own (max of conf, number of positives) */) {
    size_t BP = (size_t)B * P;
    float *loc_t = (float *)malloc(sizeof(float) * 4 * BP);
    int64_t *conf_t = (int64_t *)malloc(sizeof(int64_t) * BP);
    float *key = (float *)malloc(sizeof(float) * BP);
    uint8_t *neg = (uint8_t *)calloc(BP, 1);
    int rc = GSSD_OK;
    for (int b = 0; b < B; ++b) {
        int G = gt_off[b + 1] - gt_off[b];
        if (G <= 0) rc = GSSD_ERR_EMPTY;
    }
    if (rc == GSSD_OK) {
#pragma omp parallel for schedule(dynamic)
        for (int b = 0; b < B; ++b) {                         /* 67-72 */
            int G = gt_off[b + 1] - gt_off[b];
            float *tr = (float *)malloc(sizeof(float) * 4 * G);
            float *lab = (float *)malloc(sizeof(float) * G);
            for (int g = 0; g < G; ++g) {
                const float *row = gt + 5 * (size_t)(gt_off[b] + g);
                memcpy(tr + 4 * g, row, 4 * sizeof(float)); lab[g] = row[4];
            }
            gssd_oracle_match(threshold, tr, lab, G, priors, P, v0, v1,
                              loc_t + 4 * (size_t)b * P, conf_t + (size_t)b * P, NULL, NULL);
            free(tr); free(lab);
        }
    }
    if (rc != GSSD_OK) goto done;
    {
        /* 80-88: smooth L1 over positives, summed */
        double loss_l = 0.0;
        long N = 0;
        for (size_t i = 0; i < BP; ++i) {
            if (conf_t[i] > 0) {
                ++N;
                for (int k = 0; k < 4; ++k) {
                    float d = loc[4 * i + k] - loc_t[4 * i + k];
                    float ad = fabsf(d);
                    loss_l += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
                }
            }
        }
        /* 91-99: mining key with the batch-global max, positives zeroed */
        float x_max = tensor_max(conf, BP * C);
        if (stats_out) { stats_out[0] = x_max; stats_out[1] = (float)N; }
        if (x_max_override) x_max = *x_max_override;
        if (n_override) N = *n_override;
#pragma omp parallel for
        for (size_t i = 0; i < BP; ++i) {
            float k = lse_row(conf + i * C, C, x_max) - conf[i * C + conf_t[i]];
            key[i] = conf_t[i] > 0 ? 0.f : k;
        }
        /* 102-106: rank = argsort(argsort(key, desc)); neg = rank < min(ratio*num_pos, P-1) */
#pragma omp parallel for schedule(dynamic)
        for (int b = 0; b < B; ++b) {
            key_idx *ki = (key_idx *)malloc(sizeof(key_idx) * P);
            int np = 0;
            for (int p = 0; p < P; ++p) { ki[p].key = key[(size_t)b * P + p]; ki[p].idx = p; np += conf_t[(size_t)b * P + p] > 0; }
            qsort(ki, P, sizeof(key_idx), cmp_key_desc);
            long nn = (long)negpos_ratio * np;
            if (nn > P - 1) nn = P - 1;
            for (long r = 0; r < nn; ++r) neg[(size_t)b * P + ki[r].idx] = 1;
            if (num_pos_out) num_pos_out[b] = np;
            if (kth_gap_out) kth_gap_out[b] = (nn > 0 && nn < P) ? ki[nn - 1].key - ki[nn].key : 1.f;
            free(ki);
        }
        /* 108-113: cross entropy (torch log_softmax: row max) over pos|neg, summed */
        double loss_c = 0.0;
        float inv_n = 1.f / (float)N;
#pragma omp parallel for reduction(+ : loss_c)
        for (size_t i = 0; i < BP; ++i) {
            int sel = conf_t[i] > 0 || neg[i];
            const float *x = conf + i * C;
            if (grad_conf) for (int c = 0; c < C; ++c) grad_conf[i * C + c] = 0.f;
            if (grad_loc) for (int k = 0; k < 4; ++k) grad_loc[4 * i + k] = 0.f;
            if (sel) {
                float m = x[0];
                for (int c = 1; c < C; ++c) if (x[c] > m) m = x[c];
                float s = 0.f;
                for (int c = 0; c < C; ++c) s += expf(x[c] - m);
                float ls = logf(s);
                loss_c += -((x[conf_t[i]] - m) - ls);
                if (grad_conf)
                    for (int c = 0; c < C; ++c) {
                        float sm = expf((x[c] - m) - ls);
                        grad_conf[i * C + c] = (sm - (c == conf_t[i] ? 1.f : 0.f)) * inv_n;
                    }
            }
            if (conf_t[i] > 0 && grad_loc)
                for (int k = 0; k < 4; ++k) {
                    float d = loc[4 * i + k] - loc_t[4 * i + k];
                    float g = fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
                    grad_loc[4 * i + k] = g * inv_n;
                }
        }
        losses[0] = (float)loss_l / (float)N;                /* 117-119 */
        losses[1] = (float)loss_c / (float)N;
    }
    if (loc_t_out) memcpy(loc_t_out, loc_t, sizeof(float) * 4 * BP);
    if (conf_t_out) memcpy(conf_t_out, conf_t, sizeof(int64_t) * BP);
    if (pos_out) for (size_t i = 0; i < BP; ++i) pos_out[i] = conf_t[i] > 0;
    if (neg_out) memcpy(neg_out, neg, BP);
    if (key_out) memcpy(key_out, key, sizeof(float) * BP);
done:
    free(loc_t); free(conf_t); free(key); free(neg);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* nms — box_utils.py:174-238 */
typedef struct { float s; int32_t idx; } score_idx;

static int cmp_score_asc(const void *a, const void *b) {       /* stable ascending */
    const score_idx *x = (const score_idx *)a, *y = (const score_idx *)b;
    if (x->s < y->s) return -1;
    if (x->s > y->s) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* keep[n] zero padded; returns count.  min_margin (opt, [2]): [0] = smallest |IoU - overlap| over all
 * suppression decisions taken (how close the keep list is to flipping under 1-ulp box noise),
 * [1] = score gap at the top_k cut (0 = an exact tie there: order undefined in the reference). */
EXPORT int gssd_oracle_nms(const float *boxes, const float *scores, int n, float overlap, int top_k,
                           int64_t *keep, float *min_margin) {
    for (int i = 0; i < n; ++i) keep[i] = 0;                  /* 186 */
    float margin = INFINITY, cut_gap = INFINITY;
    if (n <= 0) { if (min_margin) { min_margin[0] = margin; min_margin[1] = cut_gap; } return 0; }   /* 187-188 */
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    score_idx *si = (score_idx *)malloc(sizeof(score_idx) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        const float *b = boxes + 4 * i;
        area[i] = (b[2] - b[0]) * (b[3] - b[1]);              /* 193 */
        si[i].s = scores[i]; si[i].idx = i;
    }
    qsort(si, n, sizeof(score_idx), cmp_score_asc);           /* 194 */
    int m = n < top_k ? n : top_k;                            /* 196: idx[-top_k:] */
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
    for (int i = 0; i < m; ++i) idx[i] = si[n - m + i].idx;
    if (n > m) cut_gap = si[n - m].s - si[n - m - 1].s;
    int count = 0;
    while (m > 0) {                                           /* 206 */
        int i = idx[m - 1];                                   /* 207 */
        keep[count++] = i;                                    /* 209-210 */
        if (m == 1) break;                                    /* 211-212 */
        --m;                                                  /* 213 */
        const float *bi = boxes + 4 * i;
        int w_out = 0;
        for (int q = 0; q < m; ++q) {
            const float *bj = boxes + 4 * idx[q];
            float xx1 = bj[0] < bi[0] ? bi[0] : bj[0];        /* 220-223: clamp to box i */
            float yy1 = bj[1] < bi[1] ? bi[1] : bj[1];
            float xx2 = bj[2] > bi[2] ? bi[2] : bj[2];
            float yy2 = bj[3] > bi[3] ? bi[3] : bj[3];
            float w = clamp0(xx2 - xx1), h = clamp0(yy2 - yy1);   /* 226-230 */
            float inter = w * h;                              /* 231 */
            float uni = (area[idx[q]] - inter) + area[i];     /* 233-234 */
            float iou = inter / uni;                          /* 235 */
            float d = fabsf(iou - overlap);
            if (d < margin) margin = d;
            if (iou <= overlap) idx[w_out++] = idx[q];        /* 237 */
        }
        m = w_out;
    }
    free(area); free(si); free(idx);
    if (min_margin) { min_margin[0] = margin; min_margin[1] = cut_gap; }
    return count;
}

/* ------------------------------------------------------------------------------------------ */
/* Detect.forward — layers/functions/detection_pytorch_ver_1point5.py:33-89 */
EXPORT int gssd_oracle_detect(const float *loc, const float *conf, const float *priors,
                              int B, int P, int C, int top_k, float conf_thresh, float nms_thresh,
                              float v0, float v1, float *out /* [B,C,top_k,5] */,
                              int32_t *count_out /* [B,C] opt */, int32_t *keep_idx_out /* [B,C,top_k] opt */,
                              float *min_margin /* [B,C] opt: IoU margin */, float *cut_gap /* [B,C] opt */) {
    if (nms_thresh <= 0) return GSSD_ERR_VALUE;               /* 39-40 */
    memset(out, 0, sizeof(float) * (size_t)B * C * top_k * 5);    /* 56 */
    if (keep_idx_out) for (size_t i = 0; i < (size_t)B * C * top_k; ++i) keep_idx_out[i] = -1;
    if (count_out) memset(count_out, 0, sizeof(int32_t) * (size_t)B * C);
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {                             /* 62 */
        float *dec = (float *)malloc(sizeof(float) * 4 * P);
        float *bx = (float *)malloc(sizeof(float) * 4 * P);
        float *sc = (float *)malloc(sizeof(float) * P);
        int32_t *orig = (int32_t *)malloc(sizeof(int32_t) * P);
        int64_t *keep = (int64_t *)malloc(sizeof(int64_t) * P);
        gssd_oracle_decode(loc + 4 * (size_t)b * P, priors, P, v0, v1, dec);   /* 63 */
        for (int cl = 1; cl < C; ++cl) {                      /* 67 */
            int n = 0;
            for (int p = 0; p < P; ++p) {
                float s = conf[((size_t)b * P + p) * C + cl];
                if (s > conf_thresh) {                        /* 69-70, 76-78: order-preserving */
                    sc[n] = s; memcpy(bx + 4 * n, dec + 4 * p, 4 * sizeof(float)); orig[n] = p; ++n;
                }
            }
            float mg[2] = {INFINITY, INFINITY};
            if (min_margin) min_margin[b * C + cl] = mg[0];
            if (cut_gap) cut_gap[b * C + cl] = mg[1];
            if (n == 0) continue;                             /* 73-75 */
            int cnt = gssd_oracle_nms(bx, sc, n, nms_thresh, top_k, keep, mg);   /* 81 */
            float *o = out + (((size_t)b * C + cl) * top_k) * 5;
            for (int r = 0; r < cnt; ++r) {                   /* 82-84 */
                o[5 * r] = sc[keep[r]];
                memcpy(o + 5 * r + 1, bx + 4 * keep[r], 4 * sizeof(float));
                if (keep_idx_out) keep_idx_out[((size_t)b * C + cl) * top_k + r] = orig[keep[r]];
            }
            if (count_out) count_out[b * C + cl] = cnt;
            if (min_margin) min_margin[b * C + cl] = mg[0];
            if (cut_gap) cut_gap[b * C + cl] = mg[1];
        }
        free(dec); free(bx); free(sc); free(orig); free(keep);
    }
    /* 85-88: the trailing cross-class top-k mutates a temporary copy: no effect. */
    return GSSD_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* L2Norm.forward — layers/modules/l2norm.py:19-23; x[B,Cn,HW] */
EXPORT void gssd_oracle_l2norm(const float *x, const float *w, int B, int Cn, int HW, float eps, float *y) {
    for (int b = 0; b < B; ++b) for (int i = 0; i < HW; ++i) {
        double s = 0.0;
        for (int c = 0; c < Cn; ++c) { float v = x[((size_t)b * Cn + c) * HW + i]; s += (double)v * v; }
        float norm = sqrtf((float)s) + eps;
        for (int c = 0; c < Cn; ++c) {
            size_t k = ((size_t)b * Cn + c) * HW + i;
            y[k] = w[c] * (x[k] / norm);
        }
    }
}
