// attn.cu — the attention core of GSSD++'s Self_Attn (layers/self_attn.py:69-81 of the reference):
//
//     attn   = softmax(theta^T phi, dim = -1)         theta [B, D, N], phi [B, D, M]  ->  attn [B, N, M]     (:71-72)
//     attn_g = g attn^T                               g [B, Cv, M]                    ->  attn_g [B, Cv, N]  (:80)
//
// (N = H*W queries, M = pooled keys, D = C/8, Cv = C/2; the 1x1 convolutions and the pooling around it stay the module's.)
// The reference runs permute + bmm + softmax + permute + bmm and their five backward nodes; here the forward is ONE kernel
// and the backward two, fp32 throughout (the module returns the attention map itself, self_attn.py:86, so it is materialised
// once, by the kernel that computes it, and re-used by the backward).
//
// One CTA owns a strip of QT queries against ALL keys in shared memory (32 x 1444 floats = 185 KB at the 38 x 38 map):
// scores, softmax and the product with g never leave the SM.  Both GEMM shapes are register-tiled fp32 FMA loops fed from
// shared memory with the next operand tile prefetched into registers: the whole attention is 1.3 GFLOP per image (1 % of the
// model), latency- and launch-bound in the reference, not a tensor-core problem.
#include "common.cuh"

namespace gssd {

constexpr int AT_NT = 256;
constexpr int AT_DC = 64;            // contraction chunk of strip_gemm
constexpr int AT_KA = 128;           // keys per tile of strip_gemm (16 per warp)
constexpr int AT_TAP = AT_KA + 8;    // its row pitch: 8 mod 32, so the A fragments (4 rows x 8 columns per load) hit 32 banks
// row pitch of the query tile: 8 mod 32 (B fragments: 4 rows x 8 columns per load on 32 banks; QT + 8 made it 32 at 24 queries:
// 4-way conflicts, 22 % of the backward's samples on that load)
__host__ __device__ constexpr int at_qsp(int qt) { return (qt - 8 + 31) / 32 * 32 + 8; }
constexpr int AT_KC = 64;            // keys per tile of strip_apply: few, large tiles — with 16 keys per tile every one of the 91 steps
                                     // of a 1444-key strip exposed an L2 round trip (ncu: 44 % of the samples on the tile loads)

// Both GEMM shapes run on the tensor cores as TF32 mma.sync.m16n8k8 (fp32 accumulation): a first fp32-FMA version, register-tiled
// from shared memory, reached 13 - 16 TFLOP/s, half of cuBLAS's SIMT kernels (ncu: 42 % of the issue slots, 8 or 16 warps per SM
// waiting on shared-memory operands).  Operands are rounded to TF32 (10-bit mantissa) when they are stored into shared memory.
__device__ __forceinline__ float tf32r(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// strip[q][m] (+)= sum_{d < Dn} A[d*lda + q0 + q] * Bm[d*ldb + m]     for q < QT (zero rows beyond nq), m < L
// qs: [AT_DC][at_qsp(QT)] floats, tile: [AT_DC][AT_TAP] floats.  MMA roles: M = 16 keys (one m-tile per warp and key tile), N = 8
// queries, K = 8 of the contraction.  Every (q, m) of the strip belongs to one thread, so the read-modify-write over the chunks
// of the contraction needs no synchronisation of its own.
template <int QT>
__device__ __forceinline__ void strip_gemm(float *strip, int Lp, int L, const float *__restrict__ A, int lda, int q0, int nq,
                                           const float *__restrict__ Bm, int ldb, int Dn, float *qs, float *tile) {
    constexpr int NT = QT / 8, QSP = at_qsp(QT), TPT = AT_DC * AT_KA / AT_NT;
    static_assert(TPT % 4 == 0, "tile shape");
    const int tid = threadIdx.x, warp = tid >> 5, grp = (tid & 31) >> 2, tig = tid & 3;
    const bool vec = (ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(Bm) & 15) == 0;
    for (int d0 = 0; d0 < Dn; d0 += AT_DC) {
        const int dn = min(AT_DC, Dn - d0);
        __syncthreads();                                               // qs / tile of the previous chunk are no longer read
        for (int e = tid; e < AT_DC * QT; e += AT_NT) {
            const int dd = e / QT, q = e - dd * QT;
            qs[dd * QSP + q] = (dd < dn && q < nq) ? tf32r(A[(size_t)(d0 + dd) * lda + q0 + q]) : 0.f;
        }
        // tile loads: 16-byte vectors when the rows allow it (row length and base a multiple of 4 floats: the 38 x 38 and 10 x 10
        // maps), scalars otherwise; a thread's elements are the same in both forms: float4 index f = tid + i*256 of the tile
        float pre[TPT];
        auto fetch = [&](int k0) {
#pragma unroll
            for (int i = 0; i < TPT / 4; ++i) {
                const int f = tid + i * AT_NT, dd = f / (AT_KA / 4), kk = (f - dd * (AT_KA / 4)) * 4;
                const float *src = Bm + (size_t)(d0 + dd) * ldb + k0 + kk;
                if (vec) {
                    const float4 v = (dd < dn && k0 + kk < L) ? __ldg(reinterpret_cast<const float4 *>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    pre[4 * i] = v.x; pre[4 * i + 1] = v.y; pre[4 * i + 2] = v.z; pre[4 * i + 3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) pre[4 * i + j] = (dd < dn && k0 + kk + j < L) ? src[j] : 0.f;
                }
            }
        };
        fetch(0);
        for (int k0 = 0; k0 < L; k0 += AT_KA) {
            __syncthreads();                                           // the previous tile has been consumed (and qs is written)
#pragma unroll
            for (int i = 0; i < TPT / 4; ++i) {
                const int f = tid + i * AT_NT, dd = f / (AT_KA / 4), kk = (f - dd * (AT_KA / 4)) * 4;
                *reinterpret_cast<float4 *>(tile + dd * AT_TAP + kk) =
                    make_float4(tf32r(pre[4 * i]), tf32r(pre[4 * i + 1]), tf32r(pre[4 * i + 2]), tf32r(pre[4 * i + 3]));
            }
            __syncthreads();
            if (k0 + AT_KA < L) fetch(k0 + AT_KA);
            if (k0 + warp * 16 >= L) continue;                         // this warp's 16 keys lie beyond the end
            float acc[NT][4];
#pragma unroll
            for (int n = 0; n < NT; ++n) { acc[n][0] = 0.f; acc[n][1] = 0.f; acc[n][2] = 0.f; acc[n][3] = 0.f; }
#pragma unroll
            for (int d8 = 0; d8 < AT_DC; d8 += 8) {
                const float *tp = tile + (d8 + tig) * AT_TAP + warp * 16 + grp;
                const float a[4] = {tp[0], tp[8], tp[4 * AT_TAP], tp[4 * AT_TAP + 8]};
                const float *qp = qs + (d8 + tig) * QSP + grp;
#pragma unroll
                for (int n = 0; n < NT; ++n) mma_tf32(acc[n], a, qp[n * 8], qp[4 * QSP + n * 8]);
            }
#pragma unroll
            for (int n = 0; n < NT; ++n) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m = k0 + warp * 16 + grp + (i >> 1) * 8, q = n * 8 + 2 * tig + (i & 1);
                    if (m < Lp) {
                        float *p = strip + (size_t)q * Lp + m;
                        *p = (d0 ? *p : 0.f) + (m < L ? acc[n][i] : 0.f);
                    }
                }
            }
        }
    }
    __syncthreads();
}

// out[(c)*ldo + q0 + q] = sum_{m < L} strip[q][m] * Mat[c*ldm + m]      for c < CH, q < nq
// CHP channels per pass.  MMA roles: M = 16 channels, N = 8 queries, K = 8 keys; the CHP/16 x QT/8 MMA tiles are dealt to the 8
// warps as MW x NW blocks.  tile: [CHP][AT_KC + 4] floats (channel-major, as it is loaded).  The strip holds TF32-rounded values; its padding
// columns (L..Lp) must be zero; reads up to 12 floats past a row's end meet zeros of the tile.
template <int QT, int CHP>
__device__ __forceinline__ void strip_apply(const float *strip, int Lp, int L, const float *__restrict__ Mat, int ldm, int CH,
                                            float *__restrict__ out, int ldo, int q0, int nq, float *tile) {
    constexpr int MTT = CHP / 16, NTT = QT / 8, WC = MTT < 8 ? MTT : 8, WQ = 8 / WC, MW = MTT / WC, NW = (NTT + WQ - 1) / WQ;
    constexpr int NF4 = CHP * AT_KC / 4, IT = (NF4 + AT_NT - 1) / AT_NT, TP = AT_KC + 4;     // float4s of a tile, per thread; row pitch:
    // 4 mod 32, so the A fragments (8 rows x 4 columns per load) hit 32 banks and the rows take 16-byte stores as they were loaded
    static_assert(MW >= 1, "tile shape");
    const int tid = threadIdx.x, warp = tid >> 5, grp = (tid & 31) >> 2, tig = tid & 3;
    const bool vec = (ldm & 3) == 0 && (reinterpret_cast<uintptr_t>(Mat) & 15) == 0;
    const int cb = (warp % WC) * MW * 16, qb = (warp / WC) * NW * 8;
    const bool active = qb < QT;
    for (int c0 = 0; c0 < CH; c0 += CHP) {
        float acc[MW][NW][4];
#pragma unroll
        for (int mt = 0; mt < MW; ++mt)
#pragma unroll
            for (int n = 0; n < NW; ++n) { acc[mt][n][0] = 0.f; acc[mt][n][1] = 0.f; acc[mt][n][2] = 0.f; acc[mt][n][3] = 0.f; }
        float pre[4 * IT];
        auto fetch = [&](int k0) {
#pragma unroll
            for (int i = 0; i < IT; ++i) {
                const int f = tid + i * AT_NT, cc = f / (AT_KC / 4), kk = (f - cc * (AT_KC / 4)) * 4;
                if (f >= NF4) break;
                const float *src = Mat + (size_t)(c0 + cc) * ldm + k0 + kk;
                if (vec) {
                    const float4 v = (k0 + kk < L) ? __ldg(reinterpret_cast<const float4 *>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    pre[4 * i] = v.x; pre[4 * i + 1] = v.y; pre[4 * i + 2] = v.z; pre[4 * i + 3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) pre[4 * i + j] = (k0 + kk + j < L) ? src[j] : 0.f;
                }
            }
        };
        fetch(0);
        for (int k0 = 0; k0 < L; k0 += AT_KC) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < IT; ++i) {
                const int f = tid + i * AT_NT, cc = f / (AT_KC / 4), kk = (f - cc * (AT_KC / 4)) * 4;
                if (f >= NF4) break;
                *reinterpret_cast<float4 *>(tile + cc * TP + kk) =
                    make_float4(tf32r(pre[4 * i]), tf32r(pre[4 * i + 1]), tf32r(pre[4 * i + 2]), tf32r(pre[4 * i + 3]));
            }
            __syncthreads();
            if (k0 + AT_KC < L) fetch(k0 + AT_KC);
            if (active) {
#pragma unroll
                for (int ks = 0; ks < AT_KC; ks += 8) {
                    if (k0 + ks >= L) break;
                    float b[NW][2];
#pragma unroll
                    for (int n = 0; n < NW; ++n) {
                        const float *sp = strip + (size_t)(qb + n * 8 + grp) * Lp + k0 + ks + tig;
                        const bool on = qb + n * 8 < QT;                 // QT = 24: the last warp group owns one n-tile fewer
                        b[n][0] = on ? sp[0] : 0.f; b[n][1] = on ? sp[4] : 0.f;
                    }
#pragma unroll
                    for (int mt = 0; mt < MW; ++mt) {
                        const float *tp = tile + (cb + mt * 16 + grp) * TP + ks + tig;
                        const float a[4] = {tp[0], tp[8 * TP], tp[4], tp[8 * TP + 4]};
#pragma unroll
                        for (int n = 0; n < NW; ++n) mma_tf32(acc[mt][n], a, b[n][0], b[n][1]);
                    }
                }
            }
        }
        if (active) {
#pragma unroll
            for (int mt = 0; mt < MW; ++mt)
#pragma unroll
                for (int n = 0; n < NW; ++n)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c0 + cb + mt * 16 + grp + (i >> 1) * 8, q = qb + n * 8 + 2 * tig + (i & 1);
                        if (q < nq && q < QT) out[(size_t)c * ldo + q0 + q] = acc[mt][n][i];
                    }
        }
        __syncthreads();
    }
}

template <int QT>
__device__ __forceinline__ void apply_any(const float *strip, int Lp, int L, const float *Mat, int ldm, int CH, float *out, int ldo,
                                          int q0, int nq, float *tile) {
    if (CH % 256 == 0) strip_apply<QT, 256>(strip, Lp, L, Mat, ldm, CH, out, ldo, q0, nq, tile);
    else if (CH % 128 == 0) strip_apply<QT, 128>(strip, Lp, L, Mat, ldm, CH, out, ldo, q0, nq, tile);
    else if (CH % 64 == 0) strip_apply<QT, 64>(strip, Lp, L, Mat, ldm, CH, out, ldo, q0, nq, tile);
    else strip_apply<QT, 32>(strip, Lp, L, Mat, ldm, CH, out, ldo, q0, nq, tile);
}

struct AttnArgs {
    const float *theta, *phi, *g;      // [B, D, N], [B, D, M], [B, Cv, M]
    float *attn, *o;                   // [B, N, M], [B, Cv, N]
    const float *d_o;                  // backward: [B, Cv, N]
    float *ds;                         // backward scratch [B, N, M]
    float *d_theta, *d_phi, *d_g;      // backward outputs, shapes of theta / phi / g
    int D, Cv, N, M;
};

static size_t attn_smem_floats(int QT, int L, int D) {
    const int Lp = (L + 3) & ~3;
    size_t tile = (size_t)AT_DC * AT_TAP;
    if ((size_t)256 * (AT_KC + 4) > tile) tile = (size_t)256 * (AT_KC + 4);
    (void)D;
    return (size_t)QT * Lp + 16 + (size_t)AT_DC * at_qsp(QT) + tile;      // + 16: reads past the last row's end stay inside
}

// forward: scores -> softmax (attn written once) -> attn_g
template <int QT>
__global__ void __launch_bounds__(AT_NT, 1) attn_fwd_kernel(AttnArgs a) {
    extern __shared__ __align__(16) float at_smem[];
    const int b = blockIdx.y, q0 = blockIdx.x * QT, nq = min(QT, a.N - q0), Lp = (a.M + 3) & ~3;
    float *strip = at_smem, *qs = strip + (size_t)QT * Lp + 16, *tile = qs + AT_DC * at_qsp(QT);
    if (threadIdx.x < 16) strip[(size_t)QT * Lp + threadIdx.x] = 0.f;      // B fragments of the last row read up to 7 floats past its end
    strip_gemm<QT>(strip, Lp, a.M, a.theta + (size_t)b * a.D * a.N, a.N, q0, nq, a.phi + (size_t)b * a.D * a.M, a.M, a.D, qs, tile);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < QT; r += AT_NT / 32) {
        float *row = strip + (size_t)r * Lp;
        if (r >= nq) {
            for (int m = lane; m < Lp; m += 32) row[m] = 0.f;
            continue;
        }
        float mx = -INFINITY;
        for (int m = lane; m < a.M; m += 32) mx = fmaxf(mx, row[m]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int m = lane; m < a.M; m += 32) { const float e = expf(row[m] - mx); row[m] = e; sum += e; }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        const float inv = 1.f / sum;
        float *dst = a.attn + ((size_t)b * a.N + q0 + r) * a.M;
        for (int m = lane; m < a.M; m += 32) { const float p = row[m] * inv; row[m] = tf32r(p); dst[m] = p; }
    }
    __syncthreads();
    apply_any<QT>(strip, Lp, a.M, a.g + (size_t)b * a.Cv * a.M, a.M, a.Cv, a.o + (size_t)b * a.Cv * a.N, a.N, q0, nq, tile);
}

// backward, query strips: dP = dO^T g -> dS = P * (dP - rowsum(dP * P)) (written for the key-side kernel) -> d_theta = phi dS^T
template <int QT>
__global__ void __launch_bounds__(AT_NT, 1) attn_bwd_q_kernel(AttnArgs a) {
    extern __shared__ __align__(16) float at_smem[];
    const int b = blockIdx.y, q0 = blockIdx.x * QT, nq = min(QT, a.N - q0), Lp = (a.M + 3) & ~3;
    float *strip = at_smem, *qs = strip + (size_t)QT * Lp + 16, *tile = qs + AT_DC * at_qsp(QT);
    if (threadIdx.x < 16) strip[(size_t)QT * Lp + threadIdx.x] = 0.f;      // B fragments of the last row read up to 7 floats past its end
    strip_gemm<QT>(strip, Lp, a.M, a.d_o + (size_t)b * a.Cv * a.N, a.N, q0, nq, a.g + (size_t)b * a.Cv * a.M, a.M, a.Cv, qs, tile);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < QT; r += AT_NT / 32) {
        float *row = strip + (size_t)r * Lp;
        if (r >= nq) {
            for (int m = lane; m < Lp; m += 32) row[m] = 0.f;
            continue;
        }
        const float *p = a.attn + ((size_t)b * a.N + q0 + r) * a.M;
        float dot = 0.f;
        for (int m = lane; m < a.M; m += 32) dot += row[m] * p[m];
#pragma unroll
        for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
        float *dst = a.ds + ((size_t)b * a.N + q0 + r) * a.M;
        for (int m = lane; m < a.M; m += 32) { const float v = p[m] * (row[m] - dot); row[m] = tf32r(v); dst[m] = v; }
    }
    __syncthreads();
    apply_any<QT>(strip, Lp, a.M, a.phi + (size_t)b * a.D * a.M, a.M, a.D, a.d_theta + (size_t)b * a.D * a.N, a.N, q0, nq, tile);
}

// backward, key strips: d_phi = theta dS, d_g = dO P — the transposed strips [key][query] are read from dS / attn
template <int QT>
__global__ void __launch_bounds__(AT_NT, 1) attn_bwd_k_kernel(AttnArgs a) {
    extern __shared__ __align__(16) float at_smem[];
    const int b = blockIdx.y, m0 = blockIdx.x * QT, nk = min(QT, a.M - m0), Lp = (a.N + 3) & ~3;
    float *strip = at_smem, *tile = strip + (size_t)QT * Lp + 16 + AT_DC * at_qsp(QT);
    if (threadIdx.x < 16) strip[(size_t)QT * Lp + threadIdx.x] = 0.f;
    // blockIdx.z picks the product: 0 = d_phi from dS, 1 = d_g from the attention map.  Two CTAs per key strip instead of one that
    // does both: at 4 images 244 CTAs were two rounds on 148 SMs, the second two thirds empty; 488 half-size ones pack better
    const int pass = blockIdx.z;
    const float *src = (pass == 0 ? a.ds : a.attn) + (size_t)b * a.N * a.M + m0;
    __syncthreads();
    for (int e = threadIdx.x; e < QT * Lp; e += AT_NT) {
        const int q = e / QT, kl = e - q * QT;                        // consecutive threads: consecutive keys of one query row
        strip[(size_t)kl * Lp + q] = (q < a.N && kl < nk) ? tf32r(src[(size_t)q * a.M + kl]) : 0.f;
    }
    __syncthreads();
    if (pass == 0)
        apply_any<QT>(strip, Lp, a.N, a.theta + (size_t)b * a.D * a.N, a.N, a.D, a.d_phi + (size_t)b * a.D * a.M, a.M, m0, nk, tile);
    else
        apply_any<QT>(strip, Lp, a.N, a.d_o + (size_t)b * a.Cv * a.N, a.N, a.Cv, a.d_g + (size_t)b * a.Cv * a.M, a.M, m0, nk, tile);
}

static int attn_check(const AttnArgs &a, int B) {
    if (B <= 0 || a.D <= 0 || a.Cv <= 0 || a.N <= 0 || a.M <= 0) return GSSD_ERR_ARG;
    if (a.D % 32 || a.Cv % 32 || B > 65535) return GSSD_ERR_LIMIT;
    return GSSD_OK;
}

// Strip height: the kernels are bound by the L2 traffic of re-reading phi / g for every strip (8 flop per byte at 16 queries:
// 3.6 TB/s measured), so the tallest strip that fits beside 70 KB of operand tiles: 24 queries up to 1500 keys (139 KB of
// scores), then 16 and 8.
static int attn_pick_qt(int L, int D) {
    for (int qt : {24, 16, 8})
        if (attn_smem_floats(qt, L, D) * sizeof(float) <= 225 * 1024) return qt;
    return 0;
}

template <typename K>
static int attn_launch(K kern, const AttnArgs &a, int B, int tiles, size_t smem, cudaStream_t st, int nz = 1) {
    GSSD_RETURN_IF_CUDA(allow_max_smem(reinterpret_cast<const void *>(kern)));
    kern<<<dim3(tiles, B, nz), AT_NT, smem, st>>>(a);
    GSSD_AFTER_LAUNCH();
    return GSSD_OK;
}

}  // namespace gssd

using namespace gssd;

extern "C" int gssd_attn_fwd(const float *theta, const float *phi, const float *g, int B, int D, int Cv, int N, int M, float *attn,
                             float *attn_g, void *stream) {
    if (!theta || !phi || !g || !attn || !attn_g) return GSSD_ERR_ARG;
    AttnArgs a = {};
    a.theta = theta; a.phi = phi; a.g = g; a.attn = attn; a.o = attn_g; a.D = D; a.Cv = Cv; a.N = N; a.M = M;
    int rc = attn_check(a, B);
    if (rc) return rc;
    const int qt = attn_pick_qt(M, D);
    if (!qt) return GSSD_ERR_LIMIT;
    const size_t smem = attn_smem_floats(qt, M, D) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    switch (qt) {
        case 24: return attn_launch(attn_fwd_kernel<24>, a, B, ceil_div(N, 24), smem, st);
        case 16: return attn_launch(attn_fwd_kernel<16>, a, B, ceil_div(N, 16), smem, st);
        default: return attn_launch(attn_fwd_kernel<8>, a, B, ceil_div(N, 8), smem, st);
    }
}

extern "C" int gssd_attn_bwd(const float *theta, const float *phi, const float *g, const float *attn, const float *d_attn_g, int B,
                             int D, int Cv, int N, int M, float *d_theta, float *d_phi, float *d_g, float *ds_ws, void *stream) {
    if (!theta || !phi || !g || !attn || !d_attn_g || !d_theta || !d_phi || !d_g || !ds_ws) return GSSD_ERR_ARG;
    AttnArgs a = {};
    a.theta = theta; a.phi = phi; a.g = g; a.attn = const_cast<float *>(attn); a.d_o = d_attn_g; a.ds = ds_ws;
    a.d_theta = d_theta; a.d_phi = d_phi; a.d_g = d_g; a.D = D; a.Cv = Cv; a.N = N; a.M = M;
    int rc = attn_check(a, B);
    if (rc) return rc;
    const int qt = attn_pick_qt(M, D), kt = attn_pick_qt(N, D);
    if (!qt || !kt) return GSSD_ERR_LIMIT;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem_q = attn_smem_floats(qt, M, D) * sizeof(float), smem_k = attn_smem_floats(kt, N, D) * sizeof(float);
    switch (qt) {
        case 24: rc = attn_launch(attn_bwd_q_kernel<24>, a, B, ceil_div(N, 24), smem_q, st); break;
        case 16: rc = attn_launch(attn_bwd_q_kernel<16>, a, B, ceil_div(N, 16), smem_q, st); break;
        default: rc = attn_launch(attn_bwd_q_kernel<8>, a, B, ceil_div(N, 8), smem_q, st); break;
    }
    if (rc) return rc;
    switch (kt) {
        case 24: return attn_launch(attn_bwd_k_kernel<24>, a, B, ceil_div(M, 24), smem_k, st, 2);
        case 16: return attn_launch(attn_bwd_k_kernel<16>, a, B, ceil_div(M, 16), smem_k, st, 2);
        default: return attn_launch(attn_bwd_k_kernel<8>, a, B, ceil_div(M, 8), smem_k, st, 2);
    }
}
