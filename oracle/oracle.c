/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C (OpenMP) CPU oracle for the two hot loops of the reference, written so that sizes the literal
 * numpy restatement cannot hold (N >= 16k: the reference materialises ~13 N x N tensors) can still be
 * checked, and so that bench.py has a CPU baseline ("port") that uses every host core.
 *
 *   oracle_supcon_*   SupConLoss1 label / SimCLR path, contrastyou/losses/contrastive.py:14-20, :51-100,
 *                     in the closed form of SURVEY.md Appendix A1 (verified against the reference):
 *                       loss = -(1/N) sum_i [ (sum_j P_ij S_ij)/c_i - (m + log(D_i + 1e-16)) ]
 *                       dZ   = (1/t) W Z,  W_ij = (1/N)[ -P_ij(1/c_i+1/c_j) + E_ij(1/(D_i+1e-16)+1/(D_j+1e-16)) ]
 *   oracle_iic_*      IIDSegmentationLoss, contrastyou/losses/discreteMI.py:139-165, :225-261
 *                     (shifted contraction == the F.conv2d call at :229-232; epilogue; adjoint).
 *
 * Pinned (tests/test_oracle.py) against the numpy oracles in this directory and through them against the golden fixtures generated
 * from the reference itself.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  Parity status: PINNED.
 *
 * Build: oracle/Makefile  (gcc -O3 -fopenmp -shared -fPIC; no -march so the .so runs on any x86-64 host;
 * hot loops are multi-versioned with target_clones).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define MULTIVERSION __attribute__((target_clones("avx512f", "avx2,fma", "default")))
#else
#define MULTIVERSION
#endif

/* torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU baseline sets its thread count explicitly (bench.py) */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------- dot / axpy kernels (float storage; float or double accumulation) ---------- */
MULTIVERSION static float dot_f32(const float* a, const float* b, int64_t d) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int64_t k = 0;
    for (; k + 64 <= d; k += 64) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma omp simd reduction(+ : t0, t1, t2, t3)
        for (int u = 0; u < 16; ++u) {
            t0 += a[k + u] * b[k + u];
            t1 += a[k + 16 + u] * b[k + 16 + u];
            t2 += a[k + 32 + u] * b[k + 32 + u];
            t3 += a[k + 48 + u] * b[k + 48 + u];
        }
        s0 += t0; s1 += t1; s2 += t2; s3 += t3;
    }
    for (; k < d; ++k) s0 += a[k] * b[k];
    return (s0 + s1) + (s2 + s3);
}

MULTIVERSION static double dot_f64(const float* a, const float* b, int64_t d) {
    double s = 0.0;
#pragma omp simd reduction(+ : s)
    for (int64_t k = 0; k < d; ++k) s += (double)a[k] * (double)b[k];
    return s;
}

MULTIVERSION static void axpy_f32(float w, const float* x, float* acc, int64_t d) {
    for (int64_t k = 0; k < d; ++k) acc[k] += w * x[k];
}

MULTIVERSION static void axpy_f64(double w, const float* x, double* acc, int64_t d) {
    for (int64_t k = 0; k < d; ++k) acc[k] += w * (double)x[k];
}

/* 4 rows against one column row: the column row is loaded once for four dot products / four axpys */
MULTIVERSION static void dot4_f32(const float* a0, const float* a1, const float* a2, const float* a3, const float* b,
                                  int64_t d, float out[4]) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma omp simd reduction(+ : s0, s1, s2, s3)
    for (int64_t k = 0; k < d; ++k) {
        const float bv = b[k];
        s0 += a0[k] * bv; s1 += a1[k] * bv; s2 += a2[k] * bv; s3 += a3[k] * bv;
    }
    out[0] = s0; out[1] = s1; out[2] = s2; out[3] = s3;
}

MULTIVERSION static void axpy4_f32(const float w[4], const float* x, float* c0, float* c1, float* c2, float* c3, int64_t d) {
    const float w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    for (int64_t k = 0; k < d; ++k) {
        const float xv = x[k];
        c0[k] += w0 * xv; c1[k] += w1 * xv; c2[k] += w2 * xv; c3[k] += w3 * xv;
    }
}

/* float32 fast form of oracle_supcon_fwd_bwd (prec == 0, N % 4 == 0): same arithmetic, rows processed four at a time.
 * This is the leg bench.py times as the CPU baseline. */
static int supcon_fwd_bwd_f32_blocked(const float* z, const int32_t* labels, int64_t N, int64_t d, double t, double* loss,
                                      double* row_lse, double* row_cnt, float* dz) {
    const float inv_t = (float)(1.0 / t);
    double* D = (double*)malloc(sizeof(double) * (size_t)N);
    double* ps = (double*)malloc(sizeof(double) * (size_t)N);
    if (!D || !ps) return -2;
    float m = -INFINITY;
#pragma omp parallel for schedule(static) reduction(max : m)
    for (int64_t i = 0; i < N; ++i) {
        const float s = dot_f32(z + i * d, z + i * d, d) * inv_t;
        if (s > m) m = s;
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < N; i += 4) {
        const float* a0 = z + i * d;
        double Dv[4] = {0, 0, 0, 0}, pv[4] = {0, 0, 0, 0}, cv[4] = {0, 0, 0, 0};
        for (int64_t j = 0; j < N; ++j) {
            float s[4];
            dot4_f32(a0, a0 + d, a0 + 2 * d, a0 + 3 * d, z + j * d, d, s);
            for (int r = 0; r < 4; ++r) {
                if (j == i + r) continue;
                const float sv = s[r] * inv_t;
                Dv[r] += (double)expf(sv - m);
                if (labels[j] == labels[i + r]) { pv[r] += (double)sv; cv[r] += 1.0; }
            }
        }
        for (int r = 0; r < 4; ++r) {
            D[i + r] = Dv[r]; ps[i + r] = pv[r]; row_cnt[i + r] = cv[r]; row_lse[i + r] = (double)m + log(Dv[r] + 1e-16);
        }
    }
    double acc = 0.0;
    for (int64_t i = 0; i < N; ++i) acc += ps[i] / row_cnt[i] - row_lse[i];
    *loss = -acc / (double)N;
    if (dz) {
        memset(dz, 0, sizeof(float) * (size_t)N * (size_t)d);
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t i = 0; i < N; i += 4) {
            const float* a0 = z + i * d;
            float* c0 = dz + i * d;
            float invD[4], invc[4];
            for (int r = 0; r < 4; ++r) { invD[r] = (float)(1.0 / (D[i + r] + 1e-16)); invc[r] = (float)(1.0 / row_cnt[i + r]); }
            const float scale = inv_t / (float)N;
            for (int64_t j = 0; j < N; ++j) {
                float s[4], w[4];
                dot4_f32(a0, a0 + d, a0 + 2 * d, a0 + 3 * d, z + j * d, d, s);
                const float invDj = (float)(1.0 / (D[j] + 1e-16)), invcj = (float)(1.0 / row_cnt[j]);
                for (int r = 0; r < 4; ++r) {
                    float wv = 0.f;
                    if (j != i + r) {
                        wv = expf(s[r] * inv_t - m) * (invD[r] + invDj);
                        if (labels[j] == labels[i + r]) wv -= invc[r] + invcj;
                    }
                    w[r] = wv * scale;
                }
                axpy4_f32(w, z + j * d, c0, c0 + d, c0 + 2 * d, c0 + 3 * d, d);
            }
        }
    }
    free(D); free(ps);
    return 0;
}

/* ---------- SupCon, label path ----------
 * z       [N, d] float32, both views stacked (rows 0..n-1 view 1, n..2n-1 view 2)
 * labels  [N] int32, already tiled over the two views (label[i+n] == label[i]); SimCLR = arange(n) tiled
 * prec    0: float32 dot products (what the reference's fp32 torch.mm does), 1: float64 (ground truth)
 * outputs loss[1], row_lse[N] (= m + log(D_i+1e-16)), row_cnt[N] (= c_i), dz[N, d] (float32) -- dz may be NULL
 */
int oracle_supcon_fwd_bwd(const float* z, const int32_t* labels, int64_t N, int64_t d, double t, int prec,
                          double* loss, double* row_lse, double* row_cnt, float* dz) {
    if (N <= 0 || d <= 0) return -1;
    if (prec == 0 && (N % 4) == 0) return supcon_fwd_bwd_f32_blocked(z, labels, N, d, t, loss, row_lse, row_cnt, dz);
    const double inv_t = 1.0 / t;
    double* D = (double*)malloc(sizeof(double) * (size_t)N);
    double* ps = (double*)malloc(sizeof(double) * (size_t)N);
    if (!D || !ps) return -2;
    /* the reference shifts by the global max of S (contrastive.py:17-18): the largest squared row norm / t,
       unless two distinct rows are more similar than that (not for unit rows) -- take the true max */
    double m = -INFINITY;
#pragma omp parallel for schedule(dynamic, 16) reduction(max : m)
    for (int64_t i = 0; i < N; ++i) {
        double s = (prec ? dot_f64(z + i * d, z + i * d, d) : (double)dot_f32(z + i * d, z + i * d, d)) * inv_t;
        if (s > m) m = s;
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < N; ++i) {
        const float* zi = z + i * d;
        const int32_t li = labels[i];
        double Di = 0.0, psi = 0.0, ci = 0.0;
        for (int64_t j = 0; j < N; ++j) {
            if (j == i) continue;
            double s = (prec ? dot_f64(zi, z + j * d, d) : (double)dot_f32(zi, z + j * d, d)) * inv_t;
            if (s > m) { /* keep the shift an upper bound; cannot happen for unit rows */ }
            Di += exp(s - m);
            if (labels[j] == li) { psi += s; ci += 1.0; }
        }
        D[i] = Di; ps[i] = psi; row_cnt[i] = ci; row_lse[i] = m + log(Di + 1e-16);
    }
    double acc = 0.0;
    for (int64_t i = 0; i < N; ++i) acc += ps[i] / row_cnt[i] - row_lse[i];
    *loss = -acc / (double)N;
    if (dz) {
#pragma omp parallel
        {
            double* accd = (double*)malloc(sizeof(double) * (size_t)d);
            float* accf = (float*)malloc(sizeof(float) * (size_t)d);
#pragma omp for schedule(dynamic, 16)
            for (int64_t i = 0; i < N; ++i) {
                const float* zi = z + i * d;
                const int32_t li = labels[i];
                const double inv_ci = 1.0 / row_cnt[i], inv_Di = 1.0 / (D[i] + 1e-16);
                if (prec) memset(accd, 0, sizeof(double) * (size_t)d); else memset(accf, 0, sizeof(float) * (size_t)d);
                for (int64_t j = 0; j < N; ++j) {
                    if (j == i) continue;
                    double s = (prec ? dot_f64(zi, z + j * d, d) : (double)dot_f32(zi, z + j * d, d)) * inv_t;
                    double w = exp(s - m) * (inv_Di + 1.0 / (D[j] + 1e-16));
                    if (labels[j] == li) w -= inv_ci + 1.0 / row_cnt[j];
                    w *= inv_t / (double)N;
                    if (prec) axpy_f64(w, z + j * d, accd, d); else axpy_f32((float)w, z + j * d, accf, d);
                }
                if (prec) for (int64_t k = 0; k < d; ++k) dz[i * d + k] = (float)accd[k];
                else memcpy(dz + i * d, accf, sizeof(float) * (size_t)d);
            }
            free(accd); free(accf);
        }
    }
    free(D); free(ps);
    return 0;
}

/* ---------- IIC ---------- */
MULTIVERSION static double rowdot(const float* a, const float* b, int n) {
    double s = 0.0;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < n; ++i) s += (double)a[i] * (double)b[i];
    return s;
}
MULTIVERSION static float rowdot_f32(const float* a, const float* b, int n) {
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
MULTIVERSION static void rowaxpy(float w, const float* a, float* acc, int n) {
    for (int i = 0; i < n; ++i) acc[i] += w * a[i];
}

/* raw joint J[k1,k2,dy,dx] = sum_{b,h,w} x[b,k1,h+dy-p,w+dx-p] y[b,k2,h,w]  (discreteMI.py:227-232) */
int oracle_iic_raw_joint(const float* x, const float* y, int B, int K, int H, int W, int pad, int prec, double* J) {
    const int T = 2 * pad + 1;
    const size_t nj = (size_t)K * K * T * T;
    memset(J, 0, sizeof(double) * nj);
#pragma omp parallel
    {
        double* Jl = (double*)calloc(nj, sizeof(double));
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (int b = 0; b < B; ++b)
            for (int k1 = 0; k1 < K; ++k1) {
                const float* xb = x + ((size_t)b * K + k1) * H * W;
                for (int k2 = 0; k2 < K; ++k2) {
                    const float* yb = y + ((size_t)b * K + k2) * H * W;
                    for (int dy = 0; dy < T; ++dy)
                        for (int dx = 0; dx < T; ++dx) {
                            /* y pixel (h,w) pairs with x pixel (h+dy-p, w+dx-p) */
                            const int oy = dy - pad, ox = dx - pad;
                            const int h0 = oy < 0 ? -oy : 0, h1 = oy > 0 ? H - oy : H;
                            const int w0 = ox < 0 ? -ox : 0, w1 = ox > 0 ? W - ox : W;
                            double s = 0.0;
                            if (w1 > w0)
                                for (int h = h0; h < h1; ++h) {
                                    const float* xr = xb + (size_t)(h + oy) * W + (w0 + ox);
                                    const float* yr = yb + (size_t)h * W + w0;
                                    s += prec ? rowdot(xr, yr, w1 - w0) : (double)rowdot_f32(xr, yr, w1 - w0);
                                }
                            Jl[(((size_t)k1 * K + k2) * T + dy) * T + dx] += s;
                        }
                }
            }
#pragma omp critical
        for (size_t i = 0; i < nj; ++i) J[i] += Jl[i];
        free(Jl);
    }
    return 0;
}

/* epilogue on the raw joint (discreteMI.py:233-243 or :246-261, then :154-165) and dLoss/dJ.
 * J, gJ: [K,K,T,T]; P00: [K,K] = p_i_j[0][0] (:152).  n_pixels only used when pad == 0. */
int oracle_iic_epilogue(const double* J, int K, int pad, int symmetric, double lamda, double eps, double n_pixels,
                        double* loss, double* P00, double* gJ) {
    const int T = 2 * pad + 1, TT = T * T, KK = K * K;
    const size_t n = (size_t)KK * TT;
    double* P = (double*)malloc(sizeof(double) * n);    /* [TT][K][K] */
    double* Bm = (double*)malloc(sizeof(double) * n);
    double* g = (double*)malloc(sizeof(double) * n);
    double* sd = (double*)malloc(sizeof(double) * (size_t)TT);
    double total = 1.0;
#define JIDX(k1, k2, dd) ((((size_t)(k1)) * K + (k2)) * TT + (dd))
#define PIDX(dd, k1, k2) ((((size_t)(dd)) * K + (k1)) * K + (k2))
    if (pad > 0) {
        double mn = INFINITY;
        for (size_t i = 0; i < n; ++i) if (J[i] < mn) mn = J[i];
        for (int dd = 0; dd < TT; ++dd) {
            double s = 0.0;
            for (int k1 = 0; k1 < K; ++k1) for (int k2 = 0; k2 < K; ++k2) s += J[JIDX(k1, k2, dd)] - mn + 1e-8;
            sd[dd] = s;
            for (int k1 = 0; k1 < K; ++k1) for (int k2 = 0; k2 < K; ++k2)
                Bm[PIDX(dd, k1, k2)] = (J[JIDX(k1, k2, dd)] - mn + 1e-8) / s;
        }
        total = 0.0;
        for (int dd = 0; dd < TT; ++dd)
            for (int k1 = 0; k1 < K; ++k1) for (int k2 = 0; k2 < K; ++k2) {
                double c = symmetric ? 0.5 * (Bm[PIDX(dd, k1, k2)] + Bm[PIDX(dd, k2, k1)]) : Bm[PIDX(dd, k1, k2)];
                P[PIDX(dd, k1, k2)] = c; total += c;
            }
        for (size_t i = 0; i < n; ++i) P[i] /= total;
    } else {
        for (int k1 = 0; k1 < K; ++k1) for (int k2 = 0; k2 < K; ++k2) Bm[PIDX(0, k1, k2)] = J[JIDX(k1, k2, 0)] / n_pixels;
        for (int k1 = 0; k1 < K; ++k1) for (int k2 = 0; k2 < K; ++k2)
            P[PIDX(0, k1, k2)] = symmetric ? 0.5 * (Bm[PIDX(0, k1, k2)] + Bm[PIDX(0, k2, k1)]) : Bm[PIDX(0, k1, k2)];
    }
    double L = 0.0, gdotP = 0.0;
    for (int dd = 0; dd < TT; ++dd) {
        for (int k1 = 0; k1 < K; ++k1) {
            double b = 0.0;                                  /* p_j_mat: sum over k2 (dim 3) */
            for (int k2 = 0; k2 < K; ++k2) b += P[PIDX(dd, k1, k2)];
            for (int k2 = 0; k2 < K; ++k2) {
                double a = 0.0;                              /* p_i_mat: sum over k1 (dim 2) */
                for (int kk = 0; kk < K; ++kk) a += P[PIDX(dd, kk, k2)];
                const double p = P[PIDX(dd, k1, k2)];
                L += -p * (log(p + eps) - lamda * log(a + eps) - lamda * log(b + eps));
                const double gq = -(log(p + eps) + p / (p + eps) - lamda * (log(a + eps) + a / (a + eps))
                                    - lamda * (log(b + eps) + b / (b + eps))) / (double)TT;
                g[PIDX(dd, k1, k2)] = gq; gdotP += gq * p;
            }
        }
    }
    *loss = L / (double)TT;
    for (int k1 = 0; k1 < K; ++k1) for (int k2 = 0; k2 < K; ++k2) P00[k1 * K + k2] = P[PIDX(0, k1, k2)];
    if (gJ) {
        if (pad > 0) {
            for (size_t i = 0; i < n; ++i) g[i] = (g[i] - gdotP) / total;               /* through P = C/total */
            for (int dd = 0; dd < TT; ++dd) {
                double dot = 0.0;
                for (int k1 = 0; k1 < K; ++k1) for (int k2 = 0; k2 < K; ++k2) {
                    double gb = symmetric ? 0.5 * (g[PIDX(dd, k1, k2)] + g[PIDX(dd, k2, k1)]) : g[PIDX(dd, k1, k2)];
                    dot += gb * Bm[PIDX(dd, k1, k2)];
                }
                for (int k1 = 0; k1 < K; ++k1) for (int k2 = 0; k2 < K; ++k2) {
                    double gb = symmetric ? 0.5 * (g[PIDX(dd, k1, k2)] + g[PIDX(dd, k2, k1)]) : g[PIDX(dd, k1, k2)];
                    gJ[JIDX(k1, k2, dd)] = (gb - dot) / sd[dd];
                }
            }
        } else {
            for (int k1 = 0; k1 < K; ++k1) for (int k2 = 0; k2 < K; ++k2) {
                double gb = symmetric ? 0.5 * (g[PIDX(0, k1, k2)] + g[PIDX(0, k2, k1)]) : g[PIDX(0, k1, k2)];
                gJ[JIDX(k1, k2, 0)] = gb / n_pixels;
            }
        }
    }
    free(P); free(Bm); free(g); free(sd);
    return 0;
}

/* adjoint of the raw joint: dx, dy [B,K,H,W] float32 from gJ [K,K,T,T] */
int oracle_iic_input_grads(const float* x, const float* y, const double* gJ, int B, int K, int H, int W, int pad,
                           float* dxo, float* dyo) {
    const int T = 2 * pad + 1;
    memset(dxo, 0, sizeof(float) * (size_t)B * K * H * W);
    memset(dyo, 0, sizeof(float) * (size_t)B * K * H * W);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
        for (int k1 = 0; k1 < K; ++k1)
            for (int k2 = 0; k2 < K; ++k2) {
                const float* xb = x + ((size_t)b * K + k1) * H * W;
                const float* yb = y + ((size_t)b * K + k2) * H * W;
                float* gx = dxo + ((size_t)b * K + k1) * H * W;
                float* gy = dyo + ((size_t)b * K + k2) * H * W;
                for (int dy = 0; dy < T; ++dy)
                    for (int dx = 0; dx < T; ++dx) {
                        const float g = (float)gJ[(((size_t)k1 * K + k2) * T + dy) * T + dx];
                        const int oy = dy - pad, ox = dx - pad;
                        const int h0 = oy < 0 ? -oy : 0, h1 = oy > 0 ? H - oy : H;
                        const int w0 = ox < 0 ? -ox : 0, w1 = ox > 0 ? W - ox : W;
                        if (w1 <= w0) continue;
                        for (int h = h0; h < h1; ++h) {
                            rowaxpy(g, yb + (size_t)h * W + w0, gx + (size_t)(h + oy) * W + (w0 + ox), w1 - w0);
                            rowaxpy(g, xb + (size_t)(h + oy) * W + (w0 + ox), gy + (size_t)h * W + w0, w1 - w0);
                        }
                    }
            }
    return 0;
}

/* whole IIDSegmentationLoss fwd+bwd: what bench.py times as the CPU baseline */
int oracle_iic_fwd_bwd(const float* x, const float* y, int B, int K, int H, int W, int pad, int symmetric,
                       double lamda, double eps, int prec, double* loss, double* P00, float* dx, float* dy) {
    const int T = 2 * pad + 1;
    const size_t nj = (size_t)K * K * T * T;
    double* J = (double*)malloc(sizeof(double) * nj);
    double* gJ = (double*)malloc(sizeof(double) * nj);
    oracle_iic_raw_joint(x, y, B, K, H, W, pad, prec, J);
    oracle_iic_epilogue(J, K, pad, symmetric, lamda, eps, (double)B * H * W, loss, P00, gJ);
    if (dx && dy) oracle_iic_input_grads(x, y, gJ, B, K, H, W, pad, dx, dy);
    free(J); free(gJ);
    return 0;
}
