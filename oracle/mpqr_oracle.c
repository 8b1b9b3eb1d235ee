/*
 * oracle/mpqr_oracle.c — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's
 * blocked Householder QR hot path).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product library
 * (mixedprecisionblockqr_b200/csrc -> libmpqr.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 *   (1) the reference's own known answers (3x3 of Cuda/qr.cu:1397-1401 and
 *       python/test_data.py:18-22; the [0,0,2] reflector of python/test_all.py:12-20;
 *       python fixtures test_data.py:4-57),
 *   (2) golden vectors produced by the UNMODIFIED reference compiled in
 *       oracle/_ref/libref_qr.so (tests/golden/ npz files, generator tests/golden/make_golden.py),
 *   (3) oracle/_ref itself, bit-for-bit, whenever it is present.
 *
 * All matrices are row-major FP32.  The packed factor needs (m+1) rows: the unit
 * Householder vector of column k is stored at rows k+1..m of column k, i.e. one row
 * BELOW the diagonal (Cuda/qr.cu:283-285, :1062).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * Deterministic input generator shared (bit-for-bit) with the CUDA generator in
 * mixedprecisionblockqr_b200/csrc and the numpy one in tests: uniform [0,1) with 24
 * random bits, the distribution of the reference's h_generate_random_matrix
 * (Cuda/mmult.cuh:39-60) but stateless so every shard can be produced anywhere.
 * ---------------------------------------------------------------------------------- */
static inline uint64_t orc_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

float orc_uniform01(uint64_t seed, uint64_t idx) {
    uint64_t h = orc_mix64(orc_mix64(seed) + idx);
    return (float)(h >> 40) * (1.0f / 16777216.0f);
}

/* Fills rows [0,m) x cols [0,n) of a row-major buffer with leading dimension ld.
 * Element (i,j) depends only on (seed, i*n+j). */
void orc_fill_uniform(float* A, long m, long n, long ld, uint64_t seed) {
    for (long i = 0; i < m; ++i)
        for (long j = 0; j < n; ++j) A[i * ld + j] = orc_uniform01(seed, (uint64_t)(i * n + j));
}

/* ------------------------------------------------------------------------------------
 * Panel factorisation — restates h_householder_qr, Cuda/qr.cu:198-293.
 *   for k in [off, min(off+pw, n)):
 *     u = A[k:m, k]                                   (:221-226)
 *     s = +1 if u0 >= 0 else -1                       (:229-235)
 *     if sum(u^2) == 0: skip the column               (:238-244)
 *     u0 += s*||u||; w = u/||u||                      (:245-257)
 *     A[k:m, k:tau) -= 2 w (w^T A[k:m, k:tau))        (:263-280)  (panel columns only)
 *     A[k+1:m+1, k] = w                               (:283-285)  (shifted one row down)
 * Summation order is the reference's (left-to-right float adds), so the result is
 * bit-identical to oracle/_ref on the same compiler.
 * Extension (reference is undefined for m < n, Cuda/qr.cu:222-230): columns k >= m are
 * left untouched.
 * ---------------------------------------------------------------------------------- */
void orc_householder_panel(float* A, int m, int n, int off, int pw) {
    int tau = off + pw > n ? n : off + pw;
    for (int k = off; k < tau && k < m; ++k) {
        int len = m - k;
        float* w = (float*)malloc((size_t)len * sizeof(float));
        for (int i = 0; i < len; ++i) w[i] = A[(size_t)(i + k) * n + k];
        int s = w[0] >= 0 ? 1 : -1;
        float acc = 0;
        for (int i = 0; i < len; ++i) acc += w[i] * w[i];
        if (acc == 0) {
            free(w);
            continue;
        }
        acc = sqrtf(acc);
        w[0] = s * acc + w[0];
        acc = 0;
        for (int i = 0; i < len; ++i) acc += w[i] * w[i];
        acc = sqrtf(acc);
        for (int i = 0; i < len; ++i) w[i] /= acc;

        int nc = tau - k;
        float* wa = (float*)malloc((size_t)nc * sizeof(float));
        for (int c = 0; c < nc; ++c) {
            float d = 0;
            for (int i = 0; i < len; ++i) d += w[i] * A[(size_t)(i + k) * n + k + c];
            wa[c] = d;
        }
        for (int i = 0; i < len; ++i)
            for (int c = 0; c < nc; ++c) {
                float t = w[i] * wa[c];
                A[(size_t)(i + k) * n + k + c] = A[(size_t)(i + k) * n + k + c] - 2 * t;
            }
        for (int i = 0; i < len; ++i) A[(size_t)(k + 1 + i) * n + k] = w[i];
        free(wa);
        free(w);
    }
}

/* ------------------------------------------------------------------------------------
 * WY accumulation — restates h_wy_transform, Cuda/qr.cu:337-426, but RETURNS W and Y
 * (D x pw row-major, D = m-off) instead of freeing them (:420-421), and the dense
 * D x D matrix I - W Y^T only when `dense` is non-NULL (it is what the reference
 * returns, :424).  Y[:,0]=w_1, W[:,0]=2 w_1 (:349-352); z = 2 (I - W Y^T)[:, i:] w_i
 * (:361-387) with the SAME evaluation order as the reference (dense I - W Y^T formed in
 * FP32 first, then the mat-vec) so W is bit-identical; Y[idx<i, i] = 0 (:392-394).
 * ---------------------------------------------------------------------------------- */
void orc_wy_transform(const float* A, int m, int n, int off, int pw, float* W, float* Y, float* dense) {
    int D = m - off;
    float* P = dense ? dense : (float*)malloc((size_t)D * D * sizeof(float));
    float* z = (float*)malloc((size_t)D * sizeof(float));
    for (int i = 0; i < D; ++i) {
        float w = A[(size_t)(i + off + 1) * n + off];
        Y[(size_t)i * pw] = w;
        W[(size_t)i * pw] = 2 * w;
    }
    for (int c = 1; c <= pw; ++c) {
        /* P = I - W[:, :c] Y[:, :c]^T  (for c == pw this is the returned matrix, :402-418) */
        if (c < pw || dense) {
            for (int a = 0; a < D; ++a)
                for (int b = 0; b < D; ++b) {
                    float d = 0;
                    for (int t = 0; t < c; ++t) d += W[(size_t)a * pw + t] * Y[(size_t)b * pw + t];
                    P[(size_t)a * D + b] = (a == b) ? 1 - d : -d;
                }
        }
        if (c == pw) break;
        for (int a = 0; a < D; ++a) {
            float d = 0;
            for (int b = c; b < D; ++b) d += P[(size_t)a * D + b] * A[(size_t)(off + b + 1) * n + off + c];
            z[a] = 2 * d;
        }
        for (int a = 0; a < D; ++a) {
            Y[(size_t)a * pw + c] = (a < c) ? 0 : A[(size_t)(off + a + 1) * n + off + c];
            W[(size_t)a * pw + c] = z[a];
        }
    }
    free(z);
    if (!dense) free(P);
}

/* Scalable W/Y: same recurrence evaluated as z = 2 (w_i - W (Y^T w_i)) with double
 * accumulation, O(D pw^2) instead of O(D^2 pw^2).  Agrees with orc_wy_transform to FP32
 * rounding (tests/test_oracle.py).  Used by orc_block_qr for sizes the dense form cannot
 * reach. */
void orc_wy_factors(const float* A, int m, int n, int off, int pw, float* W, float* Y) {
    int D = m - off;
    double* g = (double*)malloc((size_t)pw * sizeof(double));
    for (int c = 0; c < pw; ++c) {
        for (int a = 0; a < D; ++a)
            Y[(size_t)a * pw + c] = (a < c || off + c >= n) ? 0.f : A[(size_t)(off + a + 1) * n + off + c];
        for (int t = 0; t < c; ++t) {
            double d = 0;
            for (int a = c; a < D; ++a) d += (double)Y[(size_t)a * pw + t] * Y[(size_t)a * pw + c];
            g[t] = d;
        }
        for (int a = 0; a < D; ++a) {
            double d = Y[(size_t)a * pw + c];
            for (int t = 0; t < c; ++t) d -= (double)W[(size_t)a * pw + t] * g[t];
            W[(size_t)a * pw + c] = (float)(2 * d);
        }
    }
    free(g);
}

/* ------------------------------------------------------------------------------------
 * Block QR, literal form — restates h_block_qr, Cuda/qr.cu:1275-1326: per panel
 *   panel factor -> dense panelQ = I - W Y^T -> A[l:, tau:] = panelQ^T A_old[l:, tau:]
 *   (:1295-1305, inner index runs over panelQ ROWS) -> Q[:, l:] = Q_old[:, l:] panelQ
 *   (:1309-1319).
 * O(m^2 n^2 / r); bit-identical to oracle/_ref; use only for small shapes.
 * A is (m+1) x n packed, Q is m x m and must hold the identity on entry (:1281 is
 * commented out in the reference, the caller initialises Q: :1868-1872).
 * ---------------------------------------------------------------------------------- */
void orc_block_qr_dense(float* A, float* Q, int m, int n, int r) {
    for (int lam = 0; lam < n;) {
        int tau = lam + r < n ? lam + r : n;
        int D = m - lam, pw = tau - lam;
        orc_householder_panel(A, m, n, lam, pw);
        float* W = (float*)malloc((size_t)D * pw * sizeof(float));
        float* Y = (float*)malloc((size_t)D * pw * sizeof(float));
        float* P = (float*)malloc((size_t)D * D * sizeof(float));
        orc_wy_transform(A, m, n, lam, pw, W, Y, P);
        float* old = (float*)malloc((size_t)m * n * sizeof(float));
        memcpy(old, A, (size_t)m * n * sizeof(float));
        for (int i = lam; i < m; ++i)
            for (int j = tau; j < n; ++j) {
                float d = 0;
                for (int t = 0; t < D; ++t) d += P[(size_t)t * D + (i - lam)] * old[(size_t)(t + lam) * n + j];
                A[(size_t)i * n + j] = d;
            }
        free(old);
        old = (float*)malloc((size_t)m * m * sizeof(float));
        memcpy(old, Q, (size_t)m * m * sizeof(float));
        for (int i = 0; i < m; ++i)
            for (int j = lam; j < m; ++j) {
                float d = 0;
                for (int t = 0; t < D; ++t) d += old[(size_t)i * m + t + lam] * P[(size_t)t * D + (j - lam)];
                Q[(size_t)i * m + j] = d;
            }
        free(old);
        free(P);
        free(W);
        free(Y);
        lam = tau;
    }
}

/* ------------------------------------------------------------------------------------
 * Block QR, scalable form: same panel loop (tau = min(lam+r, n), Cuda/qr.cu:1284-1285)
 * and same panel kernel, but the trailing update is applied in factored form
 *   A[l:, tau:] -= Y (W^T A[l:, tau:])       (== panelQ^T A)
 *   Q[:, l:]    -= (Q[:, l:] W) Y^T          (== Q panelQ)
 * with FP32 storage and double accumulation.  O(m n r) per panel.  Q may be NULL.
 * Extension for m < n: only min(m,n) columns are factored, every trailing column is
 * still updated.
 * ---------------------------------------------------------------------------------- */
void orc_block_qr(float* A, float* Q, int m, int n, int r) {
    int kmax = m < n ? m : n;
    for (int lam = 0; lam < kmax;) {
        int tau = lam + r < kmax ? lam + r : kmax;
        int D = m - lam, pw = tau - lam;
        orc_householder_panel(A, m, n, lam, pw);
        float* W = (float*)malloc((size_t)D * pw * sizeof(float));
        float* Y = (float*)malloc((size_t)D * pw * sizeof(float));
        orc_wy_factors(A, m, n, lam, pw, W, Y);
        int nt = n - tau;
        if (nt > 0) {
            double* S = (double*)calloc((size_t)pw * nt, sizeof(double));
            for (int a = 0; a < D; ++a) {
                const float* arow = A + (size_t)(a + lam) * n + tau;
                for (int t = 0; t < pw; ++t) {
                    double wv = W[(size_t)a * pw + t];
                    if (wv == 0) continue;
                    double* s = S + (size_t)t * nt;
                    for (int j = 0; j < nt; ++j) s[j] += wv * arow[j];
                }
            }
            double* acc = (double*)malloc((size_t)nt * sizeof(double));
            for (int a = 0; a < D; ++a) {
                float* arow = A + (size_t)(a + lam) * n + tau;
                for (int j = 0; j < nt; ++j) acc[j] = 0;
                for (int t = 0; t < pw; ++t) {
                    double yv = Y[(size_t)a * pw + t];
                    if (yv == 0) continue;
                    const double* s = S + (size_t)t * nt;
                    for (int j = 0; j < nt; ++j) acc[j] += yv * s[j];
                }
                for (int j = 0; j < nt; ++j) arow[j] = (float)((double)arow[j] - acc[j]);
            }
            free(acc);
            free(S);
        }
        if (Q) {
            double* qw = (double*)malloc((size_t)pw * sizeof(double));
            for (int i = 0; i < m; ++i) {
                float* qrow = Q + (size_t)i * m + lam;
                for (int t = 0; t < pw; ++t) qw[t] = 0;
                for (int a = 0; a < D; ++a) {
                    double qv = qrow[a];
                    if (qv == 0) continue;
                    for (int t = 0; t < pw; ++t) qw[t] += qv * W[(size_t)a * pw + t];
                }
                for (int a = 0; a < D; ++a) {
                    double d = 0;
                    for (int t = 0; t < pw; ++t) d += qw[t] * Y[(size_t)a * pw + t];
                    qrow[a] = (float)((double)qrow[a] - d);
                }
            }
            free(qw);
        }
        free(W);
        free(Y);
        lam = tau;
    }
}

/* Explicit Q by backward accumulation — restates h_q_backward_accumulation,
 * Cuda/qr.cu:296-335 (GVL Alg. 5.1.5): for j = n-1..0: Q[j:, j:] -= 2 v (v^T Q[j:, j:])
 * with v read from the shifted storage A[(row+1)*n + j].  Same float order => bit-exact. */
void orc_q_backward_accumulation(const float* A, float* Q, int m, int n) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Q[(size_t)i * m + j] = (i == j) ? 1.f : 0.f;
    int kmax = m < n ? m : n;
    float* tmp = (float*)malloc((size_t)m * sizeof(float));
    for (int j = kmax - 1; j >= 0; --j) {
        for (int c = j; c < m; ++c) {
            float d = 0.0;
            for (int i = j; i < m; ++i) d += A[(size_t)(i + 1) * n + j] * Q[(size_t)i * m + c];
            tmp[c - j] = d;
        }
        for (int i = j; i < m; ++i)
            for (int c = j; c < m; ++c)
                Q[(size_t)i * m + c] = Q[(size_t)i * m + c] - 2.0f * A[(size_t)(i + 1) * n + j] * tmp[c - j];
    }
    free(tmp);
}

/* R extraction — restates h_strip_R_from_A, Cuda/qr.cu:85-100. */
void orc_strip_R(const float* A, float* R, int m, int n) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) R[(size_t)i * n + j] = (i <= j) ? A[(size_t)i * n + j] : 0.f;
}

/* ------------------------------------------------------------------------------------
 * Metrics.  Definitions follow Cuda/qr.cu:115-196 but are evaluated in double so that
 * they measure the factorisation and not the metric's own FP32 rounding:
 *   backward error  ||A - Q R||_F / ||A||_F                          (:115-135)
 *   q error         max signed entry of Q^T Q - I (NOT a norm)        (:137-171)
 *   orthogonality   ||Q^T Q - I||_F                                   (BASELINE.json)
 * ---------------------------------------------------------------------------------- */
double orc_backward_error(const float* A0, const float* R, const float* Q, int m, int n) {
    double num = 0, den = 0;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double d = 0;
            int tmax = j < m - 1 ? j : m - 1; /* R is upper trapezoidal */
            for (int t = 0; t <= tmax; ++t) d += (double)Q[(size_t)i * m + t] * R[(size_t)t * n + j];
            double a = A0[(size_t)i * n + j];
            num += (a - d) * (a - d);
            den += a * a;
        }
    return sqrt(num) / sqrt(den);
}

/* Backward error straight from the packed factor (no explicit Q): applies the stored
 * reflectors to R in reverse order in double, O(m n^2).  A0 is the m x n input. */
double orc_backward_error_packed(const float* A0, const float* Apacked, int m, int n) {
    int kmax = m < n ? m : n;
    double* B = (double*)calloc((size_t)m * n, sizeof(double));
    for (int i = 0; i < m; ++i)
        for (int j = i; j < n; ++j) B[(size_t)i * n + j] = Apacked[(size_t)i * n + j];
    double* d = (double*)malloc((size_t)n * sizeof(double));
    for (int k = kmax - 1; k >= 0; --k) {
        for (int j = k; j < n; ++j) d[j] = 0;
        for (int i = k; i < m; ++i) {
            double w = Apacked[(size_t)(i + 1) * n + k];
            if (w == 0) continue;
            for (int j = k; j < n; ++j) d[j] += w * B[(size_t)i * n + j];
        }
        for (int i = k; i < m; ++i) {
            double w = 2.0 * Apacked[(size_t)(i + 1) * n + k];
            if (w == 0) continue;
            for (int j = k; j < n; ++j) B[(size_t)i * n + j] -= w * d[j];
        }
    }
    double num = 0, den = 0;
    for (size_t t = 0; t < (size_t)m * n; ++t) {
        double a = A0[t];
        num += (a - B[t]) * (a - B[t]);
        den += a * a;
    }
    free(d);
    free(B);
    return sqrt(num) / sqrt(den);
}

double orc_q_error_max(const float* Q, int m) {
    double mx = 0;
    for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) {
            double d = 0;
            for (int t = 0; t < m; ++t) d += (double)Q[(size_t)t * m + a] * Q[(size_t)t * m + b];
            d -= (a == b);
            if (d > mx) mx = d;
        }
    return mx;
}

double orc_orthogonality_fro(const float* Q, int m) {
    double s = 0;
    for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) {
            double d = 0;
            for (int t = 0; t < m; ++t) d += (double)Q[(size_t)t * m + a] * Q[(size_t)t * m + b];
            d -= (a == b);
            s += d * d;
        }
    return sqrt(s);
}

/* Reference flop model, Cuda/qr.cu:102-113 (4 m^2 n - m n^2 + n^3/3) and the
 * BASELINE.json Householder count (2 m n^2 - 2 n^3/3; 2 m^2 n - 2 m^3/3 when m < n). */
double orc_ref_flop_model(int m, int n) {
    double mf = m, nf = n;
    return 4.0 * mf * mf * nf - mf * nf * nf + nf * nf * nf / 3.0;
}
double orc_householder_flops(double m, double n) {
    return m >= n ? 2.0 * m * n * n - 2.0 * n * n * n / 3.0 : 2.0 * m * m * n - 2.0 * m * m * m / 3.0;
}

/* ------------------------------------------------------------------------------------
 * TSQR — restates ts_qr, python/ca_qr.py:25-43, in double, generalised from the
 * reference's fixed 4 row blocks to `nblk` (power of two) blocks with a binary tree:
 *   h = m / nblk (remainder rows dropped, as :27 does); local complete Householder QR
 *   of every block (python/qr.py:25-70 conventions: sign = -1 if u0 >= 0 else +1, skip
 *   all-zero columns, skip the last column of a square block) -> stack pairs of R,
 *   factor again, ... -> R (n x n) and thin Q = blkdiag(Q_i) blkdiag(Q_ij) ... (:39-41).
 * Needs h >= n.  Q is (h*nblk) x n, R is n x n, both row-major double.
 * ---------------------------------------------------------------------------------- */
static void orc_hh_complete(const double* Ain, int m, int n, double* Q /* m x m */, double* R /* m x n */) {
    memcpy(R, Ain, (size_t)m * n * sizeof(double));
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Q[(size_t)i * m + j] = (i == j);
    double* w = (double*)malloc((size_t)m * sizeof(double));
    double* d = (double*)malloc((size_t)(m > n ? m : n) * sizeof(double));
    for (int k = 0; k < n && k < m; ++k) {
        if (m == n && k == n - 1) break; /* python/qr.py:49-50 */
        int len = m - k;
        double nrm = 0, mxabs = 0;
        for (int i = 0; i < len; ++i) {
            w[i] = R[(size_t)(i + k) * n + k];
            nrm += w[i] * w[i];
            if (fabs(w[i]) > mxabs) mxabs = fabs(w[i]);
        }
        if (mxabs <= 1e-8) continue; /* np.allclose(col, 0): atol 1e-8, python/qr.py:54 */
        nrm = sqrt(nrm);
        double sgn = w[0] >= 0 ? -1.0 : 1.0; /* python/qr.py:19 */
        w[0] -= sgn * nrm;
        double wn = 0;
        for (int i = 0; i < len; ++i) wn += w[i] * w[i];
        wn = sqrt(wn);
        for (int i = 0; i < len; ++i) w[i] /= wn;
        /* R = H R */
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int i = 0; i < len; ++i) s += w[i] * R[(size_t)(i + k) * n + j];
            d[j] = s;
        }
        for (int i = 0; i < len; ++i)
            for (int j = 0; j < n; ++j) R[(size_t)(i + k) * n + j] -= 2 * w[i] * d[j];
        /* Q = Q H */
        for (int i = 0; i < m; ++i) {
            double s = 0;
            for (int t = 0; t < len; ++t) s += Q[(size_t)i * m + k + t] * w[t];
            for (int t = 0; t < len; ++t) Q[(size_t)i * m + k + t] -= 2 * s * w[t];
        }
    }
    free(w);
    free(d);
}

int orc_tsqr(const double* A, long m, int n, int nblk, double* Qthin, double* Rout) {
    long h = m / nblk;
    if (h < n || nblk < 1 || (nblk & (nblk - 1))) return -1;
    /* level 0 */
    double* Rs = (double*)malloc((size_t)nblk * n * n * sizeof(double));
    double* Qc = (double*)malloc((size_t)h * h * sizeof(double));
    double* Rc = (double*)malloc((size_t)h * n * sizeof(double));
    /* Qacc: thin Q of every leaf, (h x n) each, later right-multiplied by tree Q blocks */
    for (int b = 0; b < nblk; ++b) {
        orc_hh_complete(A + (size_t)b * h * n, (int)h, n, Qc, Rc);
        memcpy(Rs + (size_t)b * n * n, Rc, (size_t)n * n * sizeof(double));
        for (long i = 0; i < h; ++i)
            for (int j = 0; j < n; ++j) Qthin[((size_t)b * h + i) * n + j] = Qc[(size_t)i * h + j];
    }
    free(Qc);
    free(Rc);
    /* tree levels: pair (2p, 2p+1) */
    double* S = (double*)malloc((size_t)2 * n * n * sizeof(double));
    double* Q2 = (double*)malloc((size_t)4 * n * n * sizeof(double));
    double* R2 = (double*)malloc((size_t)2 * n * n * sizeof(double));
    double* tmp = (double*)malloc((size_t)n * sizeof(double));
    int cnt = nblk;
    long span = h; /* rows of Qthin covered by one node at this level */
    while (cnt > 1) {
        for (int p = 0; p < cnt / 2; ++p) {
            memcpy(S, Rs + (size_t)(2 * p) * n * n, (size_t)n * n * sizeof(double));
            memcpy(S + (size_t)n * n, Rs + (size_t)(2 * p + 1) * n * n, (size_t)n * n * sizeof(double));
            orc_hh_complete(S, 2 * n, n, Q2, R2);
            memcpy(Rs + (size_t)p * n * n, R2, (size_t)n * n * sizeof(double));
            /* rows of child 2p get right-multiplied by Q2[0:n, 0:n], child 2p+1 by Q2[n:2n, 0:n] */
            for (int c = 0; c < 2; ++c) {
                long r0 = ((long)(2 * p + c)) * span;
                for (long i = r0; i < r0 + span; ++i) {
                    double* q = Qthin + (size_t)i * n;
                    for (int j = 0; j < n; ++j) {
                        double s = 0;
                        for (int t = 0; t < n; ++t) s += q[t] * Q2[(size_t)(c * n + t) * (2 * n) + j];
                        tmp[j] = s;
                    }
                    memcpy(q, tmp, (size_t)n * sizeof(double));
                }
            }
        }
        cnt /= 2;
        span *= 2;
    }
    memcpy(Rout, Rs, (size_t)n * n * sizeof(double));
    free(S);
    free(Q2);
    free(R2);
    free(tmp);
    free(Rs);
    return 0;
}
