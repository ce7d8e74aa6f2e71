// dbslmm_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A dependency-free restatement of the reference's per-LD-block fit, written from
// reading /root/reference/scr/{dtpr,dbslmmfit,helpers}.cpp.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product path (libdbslmm_b200.so) never links or calls it.
//
// PARITY PIN: the restatement is checked against the UNMODIFIED reference sources
// compiled over a minimal Armadillo/Boost shim (oracle/shim, oracle/build_ref.sh ->
// oracle/_ref/) in tests/test_oracle.py and against the committed golden
// vectors in tests/golden/ that the same reference build produced.  The shim replaces
// Armadillo's BLAS/LAPACK back end with plain loops, so pins are at the 1e-12 level,
// not bit level ("Armadillo version unpinned").
//
// Each function cites the reference file:line it follows.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

// ---- Armadillo statistics as used by nomalizeVec (dtpr.cpp:375-380) -----------------
// arma::mean: straight sum / n (two-accumulator unrolled in arrayops::accumulate).
double arma_mean(const double* x, int n) {
    double a1 = 0.0, a2 = 0.0;
    int i = 0;
    for (; i + 1 < n; i += 2) { a1 += x[i]; a2 += x[i + 1]; }
    if (i < n) a1 += x[i];
    return (a1 + a2) / (double)n;
}
// arma::stddev (norm_type 0 => N-1): op_var::direct_var's mean-corrected two-pass form.
double arma_stddev(const double* x, int n) {
    if (n < 2) return 0.0;
    const double mu = arma_mean(x, n);
    double acc2 = 0.0, acc3 = 0.0;
    for (int i = 0; i < n; ++i) { const double t = mu - x[i]; acc2 += t * t; acc3 += t; }
    const double var = (acc2 - acc3 * acc3 / (double)n) / (double)(n - 1);
    return std::sqrt(var);
}

inline int64_t bed_pitch(int n_total) { return (n_total + 3) / 4; }

// ---- IO::readSNPIm (dtpr.cpp:285-364) ----------------------------------------------------
// bed points just AFTER the 3 magic bytes (the reference seeks to pos*n_bit+3, :302).
// Returns number of kept individuals (c_idv).
int read_snp_im(const uint8_t* bed, int64_t pos, int n_total, const int* indicator,
                double* geno, double* maf) {
    const int64_t n_bit = bed_pitch(n_total);                 // :293-299
    const uint8_t* row = bed + pos * n_bit;                  // :302
    double geno_mean = 0.0;
    int c = 0, c_idv = 0;
    std::vector<int> miss;
    for (int64_t i = 0; i < n_bit; ++i) {                     // :314
        const unsigned b = row[i];
        for (int j = 0; j < 4; ++j) {                         // :319
            if (i == n_bit - 1 && c == n_total) break;        // :320-322 padding samples
            if (indicator && indicator[c] == 0) { c++; continue; }  // :323-326
            c++;
            const unsigned b0 = (b >> (2 * j)) & 1u, b1 = (b >> (2 * j + 1)) & 1u;
            if (b0 == 0) {
                if (b1 == 0) { geno[c_idv] = 2.0; geno_mean += 2.0; }   // :330-334
                else         { geno[c_idv] = 1.0; geno_mean += 1.0; }   // :335-339
            } else {
                if (b1 == 1) { geno[c_idv] = 0.0; }                     // :342-346
                else         { miss.push_back(c_idv); }                 // :347-349
            }
            c_idv++;
        }
    }
    geno_mean /= (double)(c_idv - (int)miss.size());          // :358
    for (int k : miss) geno[k] = geno_mean;                   // :359-360
    double s = 0.0;
    for (int i = 0; i < c_idv; ++i) s += geno[i];
    const double af = 0.5 * s / (double)c_idv;                // :361 (geno.n_elem == c_idv for callers)
    if (maf) *maf = std::min(af, 1.0 - af);                   // :362
    return c_idv;
}

// ---- SNPPROC::nomalizeVec (dtpr.cpp:375-380) ---------------------------------------------
void normalize_vec(double* x, int n) {
    const double mu = arma_mean(x, n);
    for (int i = 0; i < n; ++i) x[i] -= mu;                   // :377
    const double sd = arma_stddev(x, n);                      // :378 (N-1)
    for (int i = 0; i < n; ++i) x[i] /= sd;
}

// C(mi x mj) = A^T B for column-major A (n x mi), B (n x mj); plain loops with 4x4
// register blocking -- a stand-in for the reference's BLAS dgemm/dsyrk call sites
// (dbslmmfit.cpp:698,700,705,752).
void atb(const double* A, int mi, const double* B, int mj, int n, double* C /*mi x mj col-major*/) {
    for (int j0 = 0; j0 < mj; j0 += 4) {
        const int jb = std::min(4, mj - j0);
        for (int i0 = 0; i0 < mi; i0 += 4) {
            const int ib = std::min(4, mi - i0);
            double acc[4][4] = {{0}};
            if (ib == 4 && jb == 4) {
                const double *a0 = A + (size_t)(i0 + 0) * n, *a1 = A + (size_t)(i0 + 1) * n,
                             *a2 = A + (size_t)(i0 + 2) * n, *a3 = A + (size_t)(i0 + 3) * n;
                const double *b0 = B + (size_t)(j0 + 0) * n, *b1 = B + (size_t)(j0 + 1) * n,
                             *b2 = B + (size_t)(j0 + 2) * n, *b3 = B + (size_t)(j0 + 3) * n;
                for (int k = 0; k < n; ++k) {
                    const double x0 = a0[k], x1 = a1[k], x2 = a2[k], x3 = a3[k];
                    const double y0 = b0[k], y1 = b1[k], y2 = b2[k], y3 = b3[k];
                    acc[0][0] += x0 * y0; acc[0][1] += x0 * y1; acc[0][2] += x0 * y2; acc[0][3] += x0 * y3;
                    acc[1][0] += x1 * y0; acc[1][1] += x1 * y1; acc[1][2] += x1 * y2; acc[1][3] += x1 * y3;
                    acc[2][0] += x2 * y0; acc[2][1] += x2 * y1; acc[2][2] += x2 * y2; acc[2][3] += x2 * y3;
                    acc[3][0] += x3 * y0; acc[3][1] += x3 * y1; acc[3][2] += x3 * y2; acc[3][3] += x3 * y3;
                }
            } else {
                for (int i = 0; i < ib; ++i)
                    for (int j = 0; j < jb; ++j) {
                        const double* a = A + (size_t)(i0 + i) * n;
                        const double* b = B + (size_t)(j0 + j) * n;
                        double s = 0.0;
                        for (int k = 0; k < n; ++k) s += a[k] * b[k];
                        acc[i][j] = s;
                    }
            }
            for (int i = 0; i < ib; ++i)
                for (int j = 0; j < jb; ++j) C[(size_t)(j0 + j) * mi + (i0 + i)] = acc[i][j];
        }
    }
}

// y = A x, A col-major m x m (symmetric in every call site).
void matvec(const double* A, const double* x, int m, double* y) {
    for (int i = 0; i < m; ++i) y[i] = 0.0;
    for (int j = 0; j < m; ++j) {
        const double xj = x[j];
        const double* col = A + (size_t)j * m;
        for (int i = 0; i < m; ++i) y[i] += col[i] * xj;
    }
}
double dotp(const double* a, const double* b, int m) {
    double s = 0.0;
    for (int i = 0; i < m; ++i) s += a[i] * b[i];
    return s;
}

// ---- DBSLMMFIT::PCGv (dbslmmfit.cpp:629-668) ---------------------------------------------
int pcgv(const double* A, const double* b, int m, int maxiter, double tol, double* x) {
    std::vector<double> Minv(m), r(b, b + m), r1(m), z(m), z1(m), p(m), Ap(m);
    for (int i = 0; i < m; ++i) {
        double d = A[(size_t)i * m + i];
        if (d == 0) d = 1e-4;                                 // :632-635
        Minv[i] = 1.0 / d;                                    // :636
    }
    for (int i = 0; i < m; ++i) { x[i] = 0.0; z[i] = Minv[i] * r[i]; p[i] = z[i]; }   // :638-644
    int iter = 0;
    double sumr2 = std::sqrt(dotp(r.data(), r.data(), m));    // :646 norm(r,2)
    while (sumr2 > tol && iter < maxiter) {                   // :648
        iter += 1;
        matvec(A, p.data(), m, Ap.data());                    // :651
        const double a = dotp(r.data(), z.data(), m) / dotp(p.data(), Ap.data(), m);   // :653
        for (int i = 0; i < m; ++i) {
            x[i] = x[i] + a * p[i];                           // :655
            r1[i] = r[i] - a * Ap[i];                         // :656
            z1[i] = Minv[i] * r1[i];                          // :657
        }
        const double bet = dotp(z1.data(), r1.data(), m) / dotp(z.data(), r.data(), m);  // :658
        for (int i = 0; i < m; ++i) { p[i] = z1[i] + bet * p[i]; z[i] = z1[i]; r[i] = r1[i]; }  // :659-661
        sumr2 = std::sqrt(dotp(r.data(), r.data(), m));       // :662
    }
    return iter;                                              // :664-667 (caller logs iter>=maxiter)
}

// In-place lower Cholesky of col-major SPD A (m x m); returns 0 or (k+1) of failing pivot.
int chol_lower(double* A, int m) {
    for (int j = 0; j < m; ++j) {
        double* cj = A + (size_t)j * m;
        for (int k = 0; k < j; ++k) {
            const double* ck = A + (size_t)k * m;
            const double ljk = ck[j];
            for (int i = j; i < m; ++i) cj[i] -= ck[i] * ljk;
        }
        const double d = cj[j];
        if (!(d > 0.0)) return j + 1;
        const double s = std::sqrt(d);
        for (int i = j; i < m; ++i) cj[i] /= s;
    }
    return 0;
}
void chol_solve(const double* L, int m, double* b) {          // b <- (L L^T)^-1 b
    for (int j = 0; j < m; ++j) {
        b[j] /= L[(size_t)j * m + j];
        const double bj = b[j];
        const double* c = L + (size_t)j * m;
        for (int i = j + 1; i < m; ++i) b[i] -= c[i] * bj;
    }
    for (int j = m - 1; j >= 0; --j) {
        const double* c = L + (size_t)j * m;
        double s = b[j];
        for (int i = j + 1; i < m; ++i) s -= c[i] * b[i];
        b[j] = s / c[j];
    }
}

// Standardised genotype matrix for a SNP list: decode + impute + z-score per column
// (calcBlock loop, dbslmmfit.cpp:419-426 / 456-463).  col-major n_ref x m.
void load_geno(const uint8_t* bed, int n_ref, const int32_t* pos, int m, std::vector<double>& X) {
    X.assign((size_t)n_ref * m, 0.0);
    for (int j = 0; j < m; ++j) {
        double maf;
        read_snp_im(bed, pos[j], n_ref, nullptr, X.data() + (size_t)j * n_ref, &maf);
        normalize_vec(X.data() + (size_t)j * n_ref, n_ref);
    }
}

struct BlockOut { int iters_max = 0; int singular = 0; };

// ---- DBSLMMFIT::estBlock, both overloads (dbslmmfit.cpp:680-738, 740-770) -----------------
// mode 0 = "ref": PCG + the reference's cancellation form for beta_s.
// mode 1 = "exact": Cholesky solve of the same systems + direct form beta_s = t / N.
BlockOut est_block(const uint8_t* bed, int n_ref, int n_obs, double sigma_s, double tau,
                   const int32_t* pos_s, const double* z_s, int ms,
                   const int32_t* pos_l, const double* z_l, int ml, int mode,
                   double* beta_s, double* beta_l, double* sigma_ss_out) {
    BlockOut out;
    if (ms == 0) return out;
    const double dn = (double)n_obs, sq = std::sqrt((double)n_obs);
    std::vector<double> Xs, Xl;
    load_geno(bed, n_ref, pos_s, ms, Xs);
    std::vector<double> Sss((size_t)ms * ms);
    atb(Xs.data(), ms, Xs.data(), ms, n_ref, Sss.data());                 // :705 / :752
    for (auto& v : Sss) v *= tau / (double)n_ref;                          // :706 / :753
    for (int i = 0; i < ms; ++i) Sss[(size_t)i * ms + i] += (1.0 - tau);   // :707-709 / :754-756
    if (sigma_ss_out) std::memcpy(sigma_ss_out, Sss.data(), sizeof(double) * ms * ms);
    const double ridge = 1.0 / (sigma_s * dn);
    auto solveA = [&](const std::vector<double>& A, const double* b, double* x,
                      const std::vector<double>* Lfac) {
        if (mode == 0) {
            const int it = pcgv(A.data(), b, ms, 1000, 1e-7, x);
            out.iters_max = std::max(out.iters_max, it);
            if (it >= 1000) out.singular++;
        } else {
            std::memcpy(x, b, sizeof(double) * ms);
            chol_solve(Lfac->data(), ms, x);
        }
    };
    if (ml == 0) {
        // ---- small-only overload (:740-770)
        std::vector<double> A = Sss;
        for (int i = 0; i < ms; ++i) A[(size_t)i * ms + i] += ridge;       // :759
        std::vector<double> L;
        if (mode == 1) { L = A; if (chol_lower(L.data(), ms)) out.singular++; }
        std::vector<double> u(ms), t(ms);
        solveA(A, z_s, u.data(), &L);                                      // :760
        if (mode == 0) {
            matvec(Sss.data(), u.data(), ms, t.data());                    // :762
            for (int i = 0; i < ms; ++i) beta_s[i] = sq * sigma_s * (z_s[i] - t[i]);   // :763-764
        } else {
            for (int i = 0; i < ms; ++i) beta_s[i] = u[i] / sq;            // algebraic identity, SURVEY 8a a6
        }
        return out;
    }
    // ---- large + small overload (:680-738)
    load_geno(bed, n_ref, pos_l, ml, Xl);
    std::vector<double> Sls((size_t)ml * ms), Sll((size_t)ml * ml);
    atb(Xl.data(), ml, Xs.data(), ms, n_ref, Sls.data());                  // :698  (ml x ms col-major)
    for (auto& v : Sls) v *= tau / (double)n_ref;                          // :699
    atb(Xl.data(), ml, Xl.data(), ml, n_ref, Sll.data());                  // :700
    for (auto& v : Sll) v *= tau / (double)n_ref;                          // :701
    for (int i = 0; i < ml; ++i) Sll[(size_t)i * ml + i] += (1.0 - tau);   // :702-704
    std::vector<double> A = Sss;
    for (int i = 0; i < ms; ++i) A[(size_t)i * ms + i] += ridge;           // :712
    std::vector<double> L;
    if (mode == 1) { L = A; if (chol_lower(L.data(), ms)) out.singular++; }
    // W = A^-1 Sigma_sl, column by column (PCGm :670-678, call :713)
    std::vector<double> W((size_t)ms * ml), col(ms);
    for (int c = 0; c < ml; ++c) {
        for (int i = 0; i < ms; ++i) col[i] = Sls[(size_t)i * ml + c];     // (Sigma_ls^T).col(c)
        solveA(A, col.data(), W.data() + (size_t)c * ms, &L);
    }
    // S = Sigma_ll - Sigma_ls W   (:714-715)
    std::vector<double> S((size_t)ml * ml);
    for (int a = 0; a < ml; ++a)
        for (int b = 0; b < ml; ++b) {
            double s = 0.0;
            for (int i = 0; i < ms; ++i) s += Sls[(size_t)i * ml + a] * W[(size_t)b * ms + i];
            S[(size_t)b * ml + a] = -s + Sll[(size_t)b * ml + a];
        }
    std::vector<double> u(ms);
    solveA(A, z_s, u.data(), &L);                                          // :716
    std::vector<double> rhs(ml);
    for (int a = 0; a < ml; ++a) {                                         // :717-718
        double s = 0.0;
        for (int i = 0; i < ms; ++i) s += Sls[(size_t)i * ml + a] * u[i];
        rhs[a] = -s + z_l[a];
    }
    if (mode == 0) {
        const int it = pcgv(S.data(), rhs.data(), ml, 1000, 1e-7, beta_l); // :719
        out.iters_max = std::max(out.iters_max, it);
        if (it >= 1000) out.singular++;
    } else {
        std::vector<double> LS = S;
        if (chol_lower(LS.data(), ml)) out.singular++;
        std::memcpy(beta_l, rhs.data(), sizeof(double) * ml);
        chol_solve(LS.data(), ml, beta_l);
    }
    for (int a = 0; a < ml; ++a) beta_l[a] /= sq;                          // :720
    // beta_s (:723-729)
    std::vector<double> t(ms), Wb(ms, 0.0);
    for (int c = 0; c < ml; ++c)
        for (int i = 0; i < ms; ++i) Wb[i] += W[(size_t)c * ms + i] * beta_l[c];
    for (int i = 0; i < ms; ++i) t[i] = sq * u[i] - dn * Wb[i];            // :723-725
    if (mode == 0) {
        std::vector<double> St(ms);
        matvec(Sss.data(), t.data(), ms, St.data());                       // :727 (ridge removed :726)
        for (int i = 0; i < ms; ++i) {
            double s = 0.0;
            for (int a = 0; a < ml; ++a) s += Sls[(size_t)i * ml + a] * beta_l[a];
            beta_s[i] = sigma_s * (sq * z_s[i] - dn * s - St[i]);          // :728-729
        }
    } else {
        for (int i = 0; i < ms; ++i) beta_s[i] = t[i] / dn;                // SURVEY 8a a7 identity
    }
    return out;
}

}  // namespace

// ============================== extern "C" surface =====================================

ORC_API int orc_read_snp_im(const uint8_t* bed, int64_t pos, int n_total, const int* indicator,
                            double* geno, double* maf) {
    return read_snp_im(bed, pos, n_total, indicator, geno, maf);
}
ORC_API void orc_normalize(double* x, int n) { normalize_vec(x, n); }

// Raw-code integer Gram the CUDA correlation builder must reproduce bit-exactly
// (SURVEY 8a "K3 exact-integer restatement"): g in {0,1,2} with missing -> 0, mask M.
// Q = G G^T, A_ij = sum g_i M_j, N = M M^T, all row-major m x m int32.
ORC_API void orc_gram_int(const uint8_t* bed, int n_ref, const int32_t* pos, int m,
                          int32_t* Q, int32_t* A, int32_t* N) {
    const int64_t pitch = bed_pitch(n_ref);
    std::vector<int8_t> G((size_t)m * n_ref), M((size_t)m * n_ref);
    for (int j = 0; j < m; ++j) {
        const uint8_t* row = bed + pos[j] * pitch;
        for (int s = 0; s < n_ref; ++s) {
            const unsigned c = (row[s >> 2] >> (2 * (s & 3))) & 3u;   // same bit order as :329-350
            G[(size_t)j * n_ref + s] = (c == 0) ? 2 : (c == 2) ? 1 : 0;
            M[(size_t)j * n_ref + s] = (c == 1) ? 0 : 1;
        }
    }
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            int32_t q = 0, a = 0, nn = 0;
            const int8_t *gi = &G[(size_t)i * n_ref], *gj = &G[(size_t)j * n_ref];
            const int8_t *mi = &M[(size_t)i * n_ref], *mj = &M[(size_t)j * n_ref];
            for (int s = 0; s < n_ref; ++s) { q += gi[s] * gj[s]; a += gi[s] * mj[s]; nn += mi[s] * mj[s]; }
            if (Q) Q[(size_t)i * m + j] = q;
            if (A) A[(size_t)i * m + j] = a;
            if (N) N[(size_t)i * m + j] = nn;
        }
}

// Per-SNP statistics of the MAF pre-pass (readBim, dtpr.cpp:93-102): maf from readSNPIm.
ORC_API void orc_snp_maf(const uint8_t* bed, int64_t n_snp, int n_ref, double* maf) {
#pragma omp parallel
    {
        std::vector<double> g(n_ref);
#pragma omp for
        for (int64_t i = 0; i < n_snp; ++i) read_snp_im(bed, i, n_ref, nullptr, g.data(), &maf[i]);
    }
}

// Sigma = tau X^T X / n + (1-tau) I through the reference's float path (col-major m x m).
ORC_API void orc_sigma(const uint8_t* bed, int n_ref, const int32_t* pos, int m, double tau,
                       double* sigma) {
    std::vector<double> X;
    load_geno(bed, n_ref, pos, m, X);
    atb(X.data(), m, X.data(), m, n_ref, sigma);
    for (size_t i = 0; i < (size_t)m * m; ++i) sigma[i] *= tau / (double)n_ref;
    for (int i = 0; i < m; ++i) sigma[(size_t)i * m + i] += (1.0 - tau);
}

ORC_API int orc_pcgv(const double* A, const double* b, int m, int maxiter, double tol, double* x) {
    return pcgv(A, b, m, maxiter, tol, x);
}

// One block.  Returns max PCG iterations (ref mode) or 0; *singular counts failed solves.
ORC_API int orc_est_block(const uint8_t* bed, int n_ref, int n_obs, double sigma_s, double tau,
                          const int32_t* pos_s, const double* z_s, int ms,
                          const int32_t* pos_l, const double* z_l, int ml, int mode,
                          double* beta_s, double* beta_l, int* singular) {
    BlockOut o = est_block(bed, n_ref, n_obs, sigma_s, tau, pos_s, z_s, ms, pos_l, z_l, ml, mode,
                           beta_s, beta_l, nullptr);
    if (singular) *singular = o.singular;
    return o.iters_max;
}

// ---- DBSLMMFIT::est (dbslmmfit.cpp:56-244 / 247-363): batches of B_MAX = min(60, n_blocks)
// blocks, each batch an OpenMP `parallel for schedule(dynamic)`; results land block-major.
// CSR inputs: block b owns small SNPs [s_off[b], s_off[b+1]) and large [l_off[b], l_off[b+1]).
// l_off == NULL selects the small-only overload for every block.
ORC_API int orc_est(const uint8_t* bed, int n_ref, int n_obs, double sigma_s, double tau,
                    int n_blocks, const int32_t* s_off, const int32_t* s_pos, const double* s_z,
                    const int32_t* l_off, const int32_t* l_pos, const double* l_z,
                    int threads, int mode, double* beta_s, double* beta_l, int* max_iters) {
    int B_MAX = 60;                                                        // :93-96
    if (n_blocks < 60) B_MAX = n_blocks;
    int singular = 0, itmax = 0;
#ifdef _OPENMP
    omp_set_num_threads(threads);                                          // :191
#endif
    for (int start = 0; start < n_blocks; start += B_MAX) {
        const int B = std::min(B_MAX, n_blocks - start);
#pragma omp parallel for schedule(dynamic) reduction(+ : singular) reduction(max : itmax)   // :192
        for (int bb = 0; bb < B; ++bb) {
            const int b = start + bb;
            const int ms = s_off[b + 1] - s_off[b];
            const int ml = l_off ? l_off[b + 1] - l_off[b] : 0;
            BlockOut o = est_block(bed, n_ref, n_obs, sigma_s, tau, s_pos + s_off[b], s_z + s_off[b], ms,
                                   l_off ? l_pos + l_off[b] : nullptr, l_off ? l_z + l_off[b] : nullptr,
                                   ml, mode, beta_s + s_off[b], l_off ? beta_l + l_off[b] : nullptr, nullptr);
            singular += o.singular;
            itmax = std::max(itmax, o.iters_max);
        }
    }
    if (max_iters) *max_iters = itmax;
    return singular;
}

// ---- the fork's asymptotic-variance side channel for ONE block ------------------
// calcBlock's test-genotype loop (dbslmmfit.cpp:427-429, 464-466: readSNPIm with the test indicator +
// nomalizeVec WITHIN the test subset) and calc_nt_by_nt_matrix (calc_asymptotic_variance.cpp:22-57):
//   Ainv = (I/(n sigma2) + Sigma_ss)^-1,  var_bl = (Sigma_ll - Sigma_sl' Ainv Sigma_sl)^-1 / n,
//   var_bs = n sigma2^2 (mat1 + mat2 n var_bl mat2'),  mat1 = Sigma_ss - Sigma_ss Ainv Sigma_ss,
//   mat2 = Sigma_sl - Sigma_ss Ainv Sigma_sl;  out_i = (X_l var_bl X_l' + X_s var_bs X_s')_ii.
// Dense and O(m^3): for test sizes only.
ORC_API int orc_variance_block(const uint8_t* bed, int n_ref, const uint8_t* tbed, int n_test_total,
                               const int* indicator, int n_obs, double sigma_s, double tau,
                               const int32_t* pos_s, const int32_t* tpos_s, int ms,
                               const int32_t* pos_l, const int32_t* tpos_l, int ml, double* out) {
    int n_test = 0;
    for (int i = 0; i < n_test_total; ++i) n_test += (indicator[i] != 0);
    const int m = ms + ml;
    std::vector<int32_t> pos(m);
    for (int j = 0; j < ms; ++j) pos[j] = pos_s[j];
    for (int j = 0; j < ml; ++j) pos[ms + j] = pos_l[j];
    std::vector<double> X;                                   // reference panel, all block SNPs (small first)
    load_geno(bed, n_ref, pos.data(), m, X);
    std::vector<double> Sg((size_t)m * m);
    atb(X.data(), m, X.data(), m, n_ref, Sg.data());
    for (auto& v : Sg) v *= tau / (double)n_ref;
    for (int i = 0; i < m; ++i) Sg[(size_t)i * m + i] += (1.0 - tau);
    auto S = [&](int i, int j) -> double { return Sg[(size_t)j * m + i]; };
    // standardised test genotypes, n_test x m (col-major)
    std::vector<double> T((size_t)n_test * m);
    for (int j = 0; j < m; ++j) {
        double maf;
        const int32_t tp = (j < ms) ? tpos_s[j] : tpos_l[j - ms];
        read_snp_im(tbed, tp, n_test_total, indicator, T.data() + (size_t)j * n_test, &maf);
        normalize_vec(T.data() + (size_t)j * n_test, n_test);
    }
    const double dn = (double)n_obs;
    // Ainv via Cholesky of A
    std::vector<double> A((size_t)ms * ms), Ainv((size_t)ms * ms, 0.0);
    for (int j = 0; j < ms; ++j) for (int i = 0; i < ms; ++i) A[(size_t)j * ms + i] = S(i, j) + (i == j ? 1.0 / (dn * sigma_s) : 0.0);
    if (chol_lower(A.data(), ms)) return 1;
    for (int c = 0; c < ms; ++c) { double* col = &Ainv[(size_t)c * ms]; col[c] = 1.0; chol_solve(A.data(), ms, col); }
    auto AI = [&](int i, int j) -> double { return Ainv[(size_t)j * ms + i]; };
    // G = Ainv Sigma_sl (ms x ml), SA = Sigma_ss Ainv (ms x ms)
    std::vector<double> G((size_t)ms * std::max(ml, 1), 0.0), SA((size_t)ms * ms, 0.0);
    for (int c = 0; c < ml; ++c) for (int i = 0; i < ms; ++i) { double s = 0; for (int k = 0; k < ms; ++k) s += AI(i, k) * S(k, ms + c); G[(size_t)c * ms + i] = s; }
    for (int j = 0; j < ms; ++j) for (int i = 0; i < ms; ++i) { double s = 0; for (int k = 0; k < ms; ++k) s += S(i, k) * AI(k, j); SA[(size_t)j * ms + i] = s; }
    std::vector<double> var_bl((size_t)std::max(ml, 1) * std::max(ml, 1), 0.0);
    if (ml > 0) {
        std::vector<double> big((size_t)ml * ml);
        for (int b = 0; b < ml; ++b) for (int a = 0; a < ml; ++a) { double s = 0; for (int i = 0; i < ms; ++i) s += S(i, ms + a) * G[(size_t)b * ms + i]; big[(size_t)b * ml + a] = S(ms + a, ms + b) - s; }
        if (chol_lower(big.data(), ml)) return 2;
        for (int c = 0; c < ml; ++c) { double* col = &var_bl[(size_t)c * ml]; col[c] = 1.0; chol_solve(big.data(), ml, col); for (int a = 0; a < ml; ++a) col[a] /= dn; }
    }
    // var_bs = n sigma^2 sigma^2 (mat1 + mat2 n var_bl mat2')
    std::vector<double> mat1((size_t)ms * ms), mat2((size_t)ms * std::max(ml, 1), 0.0), var_bs((size_t)ms * ms);
    for (int j = 0; j < ms; ++j) for (int i = 0; i < ms; ++i) { double s = 0; for (int k = 0; k < ms; ++k) s += SA[(size_t)k * ms + i] * S(k, j); mat1[(size_t)j * ms + i] = S(i, j) - s; }
    for (int c = 0; c < ml; ++c) for (int i = 0; i < ms; ++i) { double s = 0; for (int k = 0; k < ms; ++k) s += SA[(size_t)k * ms + i] * S(k, ms + c); mat2[(size_t)c * ms + i] = S(i, ms + c) - s; }
    for (int j = 0; j < ms; ++j) for (int i = 0; i < ms; ++i) {
        double s = 0;
        for (int a = 0; a < ml; ++a) for (int b = 0; b < ml; ++b) s += mat2[(size_t)a * ms + i] * dn * var_bl[(size_t)b * ml + a] * mat2[(size_t)b * ms + j];
        var_bs[(size_t)j * ms + i] = dn * sigma_s * sigma_s * (mat1[(size_t)j * ms + i] + s);
    }
    for (int t = 0; t < n_test; ++t) {
        double d = 0;
        for (int a = 0; a < ml; ++a) for (int b = 0; b < ml; ++b) d += T[(size_t)(ms + a) * n_test + t] * var_bl[(size_t)b * ml + a] * T[(size_t)(ms + b) * n_test + t];
        for (int i = 0; i < ms; ++i) { double s = 0; for (int j = 0; j < ms; ++j) s += var_bs[(size_t)j * ms + i] * T[(size_t)j * n_test + t]; d += T[(size_t)i * n_test + t] * s; }
        out[t] = d;
    }
    return 0;
}

ORC_API int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
