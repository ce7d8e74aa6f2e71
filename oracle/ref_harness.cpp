// oracle/ref_harness.cpp -- extern "C" doorway into the UNMODIFIED reference classes
// (IO, SNPPROC, DBSLMMFIT from /root/reference/scr, compiled over oracle/shim).
// TEST INFRASTRUCTURE: used to pin the oracle (tests/test_oracle.py), to generate the
// golden vectors under tests/golden/ and as bench.py's `--impl reference` arm.
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "omp.h"
#include "dtpr.hpp"
#include "dbslmmfit.hpp"
#include "calc_asymptotic_variance.hpp"

#define REF_API extern "C" __attribute__((visibility("default")))

// IO::readSNPIm on a .bed FILE (the reference reads through an ifstream)
REF_API int ref_read_snp_im(const char* bed_path, int pos, int n_total, const int* indicator, double* geno, double* maf) {
    IO io;
    std::ifstream in(bed_path, std::ios::binary);
    if (!in) return -1;
    std::vector<int> ind(indicator, indicator + n_total);
    int keep = 0;
    for (int v : ind) keep += (v != 0);
    arma::vec g = arma::zeros<arma::vec>(keep);
    double m = 0.0;
    io.readSNPIm(pos, keep, ind, in, g, m);
    std::memcpy(geno, g.memptr(), sizeof(double) * keep);
    *maf = m;
    return keep;
}

REF_API void ref_normalize(double* x, int n) {
    SNPPROC sp;
    arma::vec v(n);
    std::memcpy(v.memptr(), x, sizeof(double) * n);
    sp.nomalizeVec(v);
    std::memcpy(x, v.memptr(), sizeof(double) * n);
}

REF_API void ref_pcgv(const double* A, const double* b, int m, int maxiter, double tol, double* x) {
    DBSLMMFIT f;
    arma::mat Am(m, m);
    std::memcpy(Am.d.data(), A, sizeof(double) * m * m);
    arma::vec bv(m);
    std::memcpy(bv.memptr(), b, sizeof(double) * m);
    arma::vec xv = f.PCGv(Am, bv, (size_t)maxiter, tol);
    std::memcpy(x, xv.memptr(), sizeof(double) * m);
}

static arma::mat load_cols(const char* bed_path, int n_ref, const int32_t* pos, int m) {
    IO io; SNPPROC sp;
    std::ifstream in(bed_path, std::ios::binary);
    std::vector<int> idv(n_ref, 1);
    arma::mat X = arma::zeros<arma::mat>(n_ref, m);
    for (int j = 0; j < m; ++j) {                       // calcBlock's column loop, dbslmmfit.cpp:419-426
        arma::vec g = arma::zeros<arma::vec>(n_ref);
        double maf = 0.0;
        io.readSNPIm(pos[j], n_ref, idv, in, g, maf);
        sp.nomalizeVec(g);
        X.col(j) = g;
    }
    return X;
}

// readSNPIm + nomalizeVec + DBSLMMFIT::estBlock for one block (both overloads)
REF_API int ref_est_block(const char* bed_path, int n_ref, int n_obs, double sigma_s,
                          const int32_t* pos_s, const double* z_s, int ms,
                          const int32_t* pos_l, const double* z_l, int ml, double* beta_s, double* beta_l) {
    DBSLMMFIT f;
    arma::mat Xs = load_cols(bed_path, n_ref, pos_s, ms);
    arma::vec zs(ms), bs = arma::zeros<arma::vec>(ms);
    std::memcpy(zs.memptr(), z_s, sizeof(double) * ms);
    if (ml > 0) {
        arma::mat Xl = load_cols(bed_path, n_ref, pos_l, ml);
        arma::vec zl(ml), bl = arma::zeros<arma::vec>(ml);
        std::memcpy(zl.memptr(), z_l, sizeof(double) * ml);
        f.estBlock(n_ref, n_obs, sigma_s, Xs, Xl, zs, zl, bs, bl);
        std::memcpy(beta_l, bl.memptr(), sizeof(double) * ml);
    } else {
        f.estBlock(n_ref, n_obs, sigma_s, Xs, zs, bs);
    }
    std::memcpy(beta_s, bs.memptr(), sizeof(double) * ms);
    return 0;
}

// The north-star path under the reference's own schedule: batches of min(60, n_blocks) blocks,
// `omp parallel for schedule(dynamic)` (dbslmmfit.cpp:92-96, 189-193), each block = the calls
// above.  (DBSLMMFIT::est itself also runs the fork's variance side-channel, which is outside
// the path; ref_est_full below runs the unmodified est for completeness.)
REF_API int ref_est_path(const char* bed_path, int n_ref, int n_obs, double sigma_s, int n_blocks,
                         const int32_t* s_off, const int32_t* s_pos, const double* s_z,
                         const int32_t* l_off, const int32_t* l_pos, const double* l_z,
                         int threads, double* beta_s, double* beta_l) {
    int B_MAX = 60;
    if (n_blocks < 60) B_MAX = n_blocks;
    omp_set_num_threads(threads);
    for (int start = 0; start < n_blocks; start += B_MAX) {
        const int B = std::min(B_MAX, n_blocks - start);
#pragma omp parallel for schedule(dynamic)
        for (int bb = 0; bb < B; ++bb) {
            const int b = start + bb;
            const int ms = s_off[b + 1] - s_off[b];
            const int ml = l_off ? l_off[b + 1] - l_off[b] : 0;
            if (ms == 0) continue;
            ref_est_block(bed_path, n_ref, n_obs, sigma_s, s_pos + s_off[b], s_z + s_off[b], ms,
                          l_off ? l_pos + l_off[b] : nullptr, l_off ? l_z + l_off[b] : nullptr, ml,
                          beta_s + s_off[b], l_off ? beta_l + l_off[b] : nullptr);
        }
    }
    return 0;
}

// The fork's variance side channel for one block through the reference's own functions:
// test genotypes as calcBlock reads them (dbslmmfit.cpp:427-429, 464-466), estBlock's Sigma matrices,
// calc_nt_by_nt_matrix(...).diag() (calc_asymptotic_variance.cpp:22-57).
static arma::mat load_test_cols(const char* path, int n_total, const int* indicator, const int32_t* pos, int m) {
    IO io; SNPPROC sp;
    std::ifstream in(path, std::ios::binary);
    std::vector<int> ind(indicator, indicator + n_total);
    int n_test = 0;
    for (int v : ind) n_test += (v != 0);
    arma::mat X = arma::zeros<arma::mat>(n_test, m);
    for (int j = 0; j < m; ++j) {
        arma::vec g = arma::zeros<arma::vec>(n_test);
        double maf = 0.0;
        io.readSNPIm(pos[j], n_test, ind, in, g, maf);
        sp.nomalizeVec(g);
        X.col(j) = g;
    }
    return X;
}

REF_API int ref_variance_block(const char* bed_path, int n_ref, const char* tbed_path, int n_test_total,
                               const int* indicator, int n_obs, double sigma_s,
                               const int32_t* pos_s, const int32_t* tpos_s, int ms,
                               const int32_t* pos_l, const int32_t* tpos_l, int ml, double* out) {
    DBSLMMFIT f;
    arma::mat Xs = load_cols(bed_path, n_ref, pos_s, ms);
    arma::mat Ts = load_test_cols(tbed_path, n_test_total, indicator, tpos_s, ms);
    arma::vec zs = arma::zeros<arma::vec>(ms), bs = arma::zeros<arma::vec>(ms);
    arma::mat result;
    if (ml > 0) {
        arma::mat Xl = load_cols(bed_path, n_ref, pos_l, ml);
        arma::mat Tl = load_test_cols(tbed_path, n_test_total, indicator, tpos_l, ml);
        arma::vec zl = arma::zeros<arma::vec>(ml), bl = arma::zeros<arma::vec>(ml);
        arma::field<arma::mat> o = f.estBlock(n_ref, n_obs, sigma_s, Xs, Xl, zs, zl, bs, bl);
        result = calc_nt_by_nt_matrix(o(2), o(1), o(0), sigma_s, (unsigned)n_obs, Tl, Ts);   // as dbslmmfit.cpp:484-490
    } else {
        arma::field<arma::mat> o = f.estBlock(n_ref, n_obs, sigma_s, Xs, zs, bs);
        result = calc_nt_by_nt_matrix(o(0), sigma_s, (unsigned)n_obs, Ts);                    // :516-519
    }
    arma::vec d = result.diag();
    std::memcpy(out, d.memptr(), sizeof(double) * d.n_elem);
    return (int)d.n_elem;
}
