// oracle/eigen_qr_bench.cpp — TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
//
// Times Eigen::HouseholderQR, the library call of the reference's CPU toy (C++/main.cpp:54:
// `Eigen::HouseholderQR<Eigen::MatrixXd> householderQR(A);`), on a seeded uniform[0,1) m x n matrix in
// float (the arithmetic of the GPU path; the toy's own literal is a 3 x 3 double matrix).  Compiled by
// oracle/build_ref.sh against the Eigen 3.4.0 headers vendored in the reference tree
// (Cuda/QR/Solver/Eigen, read in place) into oracle/_ref/eigen_qr.  Single thread: Eigen's
// HouseholderQR is the unblocked level-2 algorithm and the reference builds without OpenMP.
//
//   eigen_qr m n seed reps  ->  one JSON line: seconds per factorisation (best of reps), backward error
#include <Eigen/Dense>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: eigen_qr m n seed [reps]\n"); return 2; }
    const long m = atol(argv[1]), n = atol(argv[2]);
    const uint64_t seed = strtoull(argv[3], nullptr, 10);
    const int reps = argc > 4 ? atoi(argv[4]) : 1;
    typedef Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> Mat;
    Mat A(m, n);
    const uint64_t s = mix64(seed);  // same stateless generator as oracle/mpqr_oracle.c:orc_uniform01
    for (long i = 0; i < m; ++i)
        for (long j = 0; j < n; ++j) A(i, j) = (float)(mix64(s + (uint64_t)(i * n + j)) >> 40) * (1.0f / 16777216.0f);
    double best = 1e300, err = 0;
    for (int it = 0; it < reps; ++it) {
        auto t0 = std::chrono::steady_clock::now();
        Eigen::HouseholderQR<Mat> qr(A);
        auto t1 = std::chrono::steady_clock::now();
        const double dt = std::chrono::duration<double>(t1 - t0).count();
        if (dt < best) best = dt;
        if (it == 0 && m * n <= 1024L * 1024L) {
            Mat R = qr.matrixQR().template triangularView<Eigen::Upper>();
            Mat QR = qr.householderQ() * R;
            err = (A - QR).norm() / A.norm();
        }
    }
    printf("{\"m\": %ld, \"n\": %ld, \"seconds\": %.6f, \"threads\": 1, \"backward_error\": %.3e, \"eigen\": \"%d.%d.%d\"}\n", m, n, best, err,
           EIGEN_WORLD_VERSION, EIGEN_MAJOR_VERSION, EIGEN_MINOR_VERSION);
    return 0;
}
