// loader.cu — EuRoC Jacobian text reader (host only, no CUDA calls).
//
// Replaces read_euroc_jacobian (reference Cuda/qr.cu:696-776): line 1 "<rows> <cols>", then one
// "<row> <col> <value>" triple per line (0-based, whitespace separated, later entries overwrite
// earlier ones) into a zero-filled dense row-major FP32 matrix.  The buffer is returned in the
// PACKED layout the drivers take, (rows + 1) x cols with the extra row zero
// (Cuda/qr.cu:1866-1875), so it can go straight into mpqr_block_qr_host.
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

extern "C" int mpqr_read_euroc_jacobian(const char* path, int* rows, int* cols, float** packed_out) {
    if (!path || !rows || !cols || !packed_out) {
        mpqr::set_error("mpqr_read_euroc_jacobian: null argument");
        return MPQR_EINVAL;
    }
    FILE* f = fopen(path, "r");
    if (!f) {
        mpqr::set_error("mpqr_read_euroc_jacobian: cannot open %s: %s", path, strerror(errno));
        return MPQR_EINVAL;
    }
    long m = 0, n = 0;
    if (fscanf(f, "%ld %ld", &m, &n) != 2 || m < 1 || n < 1 || m > 0x7ffffff0L || n > 0x7ffffff0L) {
        fclose(f);
        mpqr::set_error("mpqr_read_euroc_jacobian: %s: bad header (expected \"<rows> <cols>\")", path);
        return MPQR_EINVAL;
    }
    float* A = (float*)calloc((size_t)(m + 1) * (size_t)n, sizeof(float));
    if (!A) {
        fclose(f);
        mpqr::set_error("mpqr_read_euroc_jacobian: out of host memory (%ld x %ld)", m, n);
        return MPQR_ENOMEM;
    }
    long i, j, line = 1;
    double v;
    int got;
    while ((got = fscanf(f, "%ld %ld %lf", &i, &j, &v)) == 3) {
        ++line;
        if (i < 0 || i >= m || j < 0 || j >= n) {
            fclose(f);
            free(A);
            mpqr::set_error("mpqr_read_euroc_jacobian: %s line %ld: index (%ld, %ld) outside %ld x %ld", path, line, i, j, m, n);
            return MPQR_EINVAL;
        }
        A[(size_t)i * n + j] = (float)v;
    }
    const int clean_eof = (got == EOF);
    fclose(f);
    if (!clean_eof) {
        free(A);
        mpqr::set_error("mpqr_read_euroc_jacobian: %s: malformed entry after line %ld", path, line);
        return MPQR_EINVAL;
    }
    *rows = (int)m;
    *cols = (int)n;
    *packed_out = A;
    return MPQR_OK;
}

extern "C" void mpqr_free_host(void* p) { free(p); }
