/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * Plain-C restatement of the synthetic level-0 operators of SURVEY.md Appendix B, used by `bench.py --impl reference`
 * and by the cpu_baseline leg so that the reference arm never loads libamgb200.so.  tests/test_oracle.py checks that the
 * CSR arrays are byte-identical to amgb200_generate's (the product's generator) on every operator family.
 *
 * Conventions: grid index i = x + N*y + N*N*z, Dirichlet boundary, rows with ascending columns -- the order the
 * reference's loader produces for a row-major, column-ascending MatrixMarket file (/root/reference/amg/mmio_highlevel.h:289-295,
 * entry point /root/reference/amg/SSS_main.c:12-22).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int num_rows, num_cols, num_nnzs; int *row_ptr; int *col_idx; double *val; } orc_mat;   /* SSS_MAT layout, SSS_main.h:95-105 */

static uint64_t splitmix64(uint64_t x) {
    uint64_t z;
    x += 0x9E3779B97F4A7C15ull;
    z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int alloc_csr(orc_mat *A, long long n, long long nnz) {
    if (n <= 0 || nnz <= 0 || nnz > 2147483647LL) return -1;
    A->num_rows = A->num_cols = (int)n;
    A->num_nnzs = (int)nnz;
    A->row_ptr = (int *)malloc((size_t)(n + 1) * sizeof(int));
    A->col_idx = (int *)malloc((size_t)nnz * sizeof(int));
    A->val = (double *)malloc((size_t)nnz * sizeof(double));
    return (A->row_ptr && A->col_idx && A->val) ? 0 : -2;
}

/* kind 0: p2d (T(x)I + I(x)T), 1: p3d, 2: aniso3d with coefficients (1, 1, eps_z), 3: v27 */
int orc_generate(int kind, int N, double eps_z, orc_mat *A) {
    long long p = 0, n, nnz;
    int x, y, z;
    memset(A, 0, sizeof(*A));
    if (N < 2) return -1;
    if (kind >= 0 && kind <= 2) {
        const int dim = kind == 0 ? 2 : 3, Nz = kind == 0 ? 1 : N;
        const double cx = 1.0, cy = 1.0, cz = kind == 0 ? 0.0 : (kind == 1 ? 1.0 : eps_z);
        const double diag = dim == 2 ? (2.0 * cx + 2.0 * cy) : ((2.0 * cx + 2.0 * cy) + 2.0 * cz);
        n = dim == 2 ? (long long)N * N : (long long)N * N * N;
        nnz = dim == 2 ? 5LL * N * N - 4LL * N : 7LL * N * N * N - 6LL * N * N;
        if (alloc_csr(A, n, nnz)) return -1;
        for (z = 0; z < Nz; ++z) for (y = 0; y < N; ++y) for (x = 0; x < N; ++x) {
            const long long i = x + (long long)N * y + (long long)N * N * z;
            A->row_ptr[i] = (int)p;
            if (dim == 3 && z > 0) { A->col_idx[p] = (int)(i - (long long)N * N); A->val[p++] = -cz; }
            if (y > 0) { A->col_idx[p] = (int)(i - N); A->val[p++] = -cy; }
            if (x > 0) { A->col_idx[p] = (int)(i - 1); A->val[p++] = -cx; }
            A->col_idx[p] = (int)i; A->val[p++] = diag;
            if (x < N - 1) { A->col_idx[p] = (int)(i + 1); A->val[p++] = -cx; }
            if (y < N - 1) { A->col_idx[p] = (int)(i + N); A->val[p++] = -cy; }
            if (dim == 3 && z < N - 1) { A->col_idx[p] = (int)(i + (long long)N * N); A->val[p++] = -cz; }
        }
        A->row_ptr[n] = (int)p;
        return p == nnz ? 0 : -3;
    }
    if (kind == 3) {
        /* kappa_i = 10^(2u_i - 1), u_i = (splitmix64(i + 1234*0x100000001B3) >> 11) / 2^53; a_ij = -2 k_i k_j / (k_i + k_j) for the 26
         * neighbours inside the grid; a_ii = sum |a_ij| + k_i * (#missing neighbours) */
        const long long e = 3LL * N - 2;
        double *kap;
        long long i;
        n = (long long)N * N * N;
        nnz = e * e * e;
        if (alloc_csr(A, n, nnz)) return -1;
        kap = (double *)malloc((size_t)n * sizeof(double));
        if (!kap) return -2;
        for (i = 0; i < n; ++i) {
            const uint64_t h = splitmix64((uint64_t)i + 1234ull * 0x100000001B3ull);
            const double u = (double)(h >> 11) / 9007199254740992.0;
            kap[i] = pow(10.0, 2.0 * u - 1.0);
        }
        for (z = 0; z < N; ++z) for (y = 0; y < N; ++y) for (x = 0; x < N; ++x) {
            const long long ii = x + (long long)N * y + (long long)N * N * z;
            const double ki = kap[ii];
            double offsum = 0.0;
            int missing = 0, dx, dy, dz;
            long long dpos = -1;
            A->row_ptr[ii] = (int)p;
            for (dz = -1; dz <= 1; ++dz) for (dy = -1; dy <= 1; ++dy) for (dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx, yy = y + dy, zz = z + dz;
                if (dx == 0 && dy == 0 && dz == 0) { dpos = p; A->col_idx[p] = (int)ii; A->val[p++] = 0.0; continue; }
                if (xx < 0 || xx >= N || yy < 0 || yy >= N || zz < 0 || zz >= N) { ++missing; continue; }
                {
                    const long long j = xx + (long long)N * yy + (long long)N * N * zz;
                    const double kj = kap[j];
                    const double a = -2.0 * ki * kj / (ki + kj);
                    A->col_idx[p] = (int)j; A->val[p++] = a;
                    offsum += fabs(a);
                }
            }
            A->val[dpos] = offsum + ki * (double)missing;
        }
        free(kap);
        A->row_ptr[n] = (int)p;
        return p == nnz ? 0 : -3;
    }
    return -1;
}

void orc_mat_free(orc_mat *A) {
    free(A->row_ptr); free(A->col_idx); free(A->val);
    A->row_ptr = 0; A->col_idx = 0; A->val = 0;
}
