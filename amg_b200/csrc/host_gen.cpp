// Synthetic level-0 operators.  Pure host code, no device.
//
// Common conventions: grid index i = x + N*y + N*N*z (x fastest), Dirichlet boundary
// (neighbours outside the grid dropped, diagonal unchanged), CSR rows with ascending
// columns -- the order the reference's loader (amg/mmio_highlevel.h:289-295) produces for a
// row-major, column-ascending MatrixMarket file.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/amg_b200.h"

namespace {

uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int alloc_csr(amgb200_mat *A, long long n, long long nnz) {
    if (n <= 0 || nnz <= 0 || nnz > 2147483647LL || n > 2147483646LL) return -1;
    A->num_rows = A->num_cols = (int)n;
    A->num_nnzs = (int)nnz;
    A->row_ptr = (int *)malloc((size_t)(n + 1) * sizeof(int));
    A->col_idx = (int *)malloc((size_t)nnz * sizeof(int));
    A->val = (double *)malloc((size_t)nnz * sizeof(double));
    if (!A->row_ptr || !A->col_idx || !A->val) return -2;
    return 0;
}

// 5-/7-point constant-coefficient stencils: coupling -cx,-cy,-cz; diagonal 2(cx+cy+cz) (3D) or 2(cx+cy) (2D)
int gen_stencil(int dim, int N, double cx, double cy, double cz, amgb200_mat *A) {
    const long long n = dim == 2 ? (long long)N * N : (long long)N * N * N;
    const long long nnz = dim == 2 ? 5LL * N * N - 4LL * N : 7LL * N * N * N - 6LL * N * N;
    if (alloc_csr(A, n, nnz)) return -1;
    const int Nz = dim == 2 ? 1 : N;
    // same association as summing the three Kronecker terms left to right
    const double diag = dim == 2 ? (2.0 * cx + 2.0 * cy) : ((2.0 * cx + 2.0 * cy) + 2.0 * cz);
    long long p = 0;
    for (int z = 0; z < Nz; ++z)
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) {
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

// 27-point variable-coefficient diffusion (M-matrix): kappa_i = 10^(2u_i-1), edge weight = harmonic mean
int gen_v27(int N, amgb200_mat *A) {
    const long long n = (long long)N * N * N;
    const long long e = 3LL * N - 2;
    const long long nnz = e * e * e;
    if (alloc_csr(A, n, nnz)) return -1;
    std::vector<double> kap((size_t)n);
    for (long long i = 0; i < n; ++i) {
        const uint64_t h = splitmix64((uint64_t)i + 1234ull * 0x100000001B3ull);
        const double u = (double)(h >> 11) / 9007199254740992.0;   // 2^53
        kap[(size_t)i] = pow(10.0, 2.0 * u - 1.0);
    }
    long long p = 0;
    for (int z = 0; z < N; ++z)
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) {
                const long long i = x + (long long)N * y + (long long)N * N * z;
                A->row_ptr[i] = (int)p;
                const double ki = kap[(size_t)i];
                double offsum = 0.0;
                int missing = 0;
                long long dpos = -1;
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int xx = x + dx, yy = y + dy, zz = z + dz;
                            if (dx == 0 && dy == 0 && dz == 0) { dpos = p; A->col_idx[p] = (int)i; A->val[p++] = 0.0; continue; }
                            if (xx < 0 || xx >= N || yy < 0 || yy >= N || zz < 0 || zz >= N) { ++missing; continue; }
                            const long long j = xx + (long long)N * yy + (long long)N * N * zz;
                            const double kj = kap[(size_t)j];
                            const double a = -2.0 * ki * kj / (ki + kj);
                            A->col_idx[p] = (int)j; A->val[p++] = a;
                            offsum += fabs(a);
                        }
                A->val[dpos] = offsum + ki * (double)missing;
            }
    A->row_ptr[n] = (int)p;
    return p == nnz ? 0 : -3;
}

}  // namespace

extern "C" int amgb200_generate(int kind, int N, double eps_z, amgb200_mat *A) {
    if (!A || N < 2) return -1;
    memset(A, 0, sizeof(*A));
    switch (kind) {
        case 0: return gen_stencil(2, N, 1.0, 1.0, 0.0, A);
        case 1: return gen_stencil(3, N, 1.0, 1.0, 1.0, A);
        case 2: return gen_stencil(3, N, 1.0, 1.0, eps_z, A);
        case 3: return gen_v27(N, A);
        default: return -1;
    }
}

extern "C" void amgb200_mat_free(amgb200_mat *A) {
    if (!A) return;
    free(A->row_ptr); free(A->col_idx); free(A->val);
    A->row_ptr = nullptr; A->col_idx = nullptr; A->val = nullptr;
}
