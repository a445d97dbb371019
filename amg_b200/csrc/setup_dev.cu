// Device version of the step right before the hot path: direct-interpolation weights, coarse renumbering and truncation of P
// (SURVEY.md section 8 f1).  Replaces, bit for bit, the reference's interp_DIR + SSS_amg_interp_trunc
// (/root/reference/amg/Setup/SSS_inter.cu:400-547 and :16-102); the reference's own GPU attempt (DIR_Step_1, :104-210, launched
// <<<2048,64>>> at :322) covers only 131 072 rows -- here there is no cap.
//
// Every row is independent and its sums are accumulated by ONE thread in the reference's order (no FMA: the library is built with
// -fmad=false and the quotients are IEEE divisions), so the weights, the kept set and the rescaled values are identical to the CPU
// loop.  The only sequential pieces -- the numbering of the C points and the row offsets of the truncated matrix -- are two
// O(rows) prefix sums done on the host between the kernels.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/amg_b200.h"

namespace {

constexpr int TB = 256;
constexpr double TINY = 1e-20;            // SMALLFLOAT, SSS_main.h:34

#define DEV_OK(call)                                                                                          \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            fprintf(stderr, "libamgb200 (device setup): CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(70);                                                                                         \
        }                                                                                                     \
    } while (0)

// SSS_inter.cu:420-507: weights of one row.  P column indices are still FINE indices of C points here.
__global__ void __launch_bounds__(TB) interp_weights_kernel(int n, const int *__restrict__ arp, const int *__restrict__ aci, const double *__restrict__ av,
                                                            const int *__restrict__ mark, const int *__restrict__ prp, const int *__restrict__ pci,
                                                            double *pv, int *nodiag) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const int b = arp[i], e = arp[i + 1], pb = prp[i], pe = prp[i + 1];
    int dk = b;
    for (; dk < e; ++dk) if (aci[dk] == i) break;
    if (dk == e) { atomicAdd(nodiag, 1); return; }                 // (the reference would carry the previous row's diagonal: host fallback)
    double aii = av[dk];
    if (mark[i] == 0) {                                            // FGPT
        double amN = 0.0, amP = 0.0, apN = 0.0, apP = 0.0;
        int npos = 0;
        for (int j = b; j < e; ++j) {
            if (j == dk) continue;
            const int c = aci[j];
            bool strong = false;
            for (int k = pb; k < pe; ++k) if (pci[k] == c) { strong = true; break; }
            const double a = av[j];
            if (a > 0) { apN = __dadd_rn(apN, a); if (strong) { apP = __dadd_rn(apP, a); ++npos; } }
            else { amN = __dadd_rn(amN, a); if (strong) amP = __dadd_rn(amP, a); }
        }
        const double alpha = __ddiv_rn(amN, amP);
        double beta;
        if (npos > 0) beta = __ddiv_rn(apN, apP);
        else { beta = 0.0; aii = __dadd_rn(aii, apN); }
        for (int q = pb; q < pe; ++q) {
            const int c = pci[q];
            int l = b;
            for (; l < e; ++l) if (aci[l] == c) break;
            const double a = av[l];
            pv[q] = a > 0 ? __ddiv_rn(__dmul_rn(-beta, a), aii) : __ddiv_rn(__dmul_rn(-alpha, a), aii);
        }
    } else if (mark[i] == 1) {                                     // CGPT
        pv[pb] = 1.0;
    }
}

__global__ void __launch_bounds__(TB) renumber_kernel(int nnz, const int *__restrict__ cindex, int *pci) {
    const int k = blockIdx.x * TB + threadIdx.x;
    if (k < nnz) pci[k] = cindex[pci[k]];
}

// SSS_inter.cu:31-86, one row: thresholds, number of kept entries, rescaling factors
__global__ void __launch_bounds__(TB) trunc_count_kernel(int n, double eps, const int *__restrict__ prp, const double *__restrict__ pv,
                                                         int *kept, double4 *fac) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const int b = prp[i], e = prp[i + 1];
    double lo = 0.0, hi = 0.0, sneg = 0.0, spos = 0.0, tneg = 0.0, tpos = 0.0;
    for (int k = b; k < e; ++k) {
        const double v = pv[k];
        if (v > 0) { spos = __dadd_rn(spos, v); if (v > hi) hi = v; }
        else if (v < 0) { sneg = __dadd_rn(sneg, v); if (v < lo) lo = v; }
    }
    hi = __dmul_rn(hi, eps); lo = __dmul_rn(lo, eps);
    int cnt = 0;
    for (int k = b; k < e; ++k) {
        const double v = pv[k];
        if (v >= hi) { ++cnt; tpos = __dadd_rn(tpos, v); }
        else if (v <= lo) { ++cnt; tneg = __dadd_rn(tneg, v); }
    }
    kept[i] = cnt;
    fac[i] = make_double4(hi, lo, tpos > TINY ? __ddiv_rn(spos, tpos) : 1.0, tneg < -TINY ? __ddiv_rn(sneg, tneg) : 1.0);
}

__global__ void __launch_bounds__(TB) trunc_write_kernel(int n, const int *__restrict__ prp, const int *__restrict__ pci, const double *__restrict__ pv,
                                                         const int *__restrict__ nrp, const double4 *__restrict__ fac, int *oci, double *ov) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const double4 f = fac[i];
    int w = nrp[i];
    for (int k = prp[i]; k < prp[i + 1]; ++k) {
        const double v = pv[k];
        if (v >= f.x) { oci[w] = pci[k]; ov[w] = __dmul_rn(v, f.z); ++w; }
        else if (v <= f.y) { oci[w] = pci[k]; ov[w] = __dmul_rn(v, f.w); ++w; }
    }
}

template <class T>
T *up(const T *h, size_t n) {
    T *d = nullptr;
    DEV_OK(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) DEV_OK(cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

}  // namespace

// A: level matrix; mark[n]: 0 F / 1 C / 2 isolated; P: pattern from the coarsening (row_ptr, col_idx = FINE indices of the interpolatory
// C points, val allocated).  On return P holds the truncated interpolation with coarse column indices, exactly as interp_DIR leaves it.
// Returns 0, or 1 when a row has no stored diagonal (nothing is modified: the caller falls back to the host loop).
extern "C" __attribute__((visibility("default"))) int amgb200_interp_device(const amgb200_mat *A, const int *mark, amgb200_mat *P, double trunc_threshold) {
    const int n = A->num_rows, pnnz = P->num_nnzs;
    const size_t annz = (size_t)A->row_ptr[n];
    int *d_arp = up(A->row_ptr, (size_t)n + 1), *d_aci = up(A->col_idx, annz), *d_mark = up(mark, (size_t)n);
    double *d_av = up(A->val, annz);
    int *d_prp = up(P->row_ptr, (size_t)n + 1), *d_pci = up(P->col_idx, (size_t)pnnz);
    double *d_pv = nullptr;
    DEV_OK(cudaMalloc(&d_pv, std::max<size_t>(pnnz, 1) * sizeof(double)));
    DEV_OK(cudaMemset(d_pv, 0, std::max<size_t>(pnnz, 1) * sizeof(double)));
    int *d_flag = nullptr, nodiag = 0;
    DEV_OK(cudaMalloc(&d_flag, sizeof(int)));
    DEV_OK(cudaMemset(d_flag, 0, sizeof(int)));
    const int grid = std::max(1, (n + TB - 1) / TB);
    interp_weights_kernel<<<grid, TB>>>(n, d_arp, d_aci, d_av, d_mark, d_prp, d_pci, d_pv, d_flag);
    DEV_OK(cudaGetLastError());
    DEV_OK(cudaMemcpy(&nodiag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_arp); cudaFree(d_aci); cudaFree(d_av); cudaFree(d_flag);
    if (nodiag) { cudaFree(d_mark); cudaFree(d_prp); cudaFree(d_pci); cudaFree(d_pv); return 1; }
    // coarse numbering of the C points (SSS_inter.cu:513-523)
    std::vector<int> cindex((size_t)n, 0);
    int nc = 0;
    for (int i = 0; i < n; ++i) if (mark[i] == 1) cindex[i] = nc++;
    P->num_cols = nc;
    int *d_cindex = up(cindex.data(), (size_t)n);
    if (pnnz) renumber_kernel<<<(pnnz + TB - 1) / TB, TB>>>(pnnz, d_cindex, d_pci);
    // truncation (SSS_inter.cu:16-102)
    int *d_kept = nullptr;
    double4 *d_fac = nullptr;
    DEV_OK(cudaMalloc(&d_kept, (size_t)n * sizeof(int)));
    DEV_OK(cudaMalloc(&d_fac, (size_t)n * sizeof(double4)));
    trunc_count_kernel<<<grid, TB>>>(n, trunc_threshold, d_prp, d_pv, d_kept, d_fac);
    DEV_OK(cudaGetLastError());
    std::vector<int> kept((size_t)n);
    DEV_OK(cudaMemcpy(kept.data(), d_kept, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    long long tot = 0;
    for (int i = 0; i < n; ++i) { P->row_ptr[i] = (int)tot; tot += kept[i]; }
    P->row_ptr[n] = (int)tot;
    int *d_nrp = up(P->row_ptr, (size_t)n + 1), *d_oci = nullptr;
    double *d_ov = nullptr;
    DEV_OK(cudaMalloc(&d_oci, std::max<size_t>((size_t)tot, 1) * sizeof(int)));
    DEV_OK(cudaMalloc(&d_ov, std::max<size_t>((size_t)tot, 1) * sizeof(double)));
    trunc_write_kernel<<<grid, TB>>>(n, d_prp, d_pci, d_pv, d_nrp, d_fac, d_oci, d_ov);
    DEV_OK(cudaGetLastError());
    P->num_nnzs = (int)tot;
    if (tot > 0) {
        P->col_idx = (int *)realloc(P->col_idx, (size_t)tot * sizeof(int));
        P->val = (double *)realloc(P->val, (size_t)tot * sizeof(double));
        DEV_OK(cudaMemcpy(P->col_idx, d_oci, (size_t)tot * sizeof(int), cudaMemcpyDeviceToHost));
        DEV_OK(cudaMemcpy(P->val, d_ov, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost));
    }
    cudaFree(d_mark); cudaFree(d_prp); cudaFree(d_pci); cudaFree(d_pv); cudaFree(d_cindex); cudaFree(d_kept); cudaFree(d_fac);
    cudaFree(d_nrp); cudaFree(d_oci); cudaFree(d_ov);
    return 0;
}
