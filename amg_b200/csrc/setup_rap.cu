// Device versions of the two setup steps that produce the operators of the next level:
//   R = P^T                (/root/reference/amg/SSS_matvec.c:330-387, SSS_mat_trans)
//   A_{l+1} = R A P        (/root/reference/amg/SSS_matvec.c:398-534, SSS_blas_mat_rap)
// Both outputs fix the SUMMATION ORDER of the solve phase (restriction rows, Gauss-Seidel and residual rows of the coarse level), so
// they have to be the reference's arrays entry for entry, not just the same matrices:
//   * a row of R lists its entries by ascending fine row (the reference scatters P row by row);
//   * a row of R A P starts with the diagonal slot (value 0.0 + contributions), the other columns follow in DISCOVERY order of the
//     triple loop  for a in R(ic,:) / for b in A(i1,:) / for c in P(i2,:) ; a column's value is its first product (r*a)*p, later
//     products are added in traversal order, every product rounded as (r*a) first, then *p (no FMA).
// Rows are independent.  A group of lanes owns one coarse row and walks (a, b) sequentially; the entries of ONE row of P have distinct
// columns, so they are handled by the lanes in parallel: look-up in a per-row hash table in shared memory, new columns get their
// slots in lane order (ballot + rank), then every lane adds into its own slot.  Two passes (count, prefix sum on the host, fill);
// rows are binned by length into three table sizes, the largest keeps values in the output array itself (L2).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/amg_b200.h"

namespace {

#define DEV_OK(call)                                                                                          \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            fprintf(stderr, "libamgb200 (device setup): CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(70);                                                                                         \
        }                                                                                                     \
    } while (0)

constexpr int TB = 256;
constexpr unsigned FULLMASK = 0xffffffffu;

struct Csr { const int *rp; const int *ci; const double *v; };

// ---------------------------------------------------------------------------------------------------------------------------------
// transpose
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TB) col_count_kernel(int nnz, const int *__restrict__ ci, int *cnt) {
    const int k = blockIdx.x * TB + threadIdx.x;
    if (k < nnz) atomicAdd(cnt + ci[k], 1);
}
// entries land in their row of the transpose in arbitrary order ...
__global__ void __launch_bounds__(TB) transpose_scatter_kernel(int n, Csr P, int *cursor, int *tci, double *tv) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    for (int k = P.rp[i]; k < P.rp[i + 1]; ++k) {
        const int w = atomicAdd(cursor + P.ci[k], 1);
        tci[w] = i; tv[w] = P.v[k];
    }
}
// ... and are then put into ascending fine-row order, which is the order of the reference's row-by-row scatter (a row of P has
// distinct columns, so the keys of one row of the transpose are distinct).  One thread per row: binary insertion is enough for the
// rows of a restriction (mean 7, a few hundred on the coarsest levels).
__global__ void __launch_bounds__(TB) row_sort_kernel(int nrows, const int *__restrict__ rp, int *ci, double *v) {
    const int j = blockIdx.x * TB + threadIdx.x;
    if (j >= nrows) return;
    const int b = rp[j], e = rp[j + 1];
    for (int k = b + 1; k < e; ++k) {
        const int key = ci[k];
        const double val = v[k];
        int q = k - 1;
        while (q >= b && ci[q] > key) { ci[q + 1] = ci[q]; v[q + 1] = v[q]; --q; }
        ci[q + 1] = key; v[q + 1] = val;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Galerkin product
// ---------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned hash_of(int key, unsigned mask) { return ((unsigned)key * 2654435761u) >> 7 & mask; }

// LPR lanes per coarse row, GROUPS rows per CTA; HASH = table size (power of two), CAP = HASH / 2 = longest row this instance can hold.
// FILL = false: len[row] = number of columns (or -1: table too small).  FILL = true: writes the row into cci / cv at crp[row].
// VALS_GLOBAL: the values are accumulated in cv itself (L2) instead of shared memory (largest table only).
// rows == nullptr: rows 0 .. nrows-1.
template <int LPR, int GROUPS, int HASH, bool FILL, bool VALS_GLOBAL>
__global__ void __launch_bounds__(LPR *GROUPS) rap_kernel(int nrows, const int *__restrict__ rows, Csr R, Csr A, Csr P, int *len, const int *__restrict__ crp, int *cci,
                                                          double *cv) {
    constexpr int CAP = HASH / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / LPR, lg = threadIdx.x % LPR;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = LPR == 32 ? FULLMASK : ((1u << LPR) - 1u) << (lane / LPR * LPR);
    // per group: keys[HASH], slots[HASH] (FILL), vals[CAP] + cols[CAP] (FILL && !VALS_GLOBAL)
    constexpr size_t per_group = (size_t)HASH * 4 + (FILL ? (size_t)HASH * 4 : 0) + ((FILL && !VALS_GLOBAL) ? (size_t)CAP * 12 : 0);
    unsigned char *base = smem_raw + (size_t)g * per_group;
    int *keys = reinterpret_cast<int *>(base);
    int *slots = FILL ? keys + HASH : nullptr;
    double *vals = (FILL && !VALS_GLOBAL) ? reinterpret_cast<double *>(base + (size_t)HASH * 8) : nullptr;
    int *cols = (FILL && !VALS_GLOBAL) ? reinterpret_cast<int *>(base + (size_t)HASH * 8 + (size_t)CAP * 8) : nullptr;
    const unsigned mask = HASH - 1;

    for (int rq = blockIdx.x * GROUPS + g; rq < nrows; rq += gridDim.x * GROUPS) {
        const int ic = rows ? rows[rq] : rq;
        for (int h = lg; h < HASH; h += LPR) keys[h] = -1;
        __syncwarp(gmask);
        const long long row0 = FILL ? crp[ic] : 0;
        // the diagonal slot comes first (SSS_matvec.c:446-449 / :489-493), value 0.0
        int cnt = 1;
        if (lg == 0) {
            const unsigned h = hash_of(ic, mask);
            keys[h] = ic;
            if (FILL) {
                slots[h] = 0;
                if (VALS_GLOBAL) { cci[row0] = ic; __stcg(cv + row0, 0.0); } else { cols[0] = ic; vals[0] = 0.0; }
            }
        }
        __syncwarp(gmask);
        bool overflow = false;
        for (int a = R.rp[ic]; a < R.rp[ic + 1] && !overflow; ++a) {
            const double r = FILL ? R.v[a] : 0.0;
            const int i1 = R.ci[a];
            for (int bq = A.rp[i1]; bq < A.rp[i1 + 1] && !overflow; ++bq) {
                const double ra = FILL ? __dmul_rn(r, A.v[bq]) : 0.0;
                const int i2 = A.ci[bq];
                const int pb = P.rp[i2], pe = P.rp[i2 + 1];
                for (int c0 = pb; c0 < pe; c0 += LPR) {
                    const int c = c0 + lg;
                    const bool act = c < pe;
                    int i3 = -1, slot = -1;
                    double prod = 0.0;
                    bool isnew = false;
                    if (act) {
                        i3 = P.ci[c];
                        if (FILL) prod = __dmul_rn(ra, P.v[c]);
                        unsigned h = hash_of(i3, mask);
                        for (;;) {
                            const int k = keys[h];
                            if (k == i3) { if (FILL) slot = slots[h]; break; }
                            if (k == -1) { isnew = true; break; }
                            h = (h + 1) & mask;
                        }
                    }
                    // new columns take their slots in lane order = the order of the reference's inner loop
                    const unsigned newm = __ballot_sync(gmask, isnew) & gmask;
                    if (newm) {
                        const int nnew = __popc(newm);
                        if (cnt + nnew > CAP) { overflow = true; break; }
                        if (isnew) {
                            slot = cnt + __popc(newm & ((1u << lane) - 1u));
                            unsigned h = hash_of(i3, mask);
                            for (;;) {
                                const int old = atomicCAS(keys + h, -1, i3);
                                if (old == -1) { if (FILL) slots[h] = slot; break; }
                                h = (h + 1) & mask;
                            }
                        }
                        cnt += nnew;
                    }
                    if (FILL && act) {
                        if (VALS_GLOBAL) {
                            if (isnew) { cci[row0 + slot] = i3; __stcg(cv + row0 + slot, prod); }
                            else __stcg(cv + row0 + slot, __dadd_rn(__ldcg(cv + row0 + slot), prod));
                        } else {
                            if (isnew) { cols[slot] = i3; vals[slot] = prod; }
                            else vals[slot] = __dadd_rn(vals[slot], prod);
                        }
                    }
                    __syncwarp(gmask);
                }
            }
        }
        __syncwarp(gmask);
        if (!FILL) {
            if (lg == 0) len[ic] = overflow ? -1 : cnt;
        } else if (!VALS_GLOBAL) {
            for (int q = lg; q < cnt; q += LPR) { cci[row0 + q] = cols[q]; cv[row0 + q] = vals[q]; }
        }
        __syncwarp(gmask);
    }
}

template <class T>
T *up(const T *h, size_t n) {
    T *d = nullptr;
    DEV_OK(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) DEV_OK(cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}
template <class T>
T *dalloc(size_t n) {
    T *d = nullptr;
    DEV_OK(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    return d;
}
template <class T>
T *host_calloc(size_t n) {
    T *p = (T *)calloc(std::max<size_t>(n, 1), sizeof(T));      // (the reference allocates with calloc, SSS_utils.c; SSS_amg_data_destroy frees with free)
    if (!p) { fprintf(stderr, "libamgb200 (device setup): out of host memory\n"); exit(-15); }
    return p;
}

template <class T>
T *host_alloc(size_t n) {                                         // (completely overwritten by the copy that follows)
    T *p = (T *)malloc(std::max<size_t>(n, 1) * sizeof(T));
    if (!p) { fprintf(stderr, "libamgb200 (device setup): out of host memory\n"); exit(-15); }
    // first touch in parallel: the page faults of a fresh gigabyte-sized block otherwise serialise inside the D2H copy
    const size_t bytes = n * sizeof(T);
    if (bytes >= (8u << 20)) {
        char *c = reinterpret_cast<char *>(p);
#pragma omp parallel for schedule(static)
        for (long long off = 0; off < (long long)bytes; off += 4096) c[off] = 0;
    }
    return p;
}

// the three table sizes: rows of <= 64 / <= 1024 / <= 8192 columns
constexpr int T1_LPR = 8, T1_GROUPS = 32, T1_HASH = 128;
constexpr int T2_LPR = 32, T2_GROUPS = 4, T2_HASH = 2048;
constexpr int T3_LPR = 32, T3_GROUPS = 1, T3_HASH = 16384;

template <int LPR, int GROUPS, int HASH, bool FILL, bool VG>
void launch_rap(int nrows, const int *d_rows, const Csr &R, const Csr &A, const Csr &P, int *d_len, const int *d_crp, int *d_cci, double *d_cv) {
    if (nrows <= 0) return;
    constexpr size_t per_group = (size_t)HASH * 4 + (FILL ? (size_t)HASH * 4 : 0) + ((FILL && !VG) ? (size_t)(HASH / 2) * 12 : 0);
    const size_t smem = per_group * GROUPS;
    auto kern = rap_kernel<LPR, GROUPS, HASH, FILL, VG>;
    DEV_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, per_sm = 0;
    DEV_OK(cudaGetDevice(&dev));
    DEV_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DEV_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LPR * GROUPS, smem));
    const int grid = std::max(1, std::min((nrows + GROUPS - 1) / GROUPS, std::max(1, per_sm) * sms * 4));
    kern<<<grid, LPR * GROUPS, smem>>>(nrows, d_rows, R, A, P, d_len, d_crp, d_cci, d_cv);
    DEV_OK(cudaGetLastError());
}

}  // namespace

// R = P^T and Ac = R A P on the device, returned as host CSR arrays (calloc'ed, owned by the caller) exactly as SSS_mat_trans and
// SSS_blas_mat_rap produce them.  Returns 0; 1 = a row of the product has more than 8192 columns or the product exceeds 2^31 entries
// (nothing is returned: the caller falls back to the host loops).
extern "C" __attribute__((visibility("default"))) int amgb200_rap_device(const amgb200_mat *A, const amgb200_mat *P, amgb200_mat *R_out, amgb200_mat *Ac_out) {
    const int nf = A->num_rows, nc = P->num_cols;
    const size_t annz = (size_t)A->row_ptr[nf], pnnz = (size_t)P->row_ptr[nf];
    if (nc <= 0) return 1;
    const bool timing = getenv("AMGB200_SETUP_TIMING") && atoi(getenv("AMGB200_SETUP_TIMING"));
    auto now = [&]() { if (timing) cudaDeviceSynchronize(); return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_kern = 0;
    // ---- transpose ----------------------------------------------------------------------------
    int *d_prp = up(P->row_ptr, (size_t)nf + 1), *d_pci = up(P->col_idx, pnnz);
    double *d_pv = up(P->val, pnnz);
    int *d_cnt = dalloc<int>((size_t)nc);
    DEV_OK(cudaMemset(d_cnt, 0, (size_t)nc * sizeof(int)));
    double tk = now();
    if (pnnz) col_count_kernel<<<(int)((pnnz + TB - 1) / TB), TB>>>((int)pnnz, d_pci, d_cnt);
    DEV_OK(cudaGetLastError());
    t_kern += now() - tk;
    std::vector<int> cnt((size_t)nc);
    DEV_OK(cudaMemcpy(cnt.data(), d_cnt, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost));
    int *rrp = host_calloc<int>((size_t)nc + 1);
    for (int j = 0; j < nc; ++j) rrp[j + 1] = rrp[j] + cnt[j];
    int *d_rrp = up(rrp, (size_t)nc + 1);
    DEV_OK(cudaMemcpy(d_cnt, rrp, (size_t)nc * sizeof(int), cudaMemcpyHostToDevice));          // cursors
    int *d_rci = dalloc<int>(pnnz);
    double *d_rv = dalloc<double>(pnnz);
    const Csr Pd{d_prp, d_pci, d_pv};
    tk = now();
    transpose_scatter_kernel<<<std::max(1, (nf + TB - 1) / TB), TB>>>(nf, Pd, d_cnt, d_rci, d_rv);
    row_sort_kernel<<<std::max(1, (nc + TB - 1) / TB), TB>>>(nc, d_rrp, d_rci, d_rv);
    DEV_OK(cudaGetLastError());
    t_kern += now() - tk;
    // ---- Galerkin product: count -----------------------------------------------------------------
    int *d_arp = up(A->row_ptr, (size_t)nf + 1), *d_aci = up(A->col_idx, annz);
    double *d_av = up(A->val, annz);
    const Csr Rd{d_rrp, d_rci, d_rv}, Ad{d_arp, d_aci, d_av};
    int *d_len = dalloc<int>((size_t)nc);
    tk = now();
    launch_rap<T1_LPR, T1_GROUPS, T1_HASH, false, false>(nc, nullptr, Rd, Ad, Pd, d_len, nullptr, nullptr, nullptr);
    t_kern += now() - tk;
    std::vector<int> len((size_t)nc);
    DEV_OK(cudaMemcpy(len.data(), d_len, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> redo;
    for (int j = 0; j < nc; ++j) if (len[j] < 0) redo.push_back(j);
    int *d_rows = nullptr;
    bool failed = false;
    if (!redo.empty()) {
        d_rows = up(redo.data(), redo.size());
        launch_rap<T2_LPR, T2_GROUPS, T2_HASH, false, false>((int)redo.size(), d_rows, Rd, Ad, Pd, d_len, nullptr, nullptr, nullptr);
        DEV_OK(cudaMemcpy(len.data(), d_len, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<int> redo2;
        for (int j : redo) if (len[j] < 0) redo2.push_back(j);
        if (!redo2.empty()) {
            cudaFree(d_rows);
            d_rows = up(redo2.data(), redo2.size());
            launch_rap<T3_LPR, T3_GROUPS, T3_HASH, false, false>((int)redo2.size(), d_rows, Rd, Ad, Pd, d_len, nullptr, nullptr, nullptr);
            DEV_OK(cudaMemcpy(len.data(), d_len, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost));
            for (int j : redo2) if (len[j] < 0) failed = true;
        }
        cudaFree(d_rows); d_rows = nullptr;
    }
    long long tot = 0;
    int *crp = nullptr;
    if (!failed) {
        crp = host_calloc<int>((size_t)nc + 1);
        for (int j = 0; j < nc; ++j) { crp[j] = (int)tot; tot += len[j]; if (tot > 2147483647LL) { failed = true; break; } }
        if (!failed) crp[nc] = (int)tot;
    }
    if (failed) {
        free(crp); free(rrp);
        cudaFree(d_prp); cudaFree(d_pci); cudaFree(d_pv); cudaFree(d_cnt); cudaFree(d_rrp); cudaFree(d_rci); cudaFree(d_rv);
        cudaFree(d_arp); cudaFree(d_aci); cudaFree(d_av); cudaFree(d_len);
        return 1;
    }
    // ---- fill, rows binned by length -----------------------------------------------------------------
    std::vector<int> bin[3];
    for (int j = 0; j < nc; ++j) bin[len[j] <= T1_HASH / 2 ? 0 : len[j] <= T2_HASH / 2 ? 1 : 2].push_back(j);
    int *d_crp = up(crp, (size_t)nc + 1);
    int *d_cci = dalloc<int>((size_t)tot);
    double *d_cv = dalloc<double>((size_t)tot);
    tk = now();
    for (int t = 0; t < 3; ++t) {
        if (bin[t].empty()) continue;
        const bool all = (int)bin[t].size() == nc;
        int *d_list = all ? nullptr : up(bin[t].data(), bin[t].size());
        const int m = (int)bin[t].size();
        if (t == 0) launch_rap<T1_LPR, T1_GROUPS, T1_HASH, true, false>(m, d_list, Rd, Ad, Pd, nullptr, d_crp, d_cci, d_cv);
        else if (t == 1) launch_rap<T2_LPR, T2_GROUPS, T2_HASH, true, false>(m, d_list, Rd, Ad, Pd, nullptr, d_crp, d_cci, d_cv);
        else launch_rap<T3_LPR, T3_GROUPS, T3_HASH, true, true>(m, d_list, Rd, Ad, Pd, nullptr, d_crp, d_cci, d_cv);
        DEV_OK(cudaDeviceSynchronize());
        if (d_list) cudaFree(d_list);
    }
    t_kern += now() - tk;
    // ---- results ------------------------------------------------------------------------------------
    R_out->num_rows = nc; R_out->num_cols = nf; R_out->num_nnzs = (int)pnnz;
    R_out->row_ptr = rrp;
    R_out->col_idx = host_alloc<int>(pnnz);
    R_out->val = host_alloc<double>(pnnz);
    DEV_OK(cudaMemcpy(R_out->col_idx, d_rci, pnnz * sizeof(int), cudaMemcpyDeviceToHost));
    DEV_OK(cudaMemcpy(R_out->val, d_rv, pnnz * sizeof(double), cudaMemcpyDeviceToHost));
    Ac_out->num_rows = Ac_out->num_cols = nc; Ac_out->num_nnzs = (int)tot;
    Ac_out->row_ptr = crp;
    Ac_out->col_idx = host_alloc<int>((size_t)tot);
    Ac_out->val = host_alloc<double>((size_t)tot);
    DEV_OK(cudaMemcpy(Ac_out->col_idx, d_cci, (size_t)tot * sizeof(int), cudaMemcpyDeviceToHost));
    DEV_OK(cudaMemcpy(Ac_out->val, d_cv, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_prp); cudaFree(d_pci); cudaFree(d_pv); cudaFree(d_cnt); cudaFree(d_rrp); cudaFree(d_rci); cudaFree(d_rv);
    cudaFree(d_arp); cudaFree(d_aci); cudaFree(d_av); cudaFree(d_len); cudaFree(d_crp); cudaFree(d_cci); cudaFree(d_cv);
    if (timing) printf("[rap_device] %d -> %d rows, product %lld entries (bins %zu / %zu / %zu rows): kernels %.1f ms (transpose + first count + fill), total %.1f ms\n",
                       nf, nc, tot, bin[0].size(), bin[1].size(), bin[2].size(), 1e3 * t_kern, 1e3 * (now() - t_begin));
    return 0;
}
