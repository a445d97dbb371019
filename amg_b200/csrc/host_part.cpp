// Host helper of the multi-GPU harness (amg_b200/distributed.py): the ghost lists of a row-block partition of level 0 in one
// OpenMP pass over the matrix.  Every entry (row, col) whose row and column are owned by different ranks makes the column a ghost
// of the row's owner.  No reference counterpart (the reference is single-device); the numpy version it replaces took 1.3 s at 256^3.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <omp.h>

#include "../../include/amg_b200.h"

namespace {
inline int owner_of(int k, int nF, int world, const long long *fb, const long long *cb) {
    const long long *b = k < nF ? fb : cb;              // first schedule row of every rank's block, world + 1 entries
    int lo = 0, hi = world;                             // last p with b[p] <= k
    while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (b[mid] <= k) lo = mid; else hi = mid; }
    return lo;
}
}  // namespace

// order[n]: schedule position -> natural row of level 0; f_bounds / c_bounds: world + 1 ascending schedule-row offsets of the F / C
// blocks.  Returns the concatenated lists in *idx (malloc'ed, caller frees with free()) and their offsets in ptr[world*world*2 + 1]:
// list key = (reader * world + owner) * 2 + pass of the ghost (0 F, 1 C); each list holds ascending, distinct schedule indices.
extern "C" __attribute__((visibility("default"))) long long amgb200_ghost_lists_ex(const amgb200_mat *A, const int *order, int nF, int world, const long long *f_bounds,
                                                                                     const long long *c_bounds, int **idx, long long *ptr, unsigned char *reads_ghost);
extern "C" __attribute__((visibility("default"))) long long amgb200_ghost_lists(const amgb200_mat *A, const int *order, int nF, int world, const long long *f_bounds,
                                                                                  const long long *c_bounds, int **idx, long long *ptr) {
    return amgb200_ghost_lists_ex(A, order, nF, world, f_bounds, c_bounds, idx, ptr, nullptr);
}
extern "C" __attribute__((visibility("default"))) long long amgb200_ghost_lists_ex(const amgb200_mat *A, const int *order, int nF, int world, const long long *f_bounds,
                                                                                     const long long *c_bounds, int **idx, long long *ptr, unsigned char *reads_ghost) {
    const int n = A->num_rows;
    if (reads_ghost) memset(reads_ghost, 0, (size_t)n);
    std::vector<int> pos((size_t)n);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; ++k) pos[order[k]] = k;
    std::vector<unsigned short> own((size_t)n);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; ++k) own[k] = (unsigned short)owner_of(k, nF, world, f_bounds, c_bounds);
    const int nt = omp_get_max_threads();
    std::vector<std::vector<unsigned long long>> found((size_t)nt);
#pragma omp parallel
    {
        std::vector<unsigned long long> &mine = found[omp_get_thread_num()];
#pragma omp for schedule(static)
        for (int i = 0; i < n; ++i) {
            const int reader = own[pos[i]];
            for (int q = A->row_ptr[i]; q < A->row_ptr[i + 1]; ++q) {
                const int cp = pos[A->col_idx[q]];
                const int src = own[cp];
                if (src != reader && reads_ghost) __atomic_store_n(&reads_ghost[pos[i]], (unsigned char)1, __ATOMIC_RELAXED);
                if (src != reader) mine.push_back((unsigned long long)(((long long)reader * world + src) * 2 + (cp >= nF)) << 32 | (unsigned)cp);
            }
        }
    }
    std::vector<unsigned long long> all;
    for (auto &v : found) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    const long long nkeys = (long long)world * world * 2;
    *idx = (int *)malloc(std::max<size_t>(all.size(), 1) * sizeof(int));
    if (!*idx) { fprintf(stderr, "amgb200_ghost_lists: out of memory\n"); exit(-15); }
    for (long long k = 0; k <= nkeys; ++k) ptr[k] = 0;
    for (size_t e = 0; e < all.size(); ++e) { (*idx)[e] = (int)(all[e] & 0xffffffffu); ptr[(all[e] >> 32) + 1]++; }
    for (long long k = 0; k < nkeys; ++k) ptr[k + 1] += ptr[k];
    return (long long)all.size();
}
