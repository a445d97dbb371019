// Host-only introspection hooks used by the CPU test-suite (tests/test_schedule.py) to validate
// the analysis without a GPU: the wavefront schedule and a layout walk that executes the
// Gauss-Seidel rows wavefront by wavefront in REVERSE intra-wavefront order.  If the schedule is
// right (rows of one wavefront are independent) the result is bit-identical to the reference's
// sequential sweep.  Not a compute path of the product: nothing in hier.cu calls these.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/amg_b200.h"
#include "analysis.h"

using namespace amgb200;

#pragma GCC visibility push(default)
extern "C" {

// order[n], wf_row_ptr[cap]; returns the number of wavefronts (F + C), counts[0..1] = per pass,
// counts[2] = pattern_symmetric, counts[3] = rows without diagonal; -1 if cap is too small
int amgb200_debug_schedule(const amgb200_mat *A, const int *mark, int *order, int *wf_row_ptr, int cap, int counts[4]) {
    Schedule S;
    build_schedule(*A, mark, S);
    const int W = S.wf_count[0] + S.wf_count[1];
    if (W + 1 > cap) return -1;
    memcpy(order, S.order.data(), (size_t)S.n * sizeof(int));
    memcpy(wf_row_ptr, S.wf_row_ptr.data(), (size_t)(W + 1) * sizeof(int));
    counts[0] = S.wf_count[0]; counts[1] = S.wf_count[1];
    counts[2] = S.pattern_symmetric; counts[3] = S.rows_without_diag;
    return W;
}

// x, b in natural numbering; kind 0 = SELL walk (sequential per row), 1 = CSR walk
void amgb200_debug_gs_walk(const amgb200_mat *A, const int *mark, int kind, int nsweeps, double *x, const double *b) {
    Schedule S;
    build_schedule(*A, mark, S);
    DevLayout L;
    build_layout(*A, S.order.data(), S.pos.data(), kind >= 2 ? (int)KIND_CSR : kind, &S.wf_row_ptr, L);
    const int n = S.n, W = S.wf_count[0] + S.wf_count[1];
    std::vector<double> xs(n), bs(n);
    for (int k = 0; k < n; ++k) { xs[k] = x[S.order[k]]; bs[k] = b[S.order[k]]; }
    if (kind == 3) {
        // the cluster streaming smoother's algorithm on its packed blocks (C = 4 CTAs), with the WEAKEST visibility the
        // kernel guarantees: the product pass of wavefront g sees global x only through wavefront g-3; entries at
        // distance 1 and 2 come from the exchange buffers of the producing CTA
        const int Cc = 4;
        ClusterStreamLayout SL;
        build_stream_cluster(*A, S, Cc, SL);
        const int NB = SL.late_dist + 1;                               // global x is visible to the product pass through wavefront g - NB
        const int total = W * nsweeps;
        std::vector<double> vis(xs);                                   // global x as visible to the product pass
        std::vector<std::vector<std::pair<int, double>>> pending((size_t)total);
        std::vector<std::vector<std::vector<double>>> exb((size_t)total);   // [step][cta][local index]
        for (int g = 0; g < total; ++g) {
            if (g >= NB) for (auto &pr : pending[g - NB]) vis[pr.first] = pr.second;
            const int w = g % W;
            exb[g].assign(Cc, {});
            for (int c = Cc - 1; c >= 0; --c) {
                std::vector<unsigned char> blk(SL.data.data() + (size_t)SL.blk_ptr[(size_t)w * Cc + c] * 16,
                                               SL.data.data() + (size_t)SL.blk_ptr[(size_t)w * Cc + c + 1] * 16);
                const int *hd = reinterpret_cast<const int *>(blk.data());
                exb[g][c].assign(hd[0], 0.0);
                // product pass and prefix folds of every row, then ALL late products of the block from its flat list, then the
                // suffix folds
                std::vector<double> tacc(hd[0], 0.0);
                for (int li = hd[0] - 1; li >= 0; --li) {
                    unsigned char *rec = blk.data() + hd[8 + li];
                    const int *rh = reinterpret_cast<const int *>(rec);
                    double *val = reinterpret_cast<double *>(rec + 32);
                    const int *col = reinterpret_cast<const int *>(val + rh[2]);
                    if (rh[0] != hd[1] + c + li * Cc) { fprintf(stderr, "cluster stream: row dealing broken\n"); abort(); }
                    for (int p = 0; p < rh[2]; ++p) if (col[p] >= 0) val[p] = val[p] * vis[col[p]];
                    double t = bs[rh[0]];
                    for (int p = 0; p < rh[1]; ++p) t -= val[p];
                    tacc[li] = t;
                }
                const unsigned char *lt = blk.data() + hd[5];
                for (int i = 0; i < hd[4]; ++i) {
                    StreamLateC e;
                    memcpy(&e, lt + (size_t)i * 24, 24);
                    const int d = (e.src & 3) + 1, ii = e.src >> 2;
                    const double xv = g - d >= 0 ? exb[g - d][ii % Cc][ii / Cc] : vis[e.col];
                    *reinterpret_cast<double *>(blk.data() + e.pos) = e.val * xv;
                }
                for (int li = hd[0] - 1; li >= 0; --li) {
                    unsigned char *rec = blk.data() + hd[8 + li];
                    const int *rh = reinterpret_cast<const int *>(rec);
                    const double *val = reinterpret_cast<const double *>(rec + 32);
                    double t = tacc[li];
                    for (int p = rh[1]; p < rh[2]; ++p) t -= val[p];
                    const double d = reinterpret_cast<double *>(rec)[2];
                    double xn = xs[rh[0]];
                    if (fabs(d) > 1e-20) xn = t / d;
                    xs[rh[0]] = xn;
                    exb[g][c][li] = xn;
                    pending[g].emplace_back(rh[0], xn);
                }
            }
            if (g >= NB) exb[g - NB].clear();
        }
        for (int k = 0; k < n; ++k) x[S.order[k]] = xs[k];
        return;
    }
    if (kind == 2) {
        // the streaming smoother's algorithm on its packed blocks (kernels.cuh, gs_stream_cta_kernel) with three wavefronts
        // in flight, in the least favourable legal interleaving: the product pass of wavefront g+2 and the late2 patch +
        // prefix fold of wavefront g+1 run entirely BEFORE the post-barrier half of wavefront g (W >= 4; otherwise two in flight)
        StreamLayout SL;
        build_stream(L, SL);
        const int total = W * nsweeps;
        const int D = W >= 4 ? 3 : 2;
        std::vector<std::vector<unsigned char>> live(3);
        auto prod = [&](int g) {
            const int w = g % W;
            std::vector<unsigned char> &blk = live[g % 3];
            blk.assign(SL.data.data() + (size_t)SL.blk_ptr[w] * 16, SL.data.data() + (size_t)SL.blk_ptr[w + 1] * 16);
            const int *hd = reinterpret_cast<const int *>(blk.data());
            for (int ri = 0; ri < hd[0]; ++ri) {
                unsigned char *rec = blk.data() + hd[4 + ri];
                const int *rh = reinterpret_cast<const int *>(rec);
                double *val = reinterpret_cast<double *>(rec + 32);
                const int *col = reinterpret_cast<const int *>(val + rh[2]);
                for (int p = 0; p < rh[2]; ++p) if (col[p] >= 0) val[p] = val[p] * xs[col[p]];
            }
        };
        auto mid = [&](int g) {
            std::vector<unsigned char> &blk = live[g % 3];
            const int *hd = reinterpret_cast<const int *>(blk.data());
            for (int ri = 0; ri < hd[0]; ++ri) {
                unsigned char *rec = blk.data() + hd[4 + ri];
                const int *rh = reinterpret_cast<const int *>(rec);
                double *val = reinterpret_cast<double *>(rec + 32);
                const StreamLate *lt2 = reinterpret_cast<const StreamLate *>(reinterpret_cast<const int *>(val + rh[2]) + rh[2]) + (rh[3] & 0xffff);
                for (int i = 0; i < (rh[3] >> 16); ++i) val[lt2[i].pos] = lt2[i].val * xs[lt2[i].col];
                double t = bs[rh[0]];
                for (int p = 0; p < rh[1]; ++p) t -= val[p];
                reinterpret_cast<double *>(rec)[3] = t;
            }
        };
        auto post = [&](int g) {
            std::vector<unsigned char> &blk = live[g % 3];
            const int *hd = reinterpret_cast<const int *>(blk.data());
            for (int ri = hd[0] - 1; ri >= 0; --ri) {
                unsigned char *rec = blk.data() + hd[4 + ri];
                const int *rh = reinterpret_cast<const int *>(rec);
                double *val = reinterpret_cast<double *>(rec + 32);
                const StreamLate *lt = reinterpret_cast<const StreamLate *>(reinterpret_cast<const int *>(val + rh[2]) + rh[2]);
                for (int i = 0; i < (rh[3] & 0xffff); ++i) val[lt[i].pos] = lt[i].val * xs[lt[i].col];
                double t = reinterpret_cast<double *>(rec)[3];
                for (int p = rh[1]; p < rh[2]; ++p) t -= val[p];
                const double d = reinterpret_cast<double *>(rec)[2];
                if (fabs(d) > 1e-20) xs[rh[0]] = t / d;
            }
        };
        if (D == 3) {
            if (total > 0) prod(0);
            if (total > 1) prod(1);
            if (total > 0) mid(0);
            for (int g = 0; g < total; ++g) {
                if (g + 2 < total) prod(g + 2);      // needs done(g-1)
                if (g + 1 < total) mid(g + 1);       // needs done(g-1)
                post(g);
            }
        } else {
            if (total > 0) { prod(0); mid(0); }
            for (int g = 0; g < total; ++g) {
                if (g + 1 < total) { prod(g + 1); mid(g + 1); }
                post(g);
            }
        }
        for (int k = 0; k < n; ++k) x[S.order[k]] = xs[k];
        return;
    }
    for (int s = 0; s < nsweeps; ++s)
        for (int w = 0; w < W; ++w)
            for (int it = L.wf_item_ptr[w + 1] - 1; it >= L.wf_item_ptr[w]; --it) {     // reverse order inside the wavefront
                if (kind == KIND_SELL) {
                    const long long p0 = L.slice_ptr[it];
                    const int width = (int)((L.slice_ptr[it + 1] - p0) / 32);
                    for (int k = L.slice_row[it + 1] - 1; k >= L.slice_row[it]; --k) {
                        const int lane = k - L.slice_row[it];
                        double t = bs[k], d = 0.0;
                        for (int e = 0; e < width; ++e) {
                            const int j = L.col[(size_t)(p0 + 32LL * e + lane)];
                            const double a = L.val[(size_t)(p0 + 32LL * e + lane)];
                            if (j == k) d = a;
                            else if (j >= 0) t -= a * xs[j];
                        }
                        if (fabs(d) > 1e-20) xs[k] = t / d;
                    }
                } else {
                    const int k = it;
                    double t = bs[k], d = 0.0;
                    for (int p = L.rptr[k]; p < L.rptr[k + 1]; ++p) {
                        if (L.col[p] == k) d = L.val[p];
                        else t -= L.val[p] * xs[L.col[p]];
                    }
                    if (fabs(d) > 1e-20) xs[k] = t / d;
                }
            }
    for (int k = 0; k < n; ++k) x[S.order[k]] = xs[k];
}

// y = M x through the device layout (rows in row_order, columns through col_pos), natural in/out
void amgb200_debug_spmv_walk(const amgb200_mat *M, int kind, const double *x, double *y) {
    DevLayout L;
    build_layout(*M, nullptr, nullptr, kind, nullptr, L);
    if (kind == KIND_SELL) {
        const int ns = (int)L.slice_row.size() - 1;
        for (int s = 0; s < ns; ++s) {
            const long long p0 = L.slice_ptr[s];
            const int width = (int)((L.slice_ptr[s + 1] - p0) / 32);
            for (int k = L.slice_row[s]; k < L.slice_row[s + 1]; ++k) {
                const int lane = k - L.slice_row[s];
                double t = 0.0;
                for (int e = 0; e < width; ++e) {
                    const int j = L.col[(size_t)(p0 + 32LL * e + lane)];
                    if (j >= 0) t += L.val[(size_t)(p0 + 32LL * e + lane)] * x[j];
                }
                y[k] = t;
            }
        }
    } else {
        for (int k = 0; k < L.nrows; ++k) {
            double t = 0.0;
            for (int p = L.rptr[k]; p < L.rptr[k + 1]; ++p) t += L.val[p] * x[L.col[p]];
            y[k] = t;
        }
    }
}

// The fused residual (+) restriction launch (kernels.cuh, resid_restrict_kernel) executed on the CPU through the device layouts and
// the ticket list: tickets in list order (the legal execution of a single warp), r poisoned with NaN beforehand.  Checks the
// invariants the kernel relies on -- every slice exactly once, a slice of R only after every slice of A in the chunks it declares,
// every residual row it reads inside those chunks.  x, b, r (n_l) and bc (n_{l+1}) in natural numbering.  Ac / markC describe level
// l+1 (markC == NULL: its schedule is the identity, as on the coarsest level).  Returns 0, or a negative code naming the invariant.
int amgb200_debug_fused_walk(const amgb200_mat *A, const int *markA, const amgb200_mat *Ac, const int *markC, int coarsest, const amgb200_mat *R,
                             int nch, int lag, int tickets, const double *x, const double *b, double *r, double *bc) {
    Schedule Sf, Sc;
    build_schedule(*A, markA, Sf);
    if (coarsest) identity_schedule(Ac->num_rows, Sc); else build_schedule(*Ac, markC, Sc);
    DevLayout la, lr;
    build_layout(*A, Sf.order.data(), Sf.pos.data(), KIND_SELL, nullptr, la);
    build_layout(*R, Sc.order.data(), Sf.pos.data(), KIND_SELL, nullptr, lr);
    FusedPlan F;
    build_fused_plan(la, lr, Sf, Sc, *R, nch, lag, tickets, F);
    const int nA = la.nitems(), nR = lr.nitems(), n = Sf.n, nc = Sc.n;
    if (F.work.size() != F.block_info.size() * (size_t)F.tickets) return -1;
    std::vector<double> xs(n), bs(n), rs(n, std::nan("")), bcs(nc, std::nan(""));
    for (int k = 0; k < n; ++k) { xs[k] = x[Sf.order[k]]; bs[k] = b[Sf.order[k]]; }
    std::vector<unsigned> cnt((size_t)F.nch, 0u);
    std::vector<char> seenA(nA, 0), seenR(nR, 0);
    std::vector<int> slice_of(n);
    for (int s = 0; s < nA; ++s) for (int k = la.slice_row[s]; k < la.slice_row[s + 1]; ++k) slice_of[k] = s;
    auto row_sum = [](const DevLayout &L, int s, int k, const std::vector<double> &v, auto &&visit) {
        const long long p0 = L.slice_ptr[s];
        const int width = (int)((L.slice_ptr[s + 1] - p0) / 32), lane = k - L.slice_row[s];
        double t = 0.0;
        for (int e = 0; e < width; ++e) {
            const int j = L.col[(size_t)(p0 + 32LL * e + lane)];
            if (j >= 0) { visit(j); t += L.val[(size_t)(p0 + 32LL * e + lane)] * v[j]; }
        }
        return t;
    };
    for (size_t blk = 0; blk < F.block_info.size(); ++blk) {
        const unsigned info = (unsigned)F.block_info[blk];
        const int *wk = F.work.data() + blk * (size_t)F.tickets;
        if (!(info >> 31)) {
            const unsigned c = info & 0xffffu;
            unsigned done = 0;
            for (int q = 0; q < F.tickets; ++q) {
                const int w = wk[q];
                if (w == FUSED_NOP) continue;
                if (w < 0 || w >= nA || seenA[w] || (unsigned)F.slice_chunk[w] != c) return -2;
                seenA[w] = 1; ++done;
                for (int k = la.slice_row[w]; k < la.slice_row[w + 1]; ++k) rs[k] = bs[k] + row_sum(la, w, k, xs, [](int) {}) * -1.0;
            }
            if (done != info >> 16) return -6;
            cnt[c] += done;
        } else {
            const unsigned lo = info & 0xffffu, hi = (info >> 16) & 0x7fffu;
            for (unsigned c = lo; c <= hi; ++c)
                if (cnt[c] != F.chunk_items[c]) return -4;                     // the kernel would wait here for a LARGER block: deadlock
            for (int q = 0; q < F.tickets; ++q) {
                const int w = wk[q];
                if (w == FUSED_NOP) continue;
                const int t = ~w;
                if (w >= 0 || t >= nR || seenR[t]) return -3;
                seenR[t] = 1;
                int bad = 0;
                for (int k = lr.slice_row[t]; k < lr.slice_row[t + 1]; ++k)
                    bcs[k] = row_sum(lr, t, k, rs, [&](int j) {                 // every row of the 128-byte line of r_j must be final (lines stay in L1)
                        for (int p = j & ~15; p < std::min(n, (j & ~15) + 16); ++p) { const unsigned c = (unsigned)F.slice_chunk[slice_of[p]]; if (c < lo || c > hi) bad = 1; }
                    });
                if (bad) return -5;
            }
        }
    }
    for (int s = 0; s < nA; ++s) if (!seenA[s]) return -7;
    for (int t = 0; t < nR; ++t) if (!seenR[t]) return -7;
    for (int k = 0; k < n; ++k) r[Sf.order[k]] = rs[k];
    for (int k = 0; k < nc; ++k) bc[Sc.order[k]] = bcs[k];
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
