// Host-side analysis: wavefront (level-set) schedule of the reference's ordered Gauss-Seidel
// and construction of the permuted device layouts.  See DESIGN.md "Data layout in HBM".
//
// The reference's smoother (amg/Solve/SSS_smooth.c:16-49) is a true Gauss-Seidel in ascending
// row order, F rows (cfmark != 1) first, then C rows.  Row i of a pass needs the *new* x_j of
// every same-pass neighbour j < i and the *old* x_j of every same-pass neighbour j > i.  The
// wavefront number of a row is therefore 1 + max over same-pass neighbours j < i, where
// "neighbour" is taken in the symmetrised pattern (a_ij or a_ji stored) so that the
// read-old-value anti-dependencies are honoured for non-symmetric patterns too.  Rows of one
// wavefront are mutually independent; executing wavefronts in order with any intra-wavefront
// order reproduces the sequential iterates exactly.
#include "analysis.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <omp.h>

namespace amgb200 {

int choose_kind(const amgb200_mat &M, double sell_max_mean) {
    const double mean = M.num_rows ? (double)M.num_nnzs / M.num_rows : 0.0;
    return mean <= sell_max_mean ? KIND_SELL : KIND_CSR;
}

void identity_schedule(int n, Schedule &S) {
    S = Schedule();
    S.n = n;
    S.order.resize(n);
    S.pos.resize(n);
    for (int i = 0; i < n; ++i) S.order[i] = S.pos[i] = i;
    S.pass_rows[0] = n;
    S.wf_count[0] = n ? 1 : 0;
    S.wf_row_ptr = {0, n};
}

void build_schedule(const amgb200_mat &A, const int *mark, Schedule &S) {
    const int n = A.num_rows;
    S = Schedule();
    S.n = n;
    std::vector<int> lvl(n, 0), pend(n, 0);
    std::vector<unsigned char> pass(n, 0);
    if (mark) for (int i = 0; i < n; ++i) pass[i] = mark[i] == 1 ? 1 : 0;
    // the two passes are independent recurrences over disjoint rows (only same-pass neighbours enter): one thread each
    int depth[2] = {0, 0}, nodiag[2] = {0, 0}, rows[2] = {0, 0};
    bool sym[2] = {true, true};
#pragma omp parallel for schedule(static, 1) num_threads(2)
    for (int p = 0; p < 2; ++p) {
        int dmax = 0, nd = 0, nr = 0;
        bool sy = true;
        for (int i = 0; i < n; ++i) {
            if (pass[i] != p) continue;
            const int b = A.row_ptr[i], e = A.row_ptr[i + 1];
            int L = 1;
            bool diag = false;
            for (int k = b; k < e; ++k) {
                const int j = A.col_idx[k];
                if (j == i) { diag = true; continue; }
                if (j < i && pass[j] == p && lvl[j] + 1 > L) L = lvl[j] + 1;
            }
            if (pend[i] > L) { L = pend[i]; sy = false; }
            lvl[i] = L;
            if (!diag) ++nd;
            for (int k = b; k < e; ++k) {
                const int j = A.col_idx[k];
                if (j > i && pass[j] == p && pend[j] < L + 1) pend[j] = L + 1;
            }
            if (L > dmax) dmax = L;
            ++nr;
        }
        depth[p] = dmax; nodiag[p] = nd; rows[p] = nr; sym[p] = sy;
    }
    S.rows_without_diag = nodiag[0] + nodiag[1];
    S.pass_rows[0] = rows[0]; S.pass_rows[1] = rows[1];
    S.pattern_symmetric = sym[0] && sym[1];
    S.wf_count[0] = depth[0];
    S.wf_count[1] = depth[1];
    const int W = depth[0] + depth[1];
    S.wf_row_ptr.assign(W + 1, 0);
    for (int i = 0; i < n; ++i) S.wf_row_ptr[(pass[i] ? depth[0] : 0) + lvl[i] - 1 + 1]++;
    for (int w = 0; w < W; ++w) S.wf_row_ptr[w + 1] += S.wf_row_ptr[w];
    std::vector<int> fill(S.wf_row_ptr.begin(), S.wf_row_ptr.end() - 1);
    S.order.resize(n);
    S.pos.resize(n);
    for (int i = 0; i < n; ++i) {
        const int k = fill[(pass[i] ? depth[0] : 0) + lvl[i] - 1]++;
        S.order[k] = i;
        S.pos[i] = k;
    }
}

void build_layout(const amgb200_mat &M, const int *row_order, const int *col_pos, int kind,
                  const std::vector<int> *breaks, DevLayout &L) {
    L.~DevLayout();
    new (&L) DevLayout();
    L.kind = kind;
    const int n = M.num_rows;
    L.nrows = n;
    L.ncols = M.num_cols;
    L.nnz = M.row_ptr[n];
    auto nat = [&](int k) { return row_order ? row_order[k] : k; };
    int maxlen = 0;
#pragma omp parallel for reduction(max : maxlen) schedule(static)
    for (int i = 0; i < n; ++i) maxlen = std::max(maxlen, M.row_ptr[i + 1] - M.row_ptr[i]);
    L.max_row = maxlen;

    if (kind == KIND_CSR) {
        L.rptr.resize((size_t)n + 1);
        L.rptr[0] = 0;
        for (int k = 0; k < n; ++k) { const int i = nat(k); L.rptr[k + 1] = L.rptr[k] + (M.row_ptr[i + 1] - M.row_ptr[i]); }
        L.col.resize((size_t)L.nnz);
        L.val.resize((size_t)L.nnz);
#pragma omp parallel for schedule(static)
        for (int k = 0; k < n; ++k) {
            const int i = nat(k);
            int w = L.rptr[k];
            for (int q = M.row_ptr[i]; q < M.row_ptr[i + 1]; ++q, ++w) {
                const int j = M.col_idx[q];
                L.col[w] = col_pos ? col_pos[j] : j;
                L.val[w] = M.val[q];
            }
        }
        if (breaks) {
            L.wf_item_ptr = *breaks;
            // wavefront of every schedule row, then the prefix/suffix split of every row
            const int W = (int)breaks->size() - 1;
            std::vector<int> wf_of((size_t)n);
            for (int w = 0; w < W; ++w) for (int k = (*breaks)[w]; k < (*breaks)[w + 1]; ++k) wf_of[k] = w;
            L.split.resize((size_t)n);
            L.late.assign(((size_t)L.nnz + 31) / 32 + 1, 0u);
            L.late2.assign(W >= 4 ? ((size_t)L.nnz + 31) / 32 + 1 : 0, 0u);
#pragma omp parallel for schedule(static)
            for (int k = 0; k < n; ++k) {              // (bit writes of neighbouring rows share words: atomic OR)
                const int prev = wf_of[k] == 0 ? W - 1 : wf_of[k] - 1;
                const int prev2 = W >= 4 ? (wf_of[k] + W - 2) % W : -1;
                int sp = L.rptr[k + 1] - L.rptr[k];
                for (int q = L.rptr[k]; q < L.rptr[k + 1]; ++q) {
                    const int c = L.col[q];
                    if (c == k || c >= n) continue;
                    if (wf_of[c] == prev) {
                        if (q - L.rptr[k] < sp) sp = q - L.rptr[k];
                        __atomic_fetch_or(&L.late[(size_t)q >> 5], 1u << (q & 31), __ATOMIC_RELAXED);
                    } else if (wf_of[c] == prev2) __atomic_fetch_or(&L.late2[(size_t)q >> 5], 1u << (q & 31), __ATOMIC_RELAXED);
                }
                L.split[k] = sp;
            }
        }
        return;
    }

    // SELL-32
    build_sell_structure(M, row_order, breaks, L);
    const int ns = (int)L.slice_row.size() - 1;
    const size_t total = (size_t)L.slice_ptr[ns];
    L.col.resize(total);          // every slot is written below (entries or padding): no separate fill pass
    L.val.resize(total);
#pragma omp parallel for schedule(static)
    for (int s = 0; s < ns; ++s) {
        const long long base = L.slice_ptr[s];
        const int nr = L.slice_row[s + 1] - L.slice_row[s];
        const int width = (int)((L.slice_ptr[s + 1] - base) / 32);
        for (int lane = 0; lane < 32; ++lane) {
            long long w = base + lane;
            int filled = 0;
            if (lane < nr) {
                const int i = nat(L.slice_row[s] + lane);
                for (int q = M.row_ptr[i]; q < M.row_ptr[i + 1]; ++q, w += 32, ++filled) {
                    const int j = M.col_idx[q];
                    L.col[(size_t)w] = col_pos ? col_pos[j] : j;
                    L.val[(size_t)w] = M.val[q];
                }
            }
            for (; filled < width; ++filled, w += 32) { L.col[(size_t)w] = -1; L.val[(size_t)w] = 0.0; }
        }
    }
}

// The O(rows) part of a SELL-32 layout: slices (never straddling a `breaks` boundary), their widths and offsets.  col/val stay
// empty: hier.cu fills them on the device from the raw CSR arrays (sell_fill_kernel), the host only when asked (build_layout).
void build_sell_structure(const amgb200_mat &M, const int *row_order, const std::vector<int> *breaks, DevLayout &L) {
    L.~DevLayout();
    new (&L) DevLayout();
    L.kind = KIND_SELL;
    const int n = M.num_rows;
    L.nrows = n;
    L.ncols = M.num_cols;
    L.nnz = M.row_ptr[n];
    auto nat = [&](int k) { return row_order ? row_order[k] : k; };
    int maxlen = 0;
#pragma omp parallel for reduction(max : maxlen) schedule(static)
    for (int i = 0; i < n; ++i) maxlen = std::max(maxlen, M.row_ptr[i + 1] - M.row_ptr[i]);
    L.max_row = maxlen;
    std::vector<int> seg;
    if (breaks) seg = *breaks; else seg = {0, n};
    L.slice_row.clear();
    if (breaks) L.wf_item_ptr.clear();
    for (size_t s = 0; s + 1 < seg.size(); ++s) {
        if (breaks) L.wf_item_ptr.push_back((int)L.slice_row.size());
        for (int r = seg[s]; r < seg[s + 1]; r += 32) L.slice_row.push_back(r);
    }
    if (breaks) L.wf_item_ptr.push_back((int)L.slice_row.size());
    L.slice_row.push_back(n);
    const int ns = (int)L.slice_row.size() - 1;
    L.slice_ptr.assign((size_t)ns + 1, 0);
    std::vector<int> width((size_t)ns);
#pragma omp parallel for schedule(static)
    for (int s = 0; s < ns; ++s) {
        int w = 0;
        for (int k = L.slice_row[s]; k < L.slice_row[s + 1]; ++k) { const int i = nat(k); w = std::max(w, M.row_ptr[i + 1] - M.row_ptr[i]); }
        width[s] = w;
    }
    for (int s = 0; s < ns; ++s) L.slice_ptr[s + 1] = L.slice_ptr[s] + 32LL * width[s];
}

void build_stream(const DevLayout &L, StreamLayout &S) {
    const int W = (int)L.wf_item_ptr.size() - 1;
    const std::vector<int> &wip = L.wf_item_ptr;
    auto r8 = [](int v) { return (v + 7) & ~7; };
    auto r16 = [](int v) { return (v + 15) & ~15; };
    auto late_bit = [&](int q) { return (L.late[(size_t)q >> 5] >> (q & 31)) & 1u; };
    const bool have2 = !L.late2.empty();
    auto late2_bit = [&](int q) { return have2 ? (L.late2[(size_t)q >> 5] >> (q & 31)) & 1u : 0u; };
    // Bank alignment: the folding warp's lanes read 16 bytes each from DIFFERENT rows in lockstep (same term index in every
    // row slot).  Blocks start on 128-byte boundaries of shared memory, the product array of row i of a wavefront starts in
    // 16-byte bank group i % 8, and the prefix is padded to a multiple of 16 terms (128 bytes) so that the suffix starts in
    // the same group: 32 row slots then hit every bank group exactly 4 times (no conflicts beyond the 512 bytes moved).
    // per-row record size (before alignment padding)
    std::vector<int> rec_bytes((size_t)L.nrows);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < L.nrows; ++k) {
        const int p0 = L.rptr[k], len = L.rptr[k + 1] - p0, sp = L.split[k];
        int nlate = 0, nlate2 = 0;
        for (int q = p0 + sp; q < p0 + len; ++q) nlate += late_bit(q);
        for (int q = p0; q < p0 + len; ++q) nlate2 += late2_bit(q);
        if (nlate > 0xffff || nlate2 > 0x7fff) { fprintf(stderr, "libamgb200: build_stream: late list of a row exceeds the record format\n"); abort(); }
        const int len_pad = r16(sp) + r8(len - sp);
        rec_bytes[k] = 32 + len_pad * 12 + (nlate + nlate2) * 16;
    }
    auto place = [](int off, int ri) {           // first offset >= off whose product array (off + 32) lies in bank group ri % 8
        while ((((off + 32) >> 4) & 7) != (ri & 7)) off += 16;
        return off;
    };
    S.blk_ptr.assign((size_t)W + 1, 0);
    S.max_block = 0;
    for (int w = 0; w < W; ++w) {
        const int nr = wip[w + 1] - wip[w];
        int off = 16 + ((nr + 3) & ~3) * 4;
        for (int k = wip[w]; k < wip[w + 1]; ++k) off = place(off, k - wip[w]) + rec_bytes[k];
        const long long bytes = (off + 127) & ~127;
        const int i0a = wip[w] & ~1, bcnt = (wip[w + 1] - i0a + 1) & ~1;
        S.blk_ptr[w + 1] = S.blk_ptr[w] + (int)(bytes / 16);
        S.max_block = (int)std::max<long long>(S.max_block, ((bytes + (long long)bcnt * 8) + 127) & ~127LL);
    }
    S.data.resize((size_t)S.blk_ptr[W] * 16);
    S.mean_block = W ? (long long)S.blk_ptr[W] * 16 / W : 0;
#pragma omp parallel for schedule(dynamic, 8)
    for (int w = 0; w < W; ++w) {
        unsigned char *blk = S.data.data() + (size_t)S.blk_ptr[w] * 16;
        const int blk_bytes = (S.blk_ptr[w + 1] - S.blk_ptr[w]) * 16;
        memset(blk, 0, (size_t)blk_bytes);
        const int nr = wip[w + 1] - wip[w];
        int *hd = reinterpret_cast<int *>(blk);
        const int i0a = wip[w] & ~1;
        hd[0] = nr; hd[1] = i0a; hd[2] = (wip[w + 1] - i0a + 1) & ~1; hd[3] = blk_bytes;
        int *rec_off = hd + 4;
        int off = 16 + ((nr + 3) & ~3) * 4;
        for (int k = wip[w]; k < wip[w + 1]; ++k) {
            off = place(off, k - wip[w]);
            rec_off[k - wip[w]] = off;
            unsigned char *rec = blk + off;
            const int p0 = L.rptr[k], len = L.rptr[k + 1] - p0, sp = L.split[k];
            const int pre_pad = r16(sp), len_pad = pre_pad + r8(len - sp);
            int *rh = reinterpret_cast<int *>(rec);
            double *rd = reinterpret_cast<double *>(rec);
            double *val = rd + 4;
            int *col = reinterpret_cast<int *>(val + len_pad);
            StreamLate *lt = reinterpret_cast<StreamLate *>(col + len_pad);
            for (int i = 0; i < len_pad; ++i) { val[i] = 0.0; col[i] = -1; }
            double diag = 0.0;
            int nlate = 0, nlate2 = 0;
            for (int i = 0; i < len; ++i) {
                const int q = p0 + i, pos = i < sp ? i : pre_pad + (i - sp);
                const int c = L.col[q];
                if (c == k) { diag = L.val[q]; continue; }          // (SSS_smooth.c:29-30: the diagonal is not part of the sum)
                val[pos] = L.val[q]; col[pos] = c;
                if (late_bit(q)) { lt[nlate].val = L.val[q]; lt[nlate].pos = pos; lt[nlate].col = c; ++nlate; }
            }
            StreamLate *lt2 = lt + nlate;
            for (int i = 0; i < len; ++i) {
                const int q = p0 + i, pos = i < sp ? i : pre_pad + (i - sp);
                if (late2_bit(q)) { lt2[nlate2].val = L.val[q]; lt2[nlate2].pos = pos; lt2[nlate2].col = L.col[q]; ++nlate2; }
            }
            rh[0] = k; rh[1] = pre_pad; rh[2] = len_pad; rh[3] = nlate | (nlate2 << 16);
            rd[2] = diag; rd[3] = 0.0;
            off += rec_bytes[k];
        }
    }
}

void build_stream_cluster(const amgb200_mat &A, const Schedule &S, int C, ClusterStreamLayout &SL, long long smem_budget, int late_dist) {
    const int n = S.n, W = S.wf_count[0] + S.wf_count[1];
    const std::vector<int> &wrp = S.wf_row_ptr;
    auto r8 = [](int v) { return (v + 7) & ~7; };
    std::vector<int> wf_of((size_t)n);
    for (int w = 0; w < W; ++w) for (int k = wrp[w]; k < wrp[w + 1]; ++k) wf_of[k] = w;
    // cyclic distance from the wavefront of column c back to wavefront w (1 = the wavefront just before)
    auto dist = [&](int w, int c) { const int d = w - wf_of[c]; return d > 0 ? d : d + W; };
    const int LD = (late_dist >= 3 && W >= 6) ? 3 : 2;       // entries at cyclic wavefront distance <= LD come from the exchange buffers
    SL.C = C;
    SL.late_dist = LD;
    SL.max_width = 0;
    for (int w = 0; w < W; ++w) SL.max_width = std::max(SL.max_width, wrp[w + 1] - wrp[w]);
    // the three exchange buffers alone (every CTA holds the whole of the last three wavefronts) must leave room for a ring
    if (smem_budget >= 0 && 8LL * (LD + 1) * ((SL.max_width + 1) & ~1) + 8192 > smem_budget) { SL.filled = false; return; }
    std::vector<int> rec_bytes((size_t)n), split((size_t)n), nlate((size_t)n);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; ++k) {
        const int i = S.order[k];
        const int p0 = A.row_ptr[i], len = A.row_ptr[i + 1] - p0;
        int sp = len, nl = 0;
        for (int q = 0; q < len; ++q) {
            const int c = S.pos[A.col_idx[p0 + q]];
            if (c != k && dist(wf_of[k], c) <= LD) { if (q < sp) sp = q; ++nl; }
        }
        split[k] = sp; nlate[k] = nl;
        rec_bytes[k] = 32 + (r8(sp) + r8(len - sp)) * 12;
        rec_bytes[k] = (rec_bytes[k] + 15) & ~15;
    }
    SL.blk_ptr.assign((size_t)W * C + 1, 0);
    SL.max_block = 0; SL.max_local = 0; SL.max_width = 0;
    for (int w = 0; w < W; ++w) {
        const int width = wrp[w + 1] - wrp[w];
        for (int c = 0; c < C; ++c) {
            const int nr = width > c ? (width - c + C - 1) / C : 0;
            long long bytes = 32 + (long long)((nr + 3) & ~3) * 4;
            long long nflat = 0;
            for (int li = 0; li < nr; ++li) { bytes += rec_bytes[wrp[w] + c + li * C]; nflat += nlate[wrp[w] + c + li * C]; }
            bytes += ((nflat * 24 + 15) & ~15LL);
            SL.blk_ptr[(size_t)w * C + c + 1] = SL.blk_ptr[(size_t)w * C + c] + (int)(bytes / 16);
            SL.max_block = (int)std::max<long long>(SL.max_block, bytes);
            SL.max_local = std::max(SL.max_local, nr);
            SL.max_width = std::max(SL.max_width, width);
        }
    }
    SL.mean_block = W ? (long long)SL.blk_ptr[(size_t)W * C] * 16 / ((long long)W * C) : 0;
    // sizes are known: give up before the (expensive) fill when two of the largest blocks plus the three exchange
    // buffers cannot fit in one CTA's shared memory
    SL.filled = !(smem_budget >= 0 && 2LL * SL.max_block > smem_budget - 8LL * (LD + 1) * ((SL.max_width + 1) & ~1));
    if (!SL.filled) return;
    SL.data.resize((size_t)SL.blk_ptr[(size_t)W * C] * 16);
#pragma omp parallel for schedule(dynamic, 8)
    for (int w = 0; w < W; ++w) {
        const int width = wrp[w + 1] - wrp[w];
        for (int c = 0; c < C; ++c) {
            unsigned char *blk = SL.data.data() + (size_t)SL.blk_ptr[(size_t)w * C + c] * 16;
            const int nr = width > c ? (width - c + C - 1) / C : 0;
            int *hd = reinterpret_cast<int *>(blk);
            const int blk_bytes = (SL.blk_ptr[(size_t)w * C + c + 1] - SL.blk_ptr[(size_t)w * C + c]) * 16;
            hd[0] = nr; hd[1] = wrp[w]; hd[2] = width; hd[3] = blk_bytes;
            int *rec_off = hd + 8;
            int off = 32 + ((nr + 3) & ~3) * 4;
            for (int li = nr; li < ((nr + 3) & ~3); ++li) rec_off[li] = 0;
            int flat_off = off;
            for (int li = 0; li < nr; ++li) flat_off += rec_bytes[wrp[w] + c + li * C];
            unsigned char *flat = blk + flat_off;
            memset(flat, 0, (size_t)(blk_bytes - flat_off));
            int nflat = 0;
            for (int li = 0; li < nr; ++li) {
                const int k = wrp[w] + c + li * C, i = S.order[k];
                rec_off[li] = off;
                unsigned char *rec = blk + off;
                const int p0 = A.row_ptr[i], len = A.row_ptr[i + 1] - p0, sp = split[k];
                const int pre_pad = r8(sp), len_pad = pre_pad + r8(len - sp);
                int *rh = reinterpret_cast<int *>(rec);
                double *rd = reinterpret_cast<double *>(rec);
                double *val = rd + 4;
                int *col = reinterpret_cast<int *>(val + len_pad);
                memset(rec, 0, (size_t)rec_bytes[k]);
                for (int q = 0; q < len_pad; ++q) col[q] = -1;
                double diag = 0.0;
                int nl = 0;
                for (int q = 0; q < len; ++q) {
                    const int pos = q < sp ? q : pre_pad + (q - sp);
                    const int cc = S.pos[A.col_idx[p0 + q]];
                    const double v = A.val[p0 + q];
                    if (cc == k) { diag = v; continue; }
                    val[pos] = v; col[pos] = cc;
                    const int d = dist(w, cc);
                    if (d <= LD) {
                        const int ii = cc - wrp[wf_of[cc]];
                        StreamLateC e;
                        e.val = v; e.pos = off + 32 + 8 * pos; e.col = cc; e.src = (d - 1) | (ii << 2); e.pad = 0;
                        memcpy(flat + (size_t)nflat * 24, &e, 24);
                        ++nflat; ++nl;
                    }
                }
                rh[0] = k; rh[1] = pre_pad; rh[2] = len_pad; rh[3] = nl;
                rd[2] = diag; rd[3] = 0.0;
                off += rec_bytes[k];
            }
            hd[4] = nflat; hd[5] = flat_off; hd[6] = 0; hd[7] = 0;
        }
    }
}

void build_fused_plan(const DevLayout &la, const DevLayout &lr, const Schedule &Sf, const Schedule &Sc, const amgb200_mat &Rm, int nch, int lag, int tickets, FusedPlan &F) {
    const int nA = la.nitems(), nR = lr.nitems(), nf = Sf.n;
    nch = std::max(1, std::min(nch, 32767));
    tickets = std::max(1, tickets);
    F.nch = nch; F.tickets = tickets;
    F.slice_chunk.assign((size_t)nA, 0);
    // per 128-byte line of r (16 consecutive schedule positions): first and last chunk that write into it.  A slice of R waits for
    // every chunk that owns a part of a line it touches, so a line is final when an SM first loads it and may stay in its
    // (non-coherent) L1 for the rest of the launch.
    const int nlines = (nf + 15) / 16;
    std::vector<unsigned short> pos_chunk((size_t)nf);
#pragma omp parallel for schedule(static)
    for (int s = 0; s < nA; ++s) {
        const int c = (int)std::min<long long>(nch - 1, (long long)Sf.order[la.slice_row[s]] * nch / std::max(1, nf));
        F.slice_chunk[s] = c;
        for (int p = la.slice_row[s]; p < la.slice_row[s + 1]; ++p) pos_chunk[p] = (unsigned short)c;
    }
    std::vector<unsigned> line_need((size_t)nlines);
#pragma omp parallel for schedule(static)
    for (int ln = 0; ln < nlines; ++ln) {
        unsigned lo = 0xffffu, hi = 0;
        for (int p = ln * 16; p < std::min(nf, ln * 16 + 16); ++p) { lo = std::min<unsigned>(lo, pos_chunk[p]); hi = std::max<unsigned>(hi, pos_chunk[p]); }
        line_need[ln] = lo | hi << 16;
    }
    F.rneed.assign((size_t)nR, 0u);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < nR; ++t) {
        unsigned lo = 0xffffu, hi = 0;
        for (int p = lr.slice_row[t]; p < lr.slice_row[t + 1]; ++p) {
            const int j = Sc.order[p];
            for (int q = Rm.row_ptr[j]; q < Rm.row_ptr[j + 1]; ++q) {
                const unsigned nd = line_need[Sf.pos[Rm.col_idx[q]] >> 4];
                lo = std::min(lo, nd & 0xffffu); hi = std::max(hi, nd >> 16);
            }
        }
        if (lo > hi) lo = hi = 0;                                           // (slice of empty rows)
        F.rneed[t] = lo | hi << 16;
    }
    F.chunk_items.assign((size_t)nch, 0u);
    for (int s = 0; s < nA; ++s) F.chunk_items[F.slice_chunk[s]]++;
    std::vector<int> a_ptr((size_t)nch + 1, 0), r_ptr((size_t)nch + 2, 0);
    for (int c = 0; c < nch; ++c) a_ptr[c + 1] = a_ptr[c] + (int)F.chunk_items[c];
    auto r_key = [&](int t) { const unsigned lo = F.rneed[t] & 0xffffu, hi = F.rneed[t] >> 16; return hi - lo >= 32u ? nch : (int)hi; };
    for (int t = 0; t < nR; ++t) r_ptr[r_key(t) + 1]++;
    for (int c = 0; c <= nch; ++c) r_ptr[c + 1] += r_ptr[c];
    std::vector<int> a_sorted((size_t)nA), r_sorted((size_t)nR);
    {
        std::vector<int> fa(a_ptr.begin(), a_ptr.end() - 1), fr(r_ptr.begin(), r_ptr.end() - 1);
        for (int s = 0; s < nA; ++s) a_sorted[fa[F.slice_chunk[s]]++] = s;
        for (int t = 0; t < nR; ++t) r_sorted[fr[r_key(t)]++] = t;
    }
    F.work.clear();
    F.block_info.clear();
    F.work.reserve((size_t)nA + nR + (size_t)tickets * (2 * nch + 2));
    auto emit_a = [&](int c) {
        for (int q0 = a_ptr[c]; q0 < a_ptr[c + 1]; q0 += tickets) {
            const int cnt = std::min(tickets, a_ptr[c + 1] - q0);
            for (int q = 0; q < tickets; ++q) F.work.push_back(q < cnt ? a_sorted[q0 + q] : FUSED_NOP);
            F.block_info.push_back(c | cnt << 16);
        }
    };
    auto emit_r = [&](int key) {
        for (int q0 = r_ptr[key]; q0 < r_ptr[key + 1]; q0 += tickets) {
            const int cnt = std::min(tickets, r_ptr[key + 1] - q0);
            unsigned lo = 0xffffu, hi = 0;
            for (int q = 0; q < tickets; ++q) {
                F.work.push_back(q < cnt ? ~r_sorted[q0 + q] : FUSED_NOP);
                if (q < cnt) { lo = std::min(lo, F.rneed[r_sorted[q0 + q]] & 0xffffu); hi = std::max(hi, F.rneed[r_sorted[q0 + q]] >> 16); }
            }
            if (key == nch) { lo = 0; hi = (unsigned)nch - 1; }
            F.block_info.push_back((int)(0x80000000u | lo | hi << 16));
        }
    };
    for (int c = 0; c < nch + lag; ++c) {
        if (c < nch) emit_a(c);
        if (c - lag >= 0 && c - lag < nch) emit_r(c - lag);
    }
    emit_r(nch);
}

}  // namespace amgb200
