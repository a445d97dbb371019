// From-scratch host restatement of the reference's AMG *setup* phase, so that the product can
// build its input hierarchy without any reference object.  Pure host C++, no device code.
//
// The solve phase (the hot path this library accelerates) consumes a hierarchy whose C/F
// marks and CSR entry order determine the Gauss-Seidel iterates, so this file reproduces the
// reference's setup bit for bit, including its tie-breaking order:
//   level loop + stopping checks   amg/Setup/SSS_SETUP.cu:36-177
//   strength of connection         amg/Setup/SSS_coarsen.c:106-181
//   Ruge-Stueben C/F splitting     amg/Setup/SSS_coarsen.c:294-498 (bucket lists :22-105, :220-292)
//   F-F cleanup, P pattern         amg/Setup/SSS_coarsen.c:501-574, :577-630
//   direct interpolation + trunc   amg/Setup/SSS_inter.cu:400-547, :16-102
//   transpose, Galerkin product    amg/SSS_matvec.c:330-387, :398-534
// tests/test_setup_parity.py checks every array of every level against the reference's own
// setup (oracle/_ref) for byte equality.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <omp.h>

#include "../../include/amg_b200.h"

namespace {

enum { F_PT = 0, C_PT = 1, ISO_PT = 2, UNDECIDED = -1 };
constexpr int END_OF_LIST = -2, START_OF_LIST = -1;
constexpr double TINY = 1e-20;

template <class T>
T *zalloc(size_t n) {
    if (n == 0) return nullptr;
    T *p = (T *)calloc(n, sizeof(T));
    if (!p) { fprintf(stderr, "amgb200_setup: out of host memory (%zu bytes)\n", n * sizeof(T)); exit(-20); }
    return p;
}

struct Pattern {             // integer CSR pattern (strength matrix S and its transpose)
    int rows = 0, cols = 0;
    std::vector<int> ptr, idx;
};

// ---- strength of connection (SSS_coarsen.c:106-181): returns S with weak entries marked -1
void strength(const amgb200_mat &A, const amgb200_pars &pr, Pattern &S) {
    const int n = A.num_rows;
    S.rows = n; S.cols = A.num_cols;
    S.ptr.assign(A.row_ptr, A.row_ptr + n + 1);
    S.idx.assign(A.col_idx, A.col_idx + A.num_nnzs);
    const double dom = 2 - pr.max_row_sum;
    for (int i = 0; i < n; ++i) {
        const int b = A.row_ptr[i], e = A.row_ptr[i + 1];
        double dii = 0.0;
        for (int k = b; k < e; ++k) if (A.col_idx[k] == i) { dii = A.val[k]; break; }
        double scale = 0.0, asum = 0.0;
        for (int k = b; k < e; ++k) {
            const double a = fabs(A.val[k]);
            asum += a;
            if (A.col_idx[k] != i && a > scale) scale = a;
        }
        scale *= pr.strong_threshold;
        for (int k = b; k < e; ++k) if (A.col_idx[k] == i) { S.idx[k] = -1; break; }
        if (asum < dom * fabs(dii)) {
            for (int k = b; k < e; ++k) S.idx[k] = -1;
        } else {
            for (int k = b; k < e; ++k) if (-A.val[k] <= scale) S.idx[k] = -1;   // negative couplings only
        }
    }
}

// drop the -1 entries (SSS_coarsen.c:185-212); false if nothing is left
bool compress(Pattern &S) {
    int w = 0;
    for (int i = 0; i < S.rows; ++i) {
        const int b = S.ptr[i], e = S.ptr[i + 1];
        S.ptr[i] = w;
        for (int k = b; k < e; ++k) if (S.idx[k] > -1) S.idx[w++] = S.idx[k];
    }
    S.ptr[S.rows] = w;
    S.idx.resize(w);
    return w > 0;
}

void transpose_pattern(const Pattern &S, Pattern &T) {   // SSS_matvec.c:247-310 (pattern only)
    T.rows = S.cols; T.cols = S.rows;
    T.ptr.assign(T.rows + 1, 0);
    T.idx.resize(S.idx.size());
    for (int c : S.idx) T.ptr[c + 1]++;
    for (int i = 0; i < T.rows; ++i) T.ptr[i + 1] += T.ptr[i];
    std::vector<int> fill(T.ptr.begin(), T.ptr.end() - 1);
    for (int i = 0; i < S.rows; ++i)
        for (int k = S.ptr[i]; k < S.ptr[i + 1]; ++k) T.idx[fill[S.idx[k]]++] = i;
}

// Measure buckets of the RS splitting: one FIFO doubly-linked list per measure value, the
// candidate C point is the *head* of the largest non-empty measure (SSS_coarsen.c:22-105, :220-292).
// nxt/prv are the reference's `lists`/`where` arrays (zero-initialised, like its calloc).
struct Buckets {
    std::vector<int> head, tail;      // per measure; head == END_OF_LIST means "no such list"
    std::vector<int> nxt, prv;
    int top = 0;                      // largest measure with a list (0 = none)
    explicit Buckets(int n) : nxt(n, 0), prv(n, 0) {}
    bool has(int m) const { return m >= 0 && m < (int)head.size() && head[m] != END_OF_LIST; }
    void push(int m, int i) {
        if (m >= (int)head.size()) { head.resize(m + 1, END_OF_LIST); tail.resize(m + 1, START_OF_LIST); }
        if (head[m] == END_OF_LIST) {
            head[m] = tail[m] = i; nxt[i] = END_OF_LIST; prv[i] = START_OF_LIST;
            if (m > top) top = m;
        } else {
            const int t = tail[m];
            nxt[t] = i; prv[i] = t; nxt[i] = END_OF_LIST; tail[m] = i;
        }
    }
    void pop(int m, int i) {
        if (!has(m)) { printf("### ERROR: This list is empty! %s : %d\n", __FILE__, __LINE__); return; }
        if (head[m] == i && tail[m] == i) {
            head[m] = END_OF_LIST; tail[m] = START_OF_LIST;
            while (top > 0 && head[top] == END_OF_LIST) --top;
        } else if (head[m] == i) {
            head[m] = nxt[i]; prv[nxt[i]] = START_OF_LIST;
        } else if (tail[m] == i) {
            tail[m] = prv[i]; nxt[prv[i]] = END_OF_LIST;
        } else {
            nxt[prv[i]] = nxt[i]; prv[nxt[i]] = prv[i];
        }
    }
    bool empty() const { return top <= 0 || head.empty() || head[top] == END_OF_LIST; }
};

// classical Ruge-Stueben first pass + C1-criterion second pass (SSS_coarsen.c:294-498).
// Returns number of C points, <= 0 on failure.
int rs_split(Pattern &S, int *mark) {
    const int n = S.rows;
    if (!compress(S)) return -99;
    int ncoarse = 0;
    Pattern ST;
    transpose_pattern(S, ST);
    std::vector<int> lam(n);
    Buckets bk(n);
    int left = 0;
    for (int i = 0; i < n; ++i) lam[i] = ST.ptr[i + 1] - ST.ptr[i];
    for (int i = 0; i < n; ++i) {
        if (S.ptr[i + 1] == S.ptr[i]) { mark[i] = ISO_PT; lam[i] = 0; }
        else { mark[i] = UNDECIDED; ++left; }
    }
    auto bump_after_F = [&](int j) {          // j became F: its undecided strong neighbours gain weight
        for (int l = S.ptr[j]; l < S.ptr[j + 1]; ++l) {
            const int k = S.idx[l];
            if (mark[k] == UNDECIDED) { bk.pop(lam[k], k); bk.push(++lam[k], k); }
        }
    };
    for (int i = 0; i < n; ++i) {
        if (mark[i] == ISO_PT) continue;
        if (lam[i] > 0) { bk.push(lam[i], i); continue; }
        if (lam[i] < 0) printf("### WARNING: Negative lambda[%d]!\n", i);
        mark[i] = F_PT; --left;               // influences nobody: F
        for (int k = S.ptr[i]; k < S.ptr[i + 1]; ++k) {
            const int j = S.idx[k];
            if (mark[j] == ISO_PT) continue;
            if (j < i) {
                if (lam[j] > 0) bk.pop(lam[j], j);
                bk.push(++lam[j], j);
            } else {
                ++lam[j];
            }
        }
    }
    while (left > 0) {
        if (bk.empty()) break;                // (the reference would dereference NULL here)
        const int c = bk.head[bk.top];
        const int cm = lam[c];
        if (cm == 0) printf("### WARNING: Head of the list has measure 0!\n");
        mark[c] = C_PT; lam[c] = 0; --left;
        bk.pop(cm, c);
        ++ncoarse;
        for (int q = ST.ptr[c]; q < ST.ptr[c + 1]; ++q) {      // points c influences become F
            const int j = ST.idx[q];
            if (mark[j] != UNDECIDED) continue;
            mark[j] = F_PT;
            bk.pop(lam[j], j);
            --left;
            bump_after_F(j);
        }
        for (int q = S.ptr[c]; q < S.ptr[c + 1]; ++q) {        // points influencing c lose weight
            const int j = S.idx[q];
            if (mark[j] != UNDECIDED) continue;
            int m = lam[j];
            bk.pop(m, j);
            lam[j] = --m;
            if (m > 0) { bk.push(m, j); }
            else { mark[j] = F_PT; --left; bump_after_F(j); }
        }
    }
    // second pass: every strong F-F pair must share a C point (SSS_coarsen.c:441-483)
    std::vector<int> &tag = lam;
    std::fill(tag.begin(), tag.end(), -1);
    int keep = 0;
    for (int i = 0; i < n; ++i) {
        if (mark[i] != F_PT) continue;
        const int e = S.ptr[i + 1];
        for (int q = S.ptr[i]; q < e; ++q) { const int j = S.idx[q]; if (mark[j] == C_PT) tag[j] = i; }
        int promoted = 0;
        for (int q = S.ptr[i]; q < e; ++q) {
            const int j = S.idx[q];
            if (mark[j] != F_PT) continue;
            bool shares = false;
            for (int r = S.ptr[j]; r < S.ptr[j + 1]; ++r) if (tag[S.idx[r]] == i) { shares = true; break; }
            if (shares) continue;
            if (promoted == 0) { mark[j] = C_PT; ++ncoarse; tag[j] = i; keep = j; promoted = 1; }
            else { mark[i] = C_PT; mark[keep] = F_PT; break; }
        }
    }
    return ncoarse;
}

// remove F-F couplings without common C point (SSS_coarsen.c:501-574); returns #C
int clean_ff(const Pattern &S, int *mark, int ncoarse) {
    const int n = S.rows;
    std::vector<int> tag(n, -1);
    bool tentative = false;          // the reference's C_i_nonempty (deliberately outlives the row)
    int tilde = -1, tilde_row = -1;
    for (int i = 0; i < n; ++i) {
        if (mark[i] != F_PT) continue;
        for (int q = S.ptr[i]; q < S.ptr[i + 1]; ++q) {
            const int j = S.idx[q];
            tag[j] = (mark[j] == C_PT) ? i : -1;
        }
        if (tilde_row != i) tilde = -1;
        for (int q = S.ptr[i]; q < S.ptr[i + 1]; ++q) {
            const int j = S.idx[q];
            if (mark[j] != F_PT) continue;
            bool shares = false;
            for (int r = S.ptr[j]; r < S.ptr[j + 1]; ++r) if (tag[S.idx[r]] == i) { shares = true; break; }
            if (shares) continue;
            if (tentative) {
                mark[i] = C_PT; ++ncoarse;
                if (tilde > -1) { mark[tilde] = F_PT; --ncoarse; tilde = -1; }
                tentative = false;
                break;
            }
            mark[j] = C_PT; ++ncoarse;     // tentatively promote j and look at row i again
            tilde = j; tilde_row = i; tentative = true;
            --i;
            break;
        }
    }
    return ncoarse;
}

// sparsity of P for direct interpolation, columns still in fine numbering (SSS_coarsen.c:577-630)
void p_pattern(const Pattern &S, const int *mark, int ncoarse, amgb200_mat &P) {
    const int n = S.rows;
    P.num_rows = n; P.num_cols = ncoarse;
    P.row_ptr = zalloc<int>((size_t)n + 1);
    for (int i = 0; i < n; ++i) {
        int c = 0;
        if (mark[i] == F_PT) { for (int q = S.ptr[i]; q < S.ptr[i + 1]; ++q) if (mark[S.idx[q]] == C_PT) ++c; }
        else if (mark[i] == C_PT) c = 1;
        P.row_ptr[i + 1] = P.row_ptr[i] + c;
    }
    P.num_nnzs = P.row_ptr[n];
    P.col_idx = zalloc<int>((size_t)P.num_nnzs);
    P.val = zalloc<double>((size_t)P.num_nnzs);
    int w = 0;
    for (int i = 0; i < n; ++i) {
        if (mark[i] == F_PT) { for (int q = S.ptr[i]; q < S.ptr[i + 1]; ++q) { const int k = S.idx[q]; if (mark[k] == C_PT) P.col_idx[w++] = k; } }
        else if (mark[i] == C_PT) P.col_idx[w++] = i;
    }
}

// drop small interpolation weights and rescale (SSS_inter.cu:16-102)
void truncate_p(amgb200_mat &P, double eps) {
    const int n = P.num_rows;
    int kept = 0, wj = 0, wv = 0;
    for (int i = 0; i < n; ++i) {
        const int b = P.row_ptr[i], e = P.row_ptr[i + 1];
        P.row_ptr[i] = kept;
        double lo = 0, hi = 0, sneg = 0, spos = 0, tneg = 0, tpos = 0;
        for (int k = b; k < e; ++k) {
            const double v = P.val[k];
            if (v > 0) { spos += v; if (v > hi) hi = v; }
            else if (v < 0) { sneg += v; if (v < lo) lo = v; }
        }
        hi *= eps; lo *= eps;
        for (int k = b; k < e; ++k) {
            const double v = P.val[k];
            if (v >= hi) { ++kept; P.col_idx[wj++] = P.col_idx[k]; tpos += v; }
            else if (v <= lo) { ++kept; P.col_idx[wj++] = P.col_idx[k]; tneg += v; }
        }
        const double fpos = (tpos > TINY) ? spos / tpos : 1.0;
        const double fneg = (tneg < -TINY) ? sneg / tneg : 1.0;
        for (int k = b; k < e; ++k) {
            const double v = P.val[k];
            if (v >= hi) P.val[wv++] = v * fpos;
            else if (v <= lo) P.val[wv++] = v * fneg;
        }
    }
    P.num_nnzs = P.row_ptr[n] = kept;
    if (kept > 0) {
        P.col_idx = (int *)realloc(P.col_idx, (size_t)kept * sizeof(int));
        P.val = (double *)realloc(P.val, (size_t)kept * sizeof(double));
    }
}

// direct interpolation weights (SSS_inter.cu:400-547)
void interp_direct(const amgb200_mat &A, const int *mark, amgb200_mat &P, const amgb200_pars &pr) {
    const int n = A.num_rows;
    double aii = 0;                                   // carried across rows like the reference's
    for (int i = 0; i < n; ++i) {
        const int b = A.row_ptr[i], e = A.row_ptr[i + 1];
        int dk = b;
        for (; dk < e; ++dk) if (A.col_idx[dk] == i) { aii = A.val[dk]; break; }
        if (mark[i] == F_PT) {
            double neg_all = 0, neg_p = 0, pos_all = 0, pos_p = 0;
            int npos = 0;
            for (int k = b; k < e; ++k) {
                if (k == dk) continue;
                bool interp = false;
                for (int q = P.row_ptr[i]; q < P.row_ptr[i + 1]; ++q) if (P.col_idx[q] == A.col_idx[k]) { interp = true; break; }
                const double a = A.val[k];
                if (a > 0) { pos_all += a; if (interp) { pos_p += a; ++npos; } }
                else { neg_all += a; if (interp) neg_p += a; }
            }
            const double alpha = neg_all / neg_p;
            double beta;
            if (npos > 0) beta = pos_all / pos_p;
            else { beta = 0.0; aii += pos_all; }
            for (int q = P.row_ptr[i]; q < P.row_ptr[i + 1]; ++q) {
                const int c = P.col_idx[q];
                int l = b;
                for (; l < e; ++l) if (A.col_idx[l] == c) break;
                const double a = A.val[l];
                P.val[q] = (a > 0) ? (-beta * a / aii) : (-alpha * a / aii);
            }
        } else if (mark[i] == C_PT) {
            P.val[P.row_ptr[i]] = 1.0;
        }
    }
    std::vector<int> cnum(n, 0);
    int nc = 0;
    for (int i = 0; i < n; ++i) if (mark[i] == C_PT) cnum[i] = nc++;
    P.num_cols = nc;
    for (int k = 0; k < P.num_nnzs; ++k) P.col_idx[k] = cnum[P.col_idx[k]];
    truncate_p(P, pr.trunc_threshold);
}

amgb200_mat transpose(const amgb200_mat &A) {       // SSS_matvec.c:330-387
    amgb200_mat T;
    T.num_rows = A.num_cols; T.num_cols = A.num_rows; T.num_nnzs = A.num_nnzs;
    T.row_ptr = zalloc<int>((size_t)T.num_rows + 1);
    T.col_idx = zalloc<int>((size_t)A.num_nnzs);
    T.val = zalloc<double>((size_t)A.num_nnzs);
    for (int k = 0; k < A.num_nnzs; ++k) T.row_ptr[A.col_idx[k] + 1]++;
    for (int i = 0; i < T.num_rows; ++i) T.row_ptr[i + 1] += T.row_ptr[i];
    std::vector<int> fill(T.row_ptr, T.row_ptr + T.num_rows);
    for (int i = 0; i < A.num_rows; ++i)
        for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
            const int w = fill[A.col_idx[k]]++;
            T.col_idx[w] = i; T.val[w] = A.val[k];
        }
    return T;
}

// Galerkin product R*A*P (SSS_matvec.c:398-534): diagonal slot first, then columns in
// discovery order; products accumulated as (r*a)*p in traversal order.
amgb200_mat galerkin(const amgb200_mat &R, const amgb200_mat &A, const amgb200_mat &P) {
    // Rows of the product are independent; the reference's markers only have to tell "already discovered in this row"
    // (slot[i3] >= row0) and "fine column already expanded for this row" (seen[i2] == ic).  Each thread therefore walks a
    // contiguous block of coarse rows with its own marker arrays: pass 1 counts, a prefix sum places the rows, pass 2 fills.
    // The arithmetic and the order of every row are exactly those of the sequential loop.
    const int nc = R.num_rows, nf = A.num_rows;
    int *ptr = zalloc<int>((size_t)nc + 1);
    const int nt = std::max(1, std::min(omp_get_max_threads(), nc / 4096 + 1));
    std::vector<std::vector<int>> slot_t(nt), seen_t(nt);
    amgb200_mat C;
    C.num_rows = C.num_cols = nc; C.row_ptr = ptr; C.col_idx = nullptr; C.val = nullptr; C.num_nnzs = 0;
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const int r0 = (int)((long long)nc * t / nt), r1 = (int)((long long)nc * (t + 1) / nt);
        std::vector<int> &slot = slot_t[t], &seen = seen_t[t];
        slot.assign(nc, -1);
        seen.assign(nf, -1);
        int cnt = 0;
        for (int ic = r0; ic < r1; ++ic) {            // pass 1: row lengths
            const int row0 = cnt;
            slot[ic] = cnt++;
            for (int a = R.row_ptr[ic]; a < R.row_ptr[ic + 1]; ++a) {
                const int i1 = R.col_idx[a];
                for (int bq = A.row_ptr[i1]; bq < A.row_ptr[i1 + 1]; ++bq) {
                    const int i2 = A.col_idx[bq];
                    if (seen[i2] == ic) continue;
                    seen[i2] = ic;
                    for (int c = P.row_ptr[i2]; c < P.row_ptr[i2 + 1]; ++c) {
                        const int i3 = P.col_idx[c];
                        if (slot[i3] < row0) slot[i3] = cnt++;
                    }
                }
            }
            ptr[ic + 1] = cnt - row0;
        }
#pragma omp barrier
#pragma omp single
        {
            long long tot = 0;
            for (int ic = 0; ic < nc; ++ic) { const int len = ptr[ic + 1]; ptr[ic] = (int)tot; tot += len; }
            if (tot > 2147483647LL) { fprintf(stderr, "amgb200_setup: coarse matrix exceeds 2^31 entries\n"); exit(-20); }
            ptr[nc] = (int)tot;
            C.num_nnzs = (int)tot;
            C.col_idx = zalloc<int>((size_t)tot);
            C.val = zalloc<double>((size_t)tot);
        }
        std::fill(slot.begin(), slot.end(), -1);
        std::fill(seen.begin(), seen.end(), -1);
        for (int ic = r0; ic < r1; ++ic) {            // pass 2: fill
            const int row0 = ptr[ic];
            cnt = row0;
            slot[ic] = cnt; C.col_idx[cnt] = ic; C.val[cnt] = 0.0; ++cnt;
            for (int a = R.row_ptr[ic]; a < R.row_ptr[ic + 1]; ++a) {
                const double r = R.val[a];
                const int i1 = R.col_idx[a];
                for (int bq = A.row_ptr[i1]; bq < A.row_ptr[i1 + 1]; ++bq) {
                    const double ra = r * A.val[bq];
                    const int i2 = A.col_idx[bq];
                    if (seen[i2] != ic) {
                        seen[i2] = ic;
                        for (int c = P.row_ptr[i2]; c < P.row_ptr[i2 + 1]; ++c) {
                            const double rap = ra * P.val[c];
                            const int i3 = P.col_idx[c];
                            if (slot[i3] < row0) { slot[i3] = cnt; C.val[cnt] = rap; C.col_idx[cnt] = i3; ++cnt; }
                            else C.val[slot[i3]] += rap;
                        }
                    } else {
                        for (int c = P.row_ptr[i2]; c < P.row_ptr[i2 + 1]; ++c) C.val[slot[P.col_idx[c]]] += ra * P.val[c];
                    }
                }
            }
        }
    }
    return C;
}

void free_mat(amgb200_mat &M) {
    free(M.row_ptr); free(M.col_idx); free(M.val);
    M.row_ptr = nullptr; M.col_idx = nullptr; M.val = nullptr;
}

amgb200_vec make_vec(int n) { amgb200_vec v; v.n = n; v.d = zalloc<double>((size_t)n); return v; }

}  // namespace

extern "C" void amgb200_default_pars(amgb200_pars *p) {      // SSS_main.c:25-64
    memset(p, 0, sizeof(*p));
    p->smoother = 2; p->max_it = 100; p->tol = 1e-6; p->ctol = 1e-7; p->max_levels = 30; p->coarse_dof = 10;
    p->cycle_type = 1; p->cf_order = 1; p->pre_iter = 2; p->post_iter = 2; p->relax = 1.0; p->poly_deg = 3;
    p->cs_type = 1; p->interp_type = 1; p->max_row_sum = 0.9; p->strong_threshold = 0.3; p->trunc_threshold = 0.2;
}

extern "C" void amgb200_setup(amgb200_amg *mg, const amgb200_mat *A, const amgb200_pars *pars, int verbose) {
    const char *e = getenv("AMGB200_DEVICE_INTERP"), *e2 = getenv("AMGB200_DEVICE_RAP");
    amgb200_setup_ex(mg, A, pars, verbose, (e && atoi(e) ? AMGB200_SETUP_DEVICE_INTERP : 0) | (e2 && atoi(e2) ? AMGB200_SETUP_DEVICE_RAP : 0));
}

extern "C" void amgb200_setup_ex(amgb200_amg *mg, const amgb200_mat *A, const amgb200_pars *pars, int verbose, int flags) {
    if (pars->cs_type != 1 || pars->interp_type != 1) {
        fprintf(stderr, "amgb200_setup: only RS coarsening (cs_type=1) with direct interpolation (interp_type=1) is implemented\n");
        exit(-12);
    }
    memset(mg, 0, sizeof(*mg));
    mg->cg = zalloc<amgb200_comp>((size_t)pars->max_levels);
    mg->pars = *pars;
    const int min_cdof = std::max(pars->coarse_dof, 10);
    const int m = A->num_rows;
    std::vector<int> mark(m, 0);

    amgb200_mat &A0 = mg->cg[0].A;
    A0.num_rows = A->num_rows; A0.num_cols = A->num_cols; A0.num_nnzs = A->num_nnzs;
    A0.row_ptr = zalloc<int>((size_t)m + 1);
    A0.col_idx = zalloc<int>((size_t)A->num_nnzs);
    A0.val = zalloc<double>((size_t)A->num_nnzs);
    memcpy(A0.row_ptr, A->row_ptr, ((size_t)m + 1) * sizeof(int));
    memcpy(A0.col_idx, A->col_idx, (size_t)A->num_nnzs * sizeof(int));
    memcpy(A0.val, A->val, (size_t)A->num_nnzs * sizeof(double));

    int lvl = 0;
    while (mg->cg[lvl].A.num_rows > min_cdof && lvl < pars->max_levels - 1) {
        amgb200_comp &L = mg->cg[lvl];
        Pattern S;
        const double tt0 = omp_get_wtime();
        strength(L.A, *pars, S);
        const double tt1 = omp_get_wtime();
        int nc = rs_split(S, mark.data());
        const double tt2 = omp_get_wtime();
        if (nc <= 0) {
            if (verbose) { printf("### WARNING: Could not find any C-variables!\n"); printf("### WARNING: RS coarsening on level-%d failed!\n", lvl); }
            break;
        }
        nc = clean_ff(S, mark.data(), nc);
        p_pattern(S, mark.data(), nc, L.P);
        if (L.P.num_cols < min_cdof) break;
        if (L.P.num_rows > L.P.num_cols * 10 && verbose) {
            printf("### WARNING: Coarsening might be too aggressive!\n");
            printf("### WARNING: Lvl = %d ,Fine level = %d, coarse level = %d. Discard!\n", lvl, L.P.num_rows, L.P.num_cols);
        }
        L.cfmark.n = L.A.num_rows;
        L.cfmark.d = zalloc<int>((size_t)L.A.num_rows);
        memcpy(L.cfmark.d, mark.data(), (size_t)L.A.num_rows * sizeof(int));
        const double tt3 = omp_get_wtime();
        if (!((flags & AMGB200_SETUP_DEVICE_INTERP) && amgb200_interp_device(&L.A, mark.data(), &L.P, pars->trunc_threshold) == 0))
            interp_direct(L.A, mark.data(), L.P, *pars);
        const double tt4 = omp_get_wtime();
        double tt5;
        if ((flags & AMGB200_SETUP_DEVICE_RAP) && amgb200_rap_device(&L.A, &L.P, &L.R, &mg->cg[lvl + 1].A) == 0) tt5 = omp_get_wtime();
        else {
            L.R = transpose(L.P);
            tt5 = omp_get_wtime();
            mg->cg[lvl + 1].A = galerkin(L.R, L.A, L.P);
        }
        if (verbose >= 2) printf("[setup] level %d: strength %.3f split %.3f clean+pattern %.3f interp %.3f transpose %.3f galerkin %.3f s\n", lvl, tt1 - tt0, tt2 - tt1, tt3 - tt2, tt4 - tt3, tt5 - tt4, omp_get_wtime() - tt5);
        if (L.A.num_nnzs / L.A.num_rows > L.A.num_cols * 0.2) {       // (sic) tests the *fine* level, integer division
            if (verbose) { printf("### WARNING: Coarse matrix is too dense!\n"); printf("### WARNING: m = n = %d, nnz = %d!\n", L.A.num_cols, L.A.num_nnzs); }
            free_mat(mg->cg[lvl + 1].A);
            break;
        }
        ++lvl;
    }
    mg->num_levels = lvl + 1;
    mg->cg[0].wp = make_vec(m);
    for (int l = 1; l < mg->num_levels; ++l) {
        const int mm = mg->cg[l].A.num_rows;
        mg->cg[l].b = make_vec(mm);
        mg->cg[l].x = make_vec(mm);
        mg->cg[l].wp = make_vec(2 * mm);
    }
    if (verbose) {
        printf("-----------------------------------------------------------\n");
        printf("  Level   Num of rows   Num of nonzeros   Avg. NNZ / row   \n");
        printf("-----------------------------------------------------------\n");
        double gc = 0, oc = 0;
        for (int l = 0; l < mg->num_levels; ++l) {
            const amgb200_mat &M = mg->cg[l].A;
            printf("%5d %13d %17d %14.2lf\n", l, M.num_rows, M.num_nnzs, (double)M.num_nnzs / M.num_rows);
            gc += M.num_rows; oc += M.num_nnzs;
        }
        printf("-----------------------------------------------------------\n");
        printf("  Grid complexity = %.3lf  |  Operator complexity = %.3lf\n", gc / mg->cg[0].A.num_rows, oc / mg->cg[0].A.num_nnzs);
        printf("-----------------------------------------------------------\n");
    }
}

extern "C" void amgb200_amg_destroy(amgb200_amg *mg) {       // SSS_matvec.c:202-228
    if (!mg || !mg->cg) return;
    const int nl = std::max(1, mg->num_levels);
    for (int l = 0; l < nl; ++l) {
        free_mat(mg->cg[l].A); free_mat(mg->cg[l].P); free_mat(mg->cg[l].R);
        if (l > 0) { free(mg->cg[l].b.d); free(mg->cg[l].x.d); }   // level-0 x/b alias the caller's vectors
        free(mg->cg[l].wp.d); free(mg->cg[l].cfmark.d);
    }
    free(mg->cg);
    memset(mg, 0, sizeof(*mg));
}
