// Host-side analysis of a level: Gauss-Seidel wavefront schedule and device matrix layouts.
#pragma once
#include <cstdlib>
#include <vector>

#include "../../include/amg_b200.h"

namespace amgb200 {

enum MatKind { KIND_SELL = 0, KIND_CSR = 1 };

// uninitialised host array (std::vector would zero-fill ~1 GB single-threaded before the parallel fill)
template <class T>
struct RawBuf {
    T *p = nullptr;
    size_t n = 0;
    RawBuf() = default;
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    RawBuf(RawBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    RawBuf &operator=(RawBuf &&o) noexcept { if (this != &o) { free(p); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~RawBuf() { free(p); }
    void resize(size_t m) { free(p); p = m ? (T *)malloc(m * sizeof(T)) : nullptr; n = m; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T *data() { return p; }
    const T *data() const { return p; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

// Row schedule of one smoothed level.  "Schedule numbering" k = position of a row in the order
// [F-pass wavefront 0 | F-pass wavefront 1 | ... | C-pass wavefront 0 | ...], rows inside a
// wavefront in ascending natural index.  All device vectors of the level live in this numbering.
struct Schedule {
    int n = 0;
    std::vector<int> order;        // schedule position -> natural row
    std::vector<int> pos;          // natural row -> schedule position
    int pass_rows[2] = {0, 0};     // rows in the F pass (mark != 1) and the C pass (mark == 1)
    int wf_count[2] = {0, 0};      // dependency-DAG depth of each pass
    std::vector<int> wf_row_ptr;   // wf_count[0]+wf_count[1]+1 entries: first schedule row of each wavefront
    bool pattern_symmetric = true; // informational: a_ij stored <=> a_ji stored (same-pass pairs)
    int rows_without_diag = 0;
};

// A matrix permuted into schedule numbering, in the layout the kernels stream.
//  KIND_SELL: sliced ELLPACK, slice height 32 (one warp, one thread per row); entry k of lane r of
//             slice s sits at slice_ptr[s] + 32*k + r; padding has col = -1, val = 0.  Slices never
//             straddle a wavefront boundary.  Row-internal storage order is the reference's.
//  KIND_CSR : plain CSR (one warp per row, lanes stride the row).
struct DevLayout {
    int kind = KIND_SELL;
    int nrows = 0, ncols = 0;
    long long nnz = 0;             // true (unpadded) entries
    int max_row = 0;
    std::vector<int> slice_row;    // SELL: nslices+1, first schedule row of each slice
    std::vector<long long> slice_ptr;  // SELL: nslices+1
    std::vector<int> rptr;         // CSR: nrows+1
    RawBuf<int> col;
    RawBuf<double> val;
    std::vector<int> wf_item_ptr;  // (A of smoothed levels) first item (slice | row) of each wavefront
    std::vector<int> split;        // CSR + wavefronts: per row, index of the first entry whose column lies in the
                                   // cyclically preceding wavefront (row length if none)
    std::vector<unsigned> late;    // CSR + wavefronts: one bit per entry, set when its column lies in that preceding wavefront
    std::vector<unsigned> late2;   // ... when it lies in the wavefront before that one (cyclic distance 2; empty when W < 4)
    int nitems() const { return kind == KIND_SELL ? (int)slice_row.size() - 1 : nrows; }
};

// CSR + wavefronts, re-packed for the streaming single-CTA smoother (kernels.cuh, gs_stream_cta_kernel): ONE
// contiguous, 16-byte aligned block per wavefront, so that a single bulk-async copy brings everything the
// wavefront needs into shared memory.
//   block  : int32 nrows, int32 first_row & ~1, int32 rhs doubles to fetch (even), int32 block bytes
//            int32 rec_off[nrows]  (bytes from the block start, padded to a multiple of 4 ints)
//            records
//   record : int32 row, int32 prefix_pad, int32 len_pad, int32 nlate, double diag, double scratch     (32 bytes)
//            double val[len_pad]   prefix entries, zero padding to a multiple of 8, suffix entries, padding to 8
//            int32  col[len_pad]   (-1 for padding and for the diagonal entry, whose value moves to `diag`)
//            {double val; int32 pos; int32 col} late[n1]    entries whose column lies in the preceding wavefront
//            {double val; int32 pos; int32 col} late2[n2]   entries whose column lies in the wavefront before that one
//            with nlate = n1 | n2 << 16: three wavefronts are in flight in the kernel, so the product pass of wavefront
//            g only sees x through wavefront g-3; distance-2 entries are redone after g-2, distance-1 entries after g-1
// The storage order of a row is untouched: padding only inserts exact no-ops (+0.0 products).
struct StreamLayout {
    RawBuf<unsigned char> data;
    std::vector<int> blk_ptr;      // W+1, in units of 16 bytes
    int max_block = 0;             // largest block incl. its right-hand-side segment, bytes
    long long mean_block = 0;
};
struct StreamLate { double val; int pos; int col; };
void build_stream(const DevLayout &L, StreamLayout &S);

// The same idea for levels that need several SMs (kernels.cuh, gs_stream_cluster_kernel): row i of a wavefront
// (i = schedule row - first row of the wavefront) belongs to CTA i % C of a C-CTA cluster, local index i / C; block
// (w, c) = blk_ptr[w*C + c] holds CTA c's rows of wavefront w.  x stays in global memory for entries at cyclic
// wavefront distance > late_dist (2 or 3); nearer entries ("late") are read from the exchange buffers in
// shared memory (every CTA receives every x of the last late_dist + 1 wavefronts).
//   block  : int32 rows of this CTA, first row of the wavefront, width of the wavefront, block bytes,
//            int32 nflat, flat_off (bytes), 0, 0                                                       (32 bytes)
//            int32 rec_off[rows] (padded to a multiple of 4), records (header and val/col arrays as in StreamLayout, no
//            per-row late lists; the prefix ends at the first late entry),
//            flat late list: {double val; int32 dst; int32 col; int32 src; int32 0} for ALL late entries of the block,
//            dst = byte offset of the product inside the block, src = (distance-1) | index within its wavefront << 2 --
//            one flat list so that all lanes of a consumer group share the post-barrier patching evenly
struct StreamLateC { double val; int pos; int col; int src; int pad; };
struct ClusterStreamLayout {
    RawBuf<unsigned char> data;
    std::vector<int> blk_ptr;      // W*C+1, in units of 16 bytes
    int C = 0;
    int max_block = 0;             // bytes
    long long mean_block = 0;
    int max_local = 0;             // largest number of rows one CTA gets from one wavefront
    int max_width = 0;             // widest wavefront (rows)
    bool filled = false;           // false: the sizing pass showed the blocks cannot fit (data left empty)
    int late_dist = 2;             // 2 or 3: wavefront distance served from the exchange buffers (late_dist + 1 buffers per CTA)
};
// smem_budget >= 0: shared memory available for 3 exchange buffers + a ring of two blocks; the fill is skipped when it cannot fit
// late_dist: entries at cyclic wavefront distance <= late_dist (2 or 3; 3 needs >= 6 wavefronts) are served from the exchange buffers
void build_stream_cluster(const amgb200_mat &A, const Schedule &S, int C, ClusterStreamLayout &SL, long long smem_budget = -1, int late_dist = 3);

// mark == nullptr: single pass over all rows in natural order (no C/F ordering).
void build_schedule(const amgb200_mat &A, const int *mark, Schedule &S);
void identity_schedule(int n, Schedule &S);

// rows of M taken in `row_order` (schedule position -> natural row; nullptr = identity),
// columns renumbered through `col_pos` (natural -> schedule; nullptr = identity).
// `breaks` (optional, ascending schedule-row offsets incl. 0 and nrows) forces slice boundaries.
void build_layout(const amgb200_mat &M, const int *row_order, const int *col_pos, int kind,
                  const std::vector<int> *breaks, DevLayout &L);

// structure only (slices, widths, offsets, wavefront item table) of the SELL-32 layout; col/val are left empty
void build_sell_structure(const amgb200_mat &M, const int *row_order, const std::vector<int> *breaks, DevLayout &L);

int choose_kind(const amgb200_mat &M, double sell_max_mean);

// Ticket list of the fused residual (+) restriction launch (kernels.cuh, resid_restrict_kernel).  la: SELL layout of A_l (rows in
// schedule Sf), lr: SELL layout of R_l (rows in schedule Sc of level l+1), Rm: R_l as the reference stores it (natural numbering).
// The slices of A are binned into nch chunks by the natural index of their first row; a slice of R is keyed by the last chunk
// whose residual rows it reads; tickets: A(chunk 0) .. A(chunk c), R(last chunk c - lag) ..; slices of R that read 32 chunks or
// more wait for everything and come last.  The list is cut into blocks of `tickets` entries that are all slices of A of one chunk
// or all slices of R (padded with FUSED_NOP).  Invariant: every block only depends on blocks with smaller numbers.
constexpr int FUSED_NOP = 0x7fffffff;
struct FusedPlan {
    std::vector<int> work;             // >= 0: slice of A, < 0: ~slice of R, FUSED_NOP: padding
    std::vector<int> block_info;       // A: chunk | count << 16 ; R: 0x80000000 | first chunk | last chunk << 16
    std::vector<int> slice_chunk;      // slice of A -> chunk
    std::vector<unsigned> rneed;       // slice of R -> first chunk | last chunk << 16
    std::vector<unsigned> chunk_items; // slices of A per chunk
    int nch = 0, tickets = 0;
};
void build_fused_plan(const DevLayout &la, const DevLayout &lr, const Schedule &Sf, const Schedule &Sc, const amgb200_mat &Rm, int nch, int lag, int tickets, FusedPlan &F);

}  // namespace amgb200
