// Device-resident AMG hierarchy and the solve-phase driver (V-cycle, coarsest-level Krylov
// solvers, outer iteration) -- the B200 replacement of the reference's
// amg/Solve/SSS_SOLVE.c, SSS_cycle.cu, SSS_smooth.c and SSS_cuda.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <omp.h>

#include "../../include/amg_b200.h"
#include "analysis.h"
#include "kernels.cuh"

using namespace amgb200;

namespace {

long long g_launches = 0;

#define CUDA_CHECK(call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            fprintf(stderr, "libamgb200: CUDA error %s at %s:%d: %s (no CPU fallback exists)\n",      \
                    cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_));                \
            exit(70);                                                                                 \
        }                                                                                             \
    } while (0)

#define LAUNCH(kern, grid, block, stream, ...)                     \
    do {                                                           \
        kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);       \
        ++g_launches;                                              \
        CUDA_CHECK(cudaGetLastError());                            \
    } while (0)

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Device memory pool: cudaMalloc/cudaFree of ~1 GB cost 0.2-0.6 s per SSS_amg_solve call; blocks of a freed
// hierarchy are kept (size-class free lists, 2 MB granularity) and reused by the next upload.  Released at exit.
struct DevPool {
    std::vector<std::pair<size_t, void *>> free_blocks;
    std::vector<std::pair<void *, size_t>> live;
    static size_t round(size_t b) { const size_t g = b < (1u << 20) ? 512 : (2u << 20); return (std::max<size_t>(b, 1) + g - 1) / g * g; }
    void *get(size_t bytes) {
        const size_t need = round(bytes);
        int best = -1;
        for (int i = 0; i < (int)free_blocks.size(); ++i)
            if (free_blocks[i].first >= need && free_blocks[i].first <= need + need / 4 && (best < 0 || free_blocks[i].first < free_blocks[best].first)) best = i;
        void *p = nullptr;
        size_t sz = need;
        if (best >= 0) { p = free_blocks[best].second; sz = free_blocks[best].first; free_blocks.erase(free_blocks.begin() + best); }
        else {
            cudaError_t e = cudaMalloc(&p, need);
            if (e != cudaSuccess) { trim(); e = cudaMalloc(&p, need); }
            CUDA_CHECK(e);
        }
        live.emplace_back(p, sz);
        return p;
    }
    void put(void *p) {
        if (!p) return;
        for (size_t i = 0; i < live.size(); ++i)
            if (live[i].first == p) { free_blocks.emplace_back(live[i].second, p); live.erase(live.begin() + i); return; }
        cudaFree(p);
    }
    void trim() { for (auto &b : free_blocks) cudaFree(b.second); free_blocks.clear(); }
};
// Everything process-global below is keyed by the CUDA device it belongs to (amgb200_options.device lets one process keep
// hierarchies on several GPUs: a block freed on device A must never be handed to device B, and per-function attributes are
// per device) and guarded by one mutex (uploads / frees from several host threads).
constexpr int MAX_DEVICES = 64;
std::mutex &global_mutex() { static std::mutex *m = new std::mutex(); return *m; }
int current_device() { int d = 0; cudaGetDevice(&d); return d < 0 || d >= MAX_DEVICES ? 0 : d; }
DevPool &pool(int dev) { static DevPool *p = new DevPool[MAX_DEVICES]; return p[dev]; }                 // intentionally leaked: the driver reclaims at process exit
std::vector<double *> &pinned_scalars() { static std::vector<double *> *v = new std::vector<double *>(); return *v; }   // 64-byte pinned slots (host memory: device independent)
// true exactly once per (kernel, device): cudaFuncSetAttribute has to be repeated on every device, and for every template
// instance of a kernel
bool first_use(const void *kernel) {
    static std::vector<std::pair<const void *, int>> *seen = new std::vector<std::pair<const void *, int>>();
    std::lock_guard<std::mutex> lk(global_mutex());
    const std::pair<const void *, int> key(kernel, current_device());
    for (const auto &e : *seen) if (e == key) return false;
    seen->push_back(key);
    return true;
}
void dev_free(const void *p) {
    if (!p) return;
    cudaPointerAttributes at;
    int dev = current_device();
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.device >= 0 && at.device < MAX_DEVICES) dev = at.device;
    std::lock_guard<std::mutex> lk(global_mutex());
    pool(dev).put(const_cast<void *>(p));
}

template <class T>
T *dev_alloc(size_t n) {
    std::lock_guard<std::mutex> lk(global_mutex());
    return (T *)pool(current_device()).get(std::max<size_t>(n, 1) * sizeof(T));
}
// Host -> device copy of pageable memory through a small ring of pinned staging buffers (allocated once per
// process): OpenMP threads fill a staging buffer while the DMA engine drains the previous one.  ~4x faster than
// cudaMemcpy from pageable memory for the ~1 GB of a hierarchy, without pinning the caller's arrays.
struct Stager {
    static constexpr size_t CHUNK = 32u << 20;
    static constexpr int NBUF = 3;
    char *buf[NBUF] = {nullptr, nullptr, nullptr};
    cudaEvent_t done[NBUF];
    cudaStream_t stream = nullptr;
    bool ok = false;
    void init() {
        if (ok) return;
        CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (int i = 0; i < NBUF; ++i) { CUDA_CHECK(cudaMallocHost((void **)&buf[i], CHUNK)); CUDA_CHECK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming)); }
        ok = true;
    }
    std::mutex mu;                                       // (uploads from helper threads share the ring)
    void copy(void *dst, const void *src, size_t bytes) {
        std::lock_guard<std::mutex> lk(mu);
        init();
        size_t off = 0;
        int slot = 0;
        bool used[NBUF] = {false, false, false};
        while (off < bytes) {
            const size_t m = std::min(CHUNK, bytes - off);
            if (used[slot]) CUDA_CHECK(cudaEventSynchronize(done[slot]));
            const char *sp = (const char *)src + off;
            char *bp = buf[slot];
            const size_t piece = (m + 7) / 8;
#pragma omp parallel for schedule(static) num_threads(8)
            for (int t = 0; t < 8; ++t) {
                const size_t a = std::min(m, piece * t), b = std::min(m, piece * (t + 1));
                if (b > a) memcpy(bp + a, sp + a, b - a);
            }
            CUDA_CHECK(cudaMemcpyAsync((char *)dst + off, bp, m, cudaMemcpyHostToDevice, stream));
            CUDA_CHECK(cudaEventRecord(done[slot], stream));
            used[slot] = true;
            slot = (slot + 1) % NBUF;
            off += m;
        }
        CUDA_CHECK(cudaStreamSynchronize(stream));
    }
};
Stager &stager() { static Stager *s = new Stager[MAX_DEVICES]; return s[current_device()]; }   // (its stream belongs to one device)
Stager &prefetch_stager() { static Stager *s = new Stager[MAX_DEVICES]; return s[current_device()]; }

// Raw CSR arrays of the matrices whose SELL layout is filled on the device, copied to the device by a helper thread WHILE the
// calling thread runs the schedule analysis (both take ~40 ms at 128^3 and used to run one after the other).  upload_sell()
// takes a finished copy from here or, if the matrix is not listed (or already taken), uploads it itself.
struct RawPrefetch {
    struct Item { const amgb200_mat *M; int *rp = nullptr; int *ci = nullptr; double *va = nullptr; std::atomic<int> ready{0}; bool taken = false; };
    std::vector<std::unique_ptr<Item>> items;
    std::thread worker;
    void start(const std::vector<const amgb200_mat *> &mats, int device) {
        for (const amgb200_mat *M : mats) { items.emplace_back(new Item()); items.back()->M = M; }
        worker = std::thread([this, device]() {
            CUDA_CHECK(cudaSetDevice(device));
            for (auto &it : items) {
                const amgb200_mat &M = *it->M;
                const size_t nz = (size_t)M.row_ptr[M.num_rows];
                auto put = [&](const void *src, size_t bytes) -> void * {
                    void *d = dev_alloc<char>(std::max<size_t>(bytes, 1));
                    if (bytes >= (4u << 20)) prefetch_stager().copy(d, src, bytes);
                    else if (bytes) CUDA_CHECK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
                    return d;
                };
                it->rp = (int *)put(M.row_ptr, ((size_t)M.num_rows + 1) * sizeof(int));
                it->ci = (int *)put(M.col_idx, nz * sizeof(int));
                it->va = (double *)put(M.val, nz * sizeof(double));
                it->ready.store(1, std::memory_order_release);
            }
        });
    }
    bool take(const amgb200_mat &M, int *&rp, int *&ci, double *&va) {
        for (auto &it : items) {
            if (it->M == &M && !it->taken) {
                while (!it->ready.load(std::memory_order_acquire)) std::this_thread::yield();
                it->taken = true; rp = it->rp; ci = it->ci; va = it->va;
                return true;
            }
        }
        return false;
    }
    void finish(std::vector<void *> &temps) {          // copies nobody asked for are released with the other temporaries
        if (worker.joinable()) worker.join();
        for (auto &it : items) if (!it->taken) { temps.push_back(it->rp); temps.push_back(it->ci); temps.push_back(it->va); }
        items.clear();
    }
};

template <class T>
T *dev_upload_raw(const T *src, size_t n) {
    T *p = dev_alloc<T>(n);
    const size_t bytes = n * sizeof(T);
    if (bytes >= (4u << 20)) stager().copy(p, src, bytes);
    else if (bytes) CUDA_CHECK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
    return p;
}
template <class T> T *dev_upload(const std::vector<T> &v) { return dev_upload_raw(v.data(), v.size()); }
template <class T> T *dev_upload(const RawBuf<T> &v) { return dev_upload_raw(v.data(), v.size()); }

struct DevMatOwner {
    DMat v{};
    long long nnz = 0, padded = 0;
    int max_row = 0;
    bool valid = false;
    static double &t_upload() { static double t = 0; return t; }
    static void add_upload_time(double dt) { static std::mutex *m = new std::mutex(); std::lock_guard<std::mutex> lk(*m); t_upload() += dt; }
    void upload(const DevLayout &L) {
        const double t0 = now_s();
        upload_impl(L);
        add_upload_time(now_s() - t0);
    }
    void upload_impl(const DevLayout &L) {
        v.kind = L.kind; v.nrows = L.nrows; v.ncols = L.ncols; v.nitems = L.nitems(); v.max_row = L.max_row; v.recip = 0;
        v.slice_row = nullptr; v.slice_ptr = nullptr; v.rptr = nullptr; v.split = nullptr; v.late = nullptr;
        if (L.kind == KIND_SELL) { v.slice_row = dev_upload(L.slice_row); v.slice_ptr = dev_upload(L.slice_ptr); }
        else { v.rptr = dev_upload(L.rptr); if (!L.split.empty()) { v.split = dev_upload(L.split); v.late = dev_upload(L.late); } }
        v.col = dev_upload(L.col);
        v.val = dev_upload(L.val);
        nnz = L.nnz; padded = (long long)L.col.size(); max_row = L.max_row; valid = true;
    }
    // SELL layout filled on the device: L carries the slice structure only (build_sell_structure); the raw CSR arrays of M are
    // copied as they are and permuted / padded by sell_fill_kernel.  The temporaries are appended to `temps` and must stay
    // alive until the stream has been synchronised.
    void upload_sell(const DevLayout &L, const amgb200_mat &M, const int *d_order, const int *d_colpos, cudaStream_t stream, std::vector<void *> &temps,
                     RawPrefetch *pre = nullptr) {
        const double t0 = now_s();
        v.kind = KIND_SELL; v.nrows = L.nrows; v.ncols = L.ncols; v.nitems = L.nitems(); v.max_row = L.max_row; v.recip = 0;
        v.rptr = nullptr; v.split = nullptr; v.late = nullptr;
        v.slice_row = dev_upload(L.slice_row); v.slice_ptr = dev_upload(L.slice_ptr);
        const size_t total = (size_t)L.slice_ptr.back();
        int *d_col = dev_alloc<int>(total);
        double *d_val = dev_alloc<double>(total);
        v.col = d_col; v.val = d_val;
        if (total) {
            const size_t nz = (size_t)M.row_ptr[M.num_rows];
            int *d_rp = nullptr, *d_ci = nullptr;
            double *d_va = nullptr;
            if (!(pre && pre->take(M, d_rp, d_ci, d_va))) {
                d_rp = dev_upload_raw(M.row_ptr, (size_t)M.num_rows + 1);
                d_ci = dev_upload_raw(M.col_idx, nz);
                d_va = dev_upload_raw(M.val, nz);
            }
            temps.push_back(d_rp); temps.push_back(d_ci); temps.push_back(d_va);
            const int ns = L.nitems();
            LAUNCH(sell_fill_kernel, std::max(1, std::min((ns + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, 148 * 16)), BLOCK, stream,
                   ns, v.slice_row, v.slice_ptr, d_order, d_colpos, (const int *)d_rp, (const int *)d_ci, (const double *)d_va, d_col, d_val);
        }
        nnz = L.nnz; padded = (long long)total; max_row = L.max_row; valid = true;
        add_upload_time(now_s() - t0);
    }
    void release() {
        if (!valid) return;
        dev_free(v.slice_row); dev_free(v.slice_ptr); dev_free(v.rptr); dev_free(v.split); dev_free(v.late);
        dev_free(v.col); dev_free(v.val);
        valid = false;
    }
};

struct Level {
    int n = 0;
    DevMatOwner A, P, R;               // A: layout used by the smoother
    DevMatOwner Asp;                   // A in the layout used by SpMV/residual when it differs from the smoother's
    const DMat &spmvA() const { return Asp.valid ? Asp.v : A.v; }
    int *d_order = nullptr;            // schedule position -> natural row
    double *x = nullptr, *b = nullptr, *wp = nullptr;
    bool smoothed = false, ordered = false;
    int W = 0, wf_count[2] = {0, 0}, pass_items[2] = {0, 0}, pass_rows[2] = {0, 0}, max_width = 0;
    int *d_item_wf = nullptr, *d_wf_item_ptr = nullptr, *d_wf_row_ptr = nullptr;
    bool pattern_symmetric = true;
    int strategy = 0;                  // 0 = parallel passes, 2 = ordered inside one CTA, 3 = one cluster (barrier per wavefront), 4 = streaming CTA, 5 = streaming cluster, 6 = data-flow grid
    double prof_ms[4] = {0, 0, 0, 0};  // AMGB200_PROFILE: GS, residual, restrict, prolong of the last solve
    bool x_in_smem = false;            // strategy 2 only: x fits in the CTA's shared memory
    int cta_G = 1, cta_D = 1;          // strategy 2: D groups of G warps (pipeline depth D)
    int cta_cap = 0;                   // strategy 2, two-phase rows: parked suffix products per warp (0 = stream the suffix)
    unsigned char *d_stream = nullptr; // strategy 4 (streaming single CTA): per-wavefront blocks (analysis.h, StreamLayout)
    int *d_blk_ptr = nullptr;
    int stream_G = 1, stream_S = 1, stream_D = 2, stream_ring = 0;   // consumer warps per group, row slots per warp, groups (wavefronts in flight)
    long long chain_terms = 0;                                    // ordered levels: sum over wavefronts of the longest post-barrier chain (terms per sweep)
    int xc_NB = 3;                                                // strategy 5: exchange buffers per CTA (late distance + 1)
    int xc_D = 2;                                                 // strategy 5: consumer groups (wavefronts in flight)
    int xc_F = 1, xc_S = 1, xc_P = 32, xc_ring = 0, xc_cap = 0;   // strategy 5 (streaming cluster): folding warps per group, row slots, ring bytes, exchange-buffer doubles
    XRec *d_rec = nullptr;             // strategy 6 (data-flow): one {x_k, version} record per row,
    unsigned *d_hint = nullptr;        //   per-wavefront "closed" hint words of a launch,
    int hint_cap = 0;
    unsigned df_vbase = 0;             //   version base of the next launch (versions grow monotonically: records are never reset)
    int dfw_cap = 0;                   // strategy 7 (data-flow, warp per row): staging doubles per warp
    int df_grid = 0, df_sch = 20;
    bool natural = false;              // natural-order Gauss-Seidel (cf_order = 0 or no cfmark): forward sweeps use this level's
    Level *bk = nullptr;               // schedule, backward sweeps (post-smoothing) the schedule/layout/vectors of *bk
    Level *lo = nullptr;               // pre-smoothing from x = 0 (levels >= 1): the first sweep only sees the entries whose column has already
                                       // been updated -- the strictly lower triangle in schedule numbering; same schedule, shares x and b
    bool x_is_zero = false;            // x was zero-filled by the cycle and not touched since
    int *d_fb = nullptr;               // position in this level's numbering of row k of bk's numbering
    // residual (+) restriction in one launch (resid_restrict_kernel): ticket list, chunk tables, completion counters
    struct Fused {
        bool valid = false;
        int *d_work = nullptr, *d_block_info = nullptr;
        unsigned *d_chunk_items = nullptr, *d_cnt = nullptr;
        int nblocks = 0, nch = 0, grid = 0;
        unsigned epoch = 0;
        bool one_a = false, one_r = false;
        void release() { dev_free(d_work); dev_free(d_block_info); dev_free(d_chunk_items); dev_free(d_cnt); valid = false; }
    } rr;
};

}  // namespace

struct amgb200_hier {
    int nl = 0;
    std::vector<Level> L;
    amgb200_pars pars{};
    amgb200_options opt{};
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    bool exact = true;
    int max_dyn_smem = 0;
    int cluster_block = 256;
    int df_ahead = 2;                  // data-flow smoother: wavefronts ahead of the completed frontier that poll their records
    int pcg_ctas = 0;                  // coarsest-level CG in one cooperative launch: CTAs at most (0 = host-driven loop)
    double *d_partial = nullptr;       // 4 x partial_stride block partials
    int partial_stride = 0;
    double *d_scal = nullptr;          // 8 reduced scalars
    double *h_scal = nullptr;          // pinned mirror
    double *d_xnat = nullptr, *d_bnat = nullptr;
    double *kry = nullptr;             // Krylov work space on the coarsest level
    size_t kry_len = 0;
    bool profile = false;
    double phase_ms[8] = {0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double upload_s = 0, analysis_s = 0;
    double last_sumsq = 0;             // sum r_i^2 of the most recent dev_true_residual
    long long *d_dbg = nullptr;        // AMGB200_DEBUG_TIMING: per-warp cycle counters of the CTA kernel
    int item0 = -1, item1 = -1;        // >= 0: restrict the next spmv launch to this item range
    bool own_stream = true;
    // peer-memory exchange plans (multi-GPU): flag words [plan][source rank] of THIS rank, written by the peers
    static constexpr int PEER_MAX_PLANS = 8, PEER_MAX_RANKS = 64;
    unsigned *d_peer_flags = nullptr;
    struct PeerPlan {
        PeerPush *d_push = nullptr; int npush = 0, max_count = 0;
        unsigned **d_flag_ptr = nullptr; int nflag = 0;
        int *d_src = nullptr; int nsrc = 0;
        unsigned epoch = 0;
    } peer_plan[PEER_MAX_PLANS];
    std::vector<void *> ipc_opened;
};

namespace {

// ---- small device helpers -----------------------------------------------------------------
int grid_for(int n) { return std::max(1, (n + BLOCK - 1) / BLOCK); }
int red_grid(const amgb200_hier *h, int n) { return std::max(1, std::min(h->partial_stride, (n + BLOCK - 1) / BLOCK)); }

void fetch_scalars(amgb200_hier *h, int count) {
    CUDA_CHECK(cudaMemcpyAsync(h->h_scal, h->d_scal, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

struct PhaseTimer {                    // only active with AMGB200_PROFILE=1 (adds syncs)
    amgb200_hier *h; int id; int level;
    PhaseTimer(amgb200_hier *h_, int id_, int level_ = -1) : h(h_), id(id_), level(level_) { if (h->profile) CUDA_CHECK(cudaEventRecord(h->ev0, h->stream)); }
    ~PhaseTimer() {
        if (!h->profile) return;
        CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
        CUDA_CHECK(cudaEventSynchronize(h->ev1));
        float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->phase_ms[id] += ms;
        if (level >= 0 && id < 4) h->L[level].prof_ms[id] += ms;
    }
};

template <int KIND, int MODE, int RED, bool EXACT>
void launch_spmv_t(amgb200_hier *h, const DMat &A, const double *x, double *y, const double *b, double alpha) {
    const int item0 = h->item0 >= 0 ? h->item0 : 0, item1 = h->item0 >= 0 ? h->item1 : A.nitems;   // (item range: multi-GPU building blocks)
    const int grid = std::max(1, (item1 - item0 + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    if (RED != RED_NONE && grid > h->partial_stride) {
        fprintf(stderr, "libamgb200: internal error: reduction grid %d exceeds partial buffer %d\n", grid, h->partial_stride);
        exit(71);
    }
    if (KIND == 0 && A.max_row <= 8)
        LAUNCH((spmv_kernel<KIND, MODE, RED, EXACT, true>), grid, BLOCK, h->stream, A, x, y, b, alpha, h->d_partial, item0, item1);
    else
        LAUNCH((spmv_kernel<KIND, MODE, RED, EXACT>), grid, BLOCK, h->stream, A, x, y, b, alpha, h->d_partial, item0, item1);
    if (RED != RED_NONE) LAUNCH(reduce_partials_kernel, 1, BLOCK, h->stream, h->d_partial, grid, h->partial_stride, 1, 0, h->d_scal);
}
template <int KIND, bool EXACT>
void spmv_k(amgb200_hier *h, const DMat &A, int mode, int red, const double *x, double *y, const double *b, double alpha) {
    if (mode == MODE_MXY && red == RED_NONE) launch_spmv_t<KIND, MODE_MXY, RED_NONE, EXACT>(h, A, x, y, b, alpha);
    else if (mode == MODE_AMXPY && red == RED_NONE) launch_spmv_t<KIND, MODE_AMXPY, RED_NONE, EXACT>(h, A, x, y, b, alpha);
    else if (mode == MODE_RESID && red == RED_NONE) launch_spmv_t<KIND, MODE_RESID, RED_NONE, EXACT>(h, A, x, y, b, alpha);
    else if (mode == MODE_RESID && red == RED_SUMSQ) launch_spmv_t<KIND, MODE_RESID, RED_SUMSQ, EXACT>(h, A, x, y, b, alpha);
    else { fprintf(stderr, "libamgb200: unsupported spmv mode %d/%d\n", mode, red); exit(71); }
}
// y = A x | y += alpha A x | y = b - A x ; optional fused tree reduction sum(y^2) into d_scal[0]
void spmv(amgb200_hier *h, const DMat &A, int mode, int red, const double *x, double *y, const double *b, double alpha) {
    if (A.kind == KIND_SELL) spmv_k<0, true>(h, A, mode, red, x, y, b, alpha);
    else if (h->exact) spmv_k<1, true>(h, A, mode, red, x, y, b, alpha);
    else spmv_k<1, false>(h, A, mode, red, x, y, b, alpha);
}

static_assert(RR_NOP == FUSED_NOP, "padding ticket of the fused launch");
// wp = b - A x ; b_{l+1} = R wp ; x_{l+1} = 0 in one launch (levels whose A and R are both thread-per-row layouts)
const void *rr_kernel(bool one_a, bool one_r) {
    return one_a ? (one_r ? (const void *)resid_restrict_kernel<true, true> : (const void *)resid_restrict_kernel<true, false>)
                 : (one_r ? (const void *)resid_restrict_kernel<false, true> : (const void *)resid_restrict_kernel<false, false>);
}
void resid_restrict(amgb200_hier *h, int l, bool zero_xc) {
    Level &lv = h->L[l];
    Level &lc = h->L[l + 1];
    Level::Fused &f = lv.rr;
    RRPlan pl;
    pl.work = f.d_work; pl.nblocks = f.nblocks; pl.block_info = f.d_block_info; pl.chunk_items = f.d_chunk_items;
    pl.cnt = f.d_cnt; pl.nch = f.nch; pl.epoch = ++f.epoch;
    DMat A = lv.spmvA(), R = lv.R.v;
    const double *x = lv.x, *b = lv.b;
    double *r = lv.wp, *bc = lc.b, *xc = zero_xc ? lc.x : nullptr;
    void *args[] = {&A, &R, &pl, &x, &b, &r, &bc, &xc};
    CUDA_CHECK(cudaLaunchCooperativeKernel(rr_kernel(f.one_a, f.one_r), dim3(f.grid), dim3(BLOCK), args, 0, h->stream));
    ++g_launches;
}

const void *df_kernel(int sch) {
    return sch == 8 ? (const void *)gs_dataflow_kernel<8> : sch == 20 ? (const void *)gs_dataflow_kernel<20>
         : sch == 28 ? (const void *)gs_dataflow_kernel<28> : (const void *)gs_dataflow_kernel<16>;
}

// ---- Gauss-Seidel -------------------------------------------------------------------------
template <int KIND, bool EXACT>
void smooth_k(amgb200_hier *h, Level &lv, int nsweeps) {
    if (lv.strategy == 0) {
        for (int s = 0; s < nsweeps; ++s) {
            int first = 0;
            for (int p = 0; p < 2; ++p) {
                const int cnt = lv.pass_items[p];
                if (cnt > 0) {
                    const int grid = (cnt + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
                    if (KIND == 0 && lv.A.v.max_row <= 8) LAUNCH((gs_pass_kernel<KIND, EXACT, true>), grid, BLOCK, h->stream, lv.A.v, lv.b, lv.x, first, first + cnt);
                    else LAUNCH((gs_pass_kernel<KIND, EXACT>), grid, BLOCK, h->stream, lv.A.v, lv.b, lv.x, first, first + cnt);
                }
                first += cnt;
            }
        }
        return;
    }
    if (lv.strategy == 6) {
        const int need = nsweeps * lv.W;
        if (need > lv.hint_cap) {
            if (lv.d_hint) dev_free(lv.d_hint);
            lv.d_hint = dev_alloc<unsigned>((size_t)need * DF_HINT_STRIDE);
            lv.hint_cap = need;
            CUDA_CHECK(cudaMemsetAsync(lv.d_hint, 0, (size_t)need * DF_HINT_STRIDE * sizeof(unsigned), h->stream));
        }
        DMat A = lv.A.v;
        const double *b = lv.b; double *x = lv.x; XRec *rec = lv.d_rec;
        const int *iw = lv.d_item_wf, *wp = lv.d_wf_item_ptr;
        unsigned *hint = lv.d_hint;
        int W = lv.W, ns = nsweeps, ahead = h->df_ahead;
        unsigned vbase = lv.df_vbase;
        long long *dbg = h->d_dbg;
        void *args[] = {&A, &b, &x, &rec, &iw, &wp, &hint, &W, &ns, &ahead, &vbase, &dbg};
        const void *kern = df_kernel(lv.df_sch);
        CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(lv.df_grid), dim3(DF_BLOCK), args, 0, h->stream));
        ++g_launches;
        lv.df_vbase += (unsigned)nsweeps;
#ifdef AMGB200_DF_TIMING
        if (h->d_dbg) {
            long long hd[64];
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            CUDA_CHECK(cudaMemcpy(hd, h->d_dbg, sizeof(hd), cudaMemcpyDeviceToHost));
            for (int w = 0; w < 4; ++w) { const long long *o = hd + w * 8; const double ni = (double)std::max(1LL, o[6]);
                printf("   [df] warp %d: %lld items; cycles per item: static-load wait %.0f  gate %.0f  first round %.0f  poll rounds %.0f (%.1f rounds)  chain+push %.0f\n",
                       w, o[6], o[0] / ni, o[1] / ni, o[2] / ni, o[3] / ni, o[5] / ni, o[4] / ni); }
        }
#endif
        return;
    }
    if (lv.strategy == 7) {
        const int need = nsweeps * lv.W;
        if (need > lv.hint_cap) {
            if (lv.d_hint) dev_free(lv.d_hint);
            lv.d_hint = dev_alloc<unsigned>((size_t)need * DF_HINT_STRIDE);
            lv.hint_cap = need;
            CUDA_CHECK(cudaMemsetAsync(lv.d_hint, 0, (size_t)need * DF_HINT_STRIDE * sizeof(unsigned), h->stream));
        }
        if (first_use((const void *)gs_dataflow_csr_kernel)) CUDA_CHECK(cudaFuncSetAttribute(gs_dataflow_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_dyn_smem));
        DMat A = lv.A.v;
        const double *b = lv.b; double *x = lv.x; XRec *rec = lv.d_rec;
        const int *wp = lv.d_wf_item_ptr, *iw = lv.d_item_wf;
        unsigned *hint = lv.d_hint;
        int W = lv.W, ns = nsweeps, ahead = h->df_ahead, cap = lv.dfw_cap;
        unsigned vbase = lv.df_vbase;
        void *args[] = {&A, &b, &x, &rec, &wp, &iw, &hint, &W, &ns, &ahead, &vbase, &cap};
        CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)gs_dataflow_csr_kernel, dim3(lv.df_grid), dim3(DFW_BLOCK), args,
                                               (size_t)(DFW_BLOCK / 32) * (cap + 16) * sizeof(double), h->stream));
        ++g_launches;
        lv.df_vbase += (unsigned)nsweeps;
        return;
    }
    if (lv.strategy == 5) {
        auto k3 = &gs_stream_cluster_kernel<3>;
        auto k4 = &gs_stream_cluster_kernel<4>;
        if (first_use((const void *)k3)) {
            CUDA_CHECK(cudaFuncSetAttribute(k3, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            CUDA_CHECK(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_dyn_smem));
            CUDA_CHECK(cudaFuncSetAttribute(k4, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            CUDA_CHECK(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_dyn_smem));
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(XC_CTAS);
        cfg.blockDim = dim3(32 * (lv.xc_D * XC_G + 2));
        cfg.dynamicSmemBytes = XC_HDR + (size_t)lv.xc_NB * lv.xc_cap * 8 + (size_t)lv.xc_ring + 128;
        cfg.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = XC_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        auto kern = &gs_stream_cluster_kernel<3>;
        if (lv.xc_NB == 4) kern = &gs_stream_cluster_kernel<4>;
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, (const unsigned char *)lv.d_stream, (const int *)lv.d_blk_ptr, (const int *)lv.d_wf_row_ptr,
                                      (const double *)lv.b, lv.x, lv.W, nsweeps, lv.xc_F, lv.xc_S, lv.xc_P, lv.xc_D, lv.xc_ring, lv.xc_cap, lv.A.v.recip, h->d_dbg));
        ++g_launches;
#ifdef AMGB200_TIMELINE
        if (h->d_dbg) {
            long long hd[32];
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            CUDA_CHECK(cudaMemcpy(hd, h->d_dbg, sizeof(hd), cudaMemcpyDeviceToHost));
            const char *nm[11] = {"wait block", "wait GV(g-3)", "products", "group barrier", "prefix fold", "wait WF(g-1)", "late patch", "suffix fold", "quotient", "loop top", "push x"};
            const int iters = (lv.W * nsweeps + lv.xc_D - 1) / lv.xc_D;
            for (int g2 = 0; g2 < 2; ++g2) {
                printf("   cluster stream timeline CTA 0 group %d folder (cycles per wavefront, %d wavefronts, F=%d S=%d):", g2, iters, lv.xc_F, lv.xc_S);
                long long tot = 0;
                for (int i = 0; i < 11; ++i) { printf("  %s %lld", nm[i], hd[g2 * 16 + i] / iters); tot += hd[g2 * 16 + i]; }
                printf("  | total %lld\n", tot / iters);
            }
        }
#endif
        return;
    }
    if (lv.strategy == 4) {
        if (first_use((const void *)gs_stream_cta_kernel)) {
            CUDA_CHECK(cudaFuncSetAttribute(gs_stream_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_dyn_smem));
        }
        const size_t xb = ((size_t)lv.n * 8 + 127) & ~(size_t)127;
        const size_t smem = STREAM_HDR + xb + (size_t)lv.stream_ring + 128;
        gs_stream_cta_kernel<<<1, 32 * (lv.stream_D * lv.stream_G + 1), smem, h->stream>>>(lv.d_stream, lv.d_blk_ptr, lv.d_wf_item_ptr, lv.b, lv.x, lv.n, lv.W, nsweeps,
                                                                            lv.stream_G, lv.stream_S, lv.stream_D, lv.stream_ring, lv.A.v.recip, h->d_dbg);
        ++g_launches;
#ifdef AMGB200_TIMELINE
        if (h->d_dbg) {
            long long hd[32];
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            CUDA_CHECK(cudaMemcpy(hd, h->d_dbg, sizeof(hd), cudaMemcpyDeviceToHost));
            const char *nm[11] = {"wait block", "products", "prefix fold", "wait done(g-1)", "patch late", "suffix fold", "div+store", "group barrier", "arrive+release", "wait done(g-2)", "patch late2"};
            const int iters = (lv.W * nsweeps + lv.stream_D - 1) / lv.stream_D;
            for (int g2 = 0; g2 < 2; ++g2) {
                printf("   stream timeline group %d warp r=0 (cycles per wavefront, %d wavefronts, G=%d):", g2, iters, lv.stream_G);
                long long tot = 0;
                for (int i = 0; i < 11; ++i) { printf("  %s %lld", nm[i], hd[g2 * 16 + i] / iters); tot += hd[g2 * 16 + i]; }
                printf("  | total %lld (D=%d groups)\n", tot / iters, lv.stream_D);
            }
        }
#endif
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    if (lv.strategy == 2) {
        const int G = lv.cta_G, D = lv.cta_D, nw = G * D;
        const int cap = lv.cta_cap;
        const size_t stage = (size_t)nw * (STAGE + (cap ? cap + 24 + 2 * LATE_CAP : 0)) * sizeof(double);
        const size_t xbytes = (size_t)((lv.n + 1) & ~1) * sizeof(double);
        if (lv.x_in_smem) {
            if (first_use((const void *)gs_ordered_cta_kernel<KIND, EXACT, true>)) {
                CUDA_CHECK(cudaFuncSetAttribute(gs_ordered_cta_kernel<KIND, EXACT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_dyn_smem));
            }
            gs_ordered_cta_kernel<KIND, EXACT, true><<<1, 32 * nw, xbytes + stage, h->stream>>>(lv.A.v, lv.b, lv.x, lv.d_wf_item_ptr, lv.W, nsweeps, G, D, cap, h->d_dbg);
        } else {
            if (first_use((const void *)gs_ordered_cta_kernel<KIND, EXACT, false>)) {
                CUDA_CHECK(cudaFuncSetAttribute(gs_ordered_cta_kernel<KIND, EXACT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_dyn_smem));
            }
            gs_ordered_cta_kernel<KIND, EXACT, false><<<1, 32 * nw, stage, h->stream>>>(lv.A.v, lv.b, lv.x, lv.d_wf_item_ptr, lv.W, nsweeps, G, D, cap, h->d_dbg);
        }
        ++g_launches;
        if (h->d_dbg) {
            long long hd[16 * 8 + 1];
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            CUDA_CHECK(cudaMemcpy(hd, h->d_dbg, sizeof(hd), cudaMemcpyDeviceToHost));
            printf("[dbg] CTA kernel level n=%d G=%d D=%d W=%d sweeps=%d  warp0 suffix chunks %lld\n", lv.n, G, D, lv.W, nsweeps, hd[16 * 8]);
#ifdef AMGB200_TIMELINE
            {
                const char *nm[14] = {"prefix chain", "stage suffix", "bar.sync(prev wavefront)", "patch late", "suffix chain", "diag+div+store", "fence+arrive", "group barrier", "fetch", "L1 prefetch", "prefix load wait", "suffix begin issue", "suffix load wait", "-"};
                const int iters = (lv.W * nsweeps + 1) / 2;
                for (int g2 = 0; g2 < 2; ++g2) {
                    printf("   timeline group %d warp 0 (cycles per wavefront of this group, %d wavefronts):", g2, iters);
                    long long tot = 0;
                    for (int i = 0; i < 13; ++i) { printf("  %s %lld", nm[i], hd[g2 * 16 + i] / iters); tot += hd[g2 * 16 + i]; }
                    printf("  | total %lld\n", tot / iters);
                }
            }
#endif
            for (int w = 0; w < nw; ++w) printf("   warp %2d: prefix %9lld  wait %9lld  suffix %9lld  post(group barrier+fetch) %9lld  items %6lld | suffix: gather %8lld prod %8lld chain %8lld\n", w, hd[w*8], hd[w*8+1], hd[w*8+2], hd[w*8+3], hd[w*8+4], hd[w*8+5], hd[w*8+6], hd[w*8+7]);
        }
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    if (lv.strategy == 3) {
        if (first_use((const void *)gs_ordered_cluster_kernel<KIND, EXACT>)) {
            CUDA_CHECK(cudaFuncSetAttribute(gs_ordered_cluster_kernel<KIND, EXACT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CLUSTER_CTAS);
        {
            // enough warps in the cluster to take the widest wavefront in one round, and no more (barrier cost grows with threads)
            const int maxw = KIND == 0 ? std::min(h->cluster_block / 32, CLUSTER_WARPS_SELL) : CLUSTER_WARPS_CSR;
            const int nwc = std::max(1, std::min(maxw, (lv.max_width + CLUSTER_CTAS - 1) / CLUSTER_CTAS));
            cfg.blockDim = dim3(32 * nwc);
        }
        cfg.dynamicSmemBytes = 0;
        cfg.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CLUSTER_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (KIND == 0 && lv.A.v.max_row <= 20) {
            if (first_use((const void *)gs_ordered_cluster_kernel<KIND, EXACT, true>)) {
                CUDA_CHECK(cudaFuncSetAttribute(gs_ordered_cluster_kernel<KIND, EXACT, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            }
            CUDA_CHECK(cudaLaunchKernelEx(&cfg, gs_ordered_cluster_kernel<KIND, EXACT, true>, lv.A.v, (const double *)lv.b, lv.x, (const int *)lv.d_wf_item_ptr, lv.W, nsweeps, h->d_dbg));
        } else if (KIND == 0 && lv.A.v.max_row <= 28) {
            auto k28 = &gs_ordered_cluster_kernel<KIND, EXACT, true, 28>;
            if (first_use((const void *)k28)) {
                CUDA_CHECK(cudaFuncSetAttribute(k28, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            }
            CUDA_CHECK(cudaLaunchKernelEx(&cfg, k28, lv.A.v, (const double *)lv.b, lv.x, (const int *)lv.d_wf_item_ptr, lv.W, nsweeps, h->d_dbg));
        } else
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, gs_ordered_cluster_kernel<KIND, EXACT>, lv.A.v, (const double *)lv.b, lv.x, (const int *)lv.d_wf_item_ptr, lv.W, nsweeps, h->d_dbg));
        ++g_launches;
        if (h->d_dbg) {
            long long hd[5 * 8];
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            CUDA_CHECK(cudaMemcpy(hd, h->d_dbg, sizeof(hd), cudaMemcpyDeviceToHost));
            printf("[dbg] cluster kernel n=%d W=%d sweeps=%d\n", lv.n, lv.W, nsweeps);
            for (int w = 0; w < 5; ++w) printf("   warp slot %d: finish(first) %9lld  finish(extra rounds) %9lld  arrive %9lld  prefetch %9lld  wait %9lld  items %6lld\n", w, hd[w*8], hd[w*8+1], hd[w*8+2], hd[w*8+3], hd[w*8+4], hd[w*8+5]);
        }
        return;
    }
    fprintf(stderr, "libamgb200: internal error: no smoother kernel for strategy %d\n", lv.strategy);
    exit(71);
}

void smooth_level(amgb200_hier *h, Level &lv, int nsweeps) {
    if (lv.A.v.kind == KIND_SELL) smooth_k<0, true>(h, lv, nsweeps);
    else if (h->exact) smooth_k<1, true>(h, lv, nsweeps);
    else smooth_k<1, false>(h, lv, nsweeps);
}
// pre-smoothing (post = false) or post-smoothing.  With C/F ordering both run the same F-then-C sweeps
// (SSS_smooth.c:16: only `if (order)` is tested); in natural order the post-smoother runs backwards (:122-134).
void smooth(amgb200_hier *h, int l, int nsweeps, bool post = false) {
    Level &lv = h->L[l];
    if (nsweeps <= 0) return;
    if (lv.natural && post) {
        Level &bk = *lv.bk;
        LAUNCH(gather_idx_kernel, grid_for(lv.n), BLOCK, h->stream, lv.n, lv.d_fb, lv.x, bk.x);
        LAUNCH(gather_idx_kernel, grid_for(lv.n), BLOCK, h->stream, lv.n, lv.d_fb, lv.b, bk.b);
        smooth_level(h, bk, nsweeps);
        LAUNCH(scatter_idx_kernel, grid_for(lv.n), BLOCK, h->stream, lv.n, lv.d_fb, bk.x, lv.x);
        return;
    }
    smooth_level(h, lv, nsweeps);
}

// ---- coarsest-level Krylov solvers (host control flow, device vectors) ---------------------
constexpr double BIGF = 1e+20, SMALLF = 1e-20, SMALLF2 = 1e-40;
constexpr int MAX_STAG = 20, MAX_RESTART = 30;
constexpr int ERR_STAG = -42, ERR_SOLSTAG = -43, ERR_TOLSMALL = -44, ERR_MAXIT = -48;

// (x, y) feeding a Krylov coefficient: summed left to right like SSS_blas_array_dot (EXACT) or by a tree (FAST)
void launch_dot(amgb200_hier *h, int n, const double *x, const double *y, double *d_out) {
    LAUNCH(dot_seq_kernel, 1, 1024, h->stream, n, x, y, d_out, h->exact ? 1 : 0);
}
double dev_dot(amgb200_hier *h, int n, const double *x, const double *y) {
    launch_dot(h, n, x, y, h->d_scal);
    fetch_scalars(h, 1);
    return h->h_scal[0];
}
double dev_norm2(amgb200_hier *h, int n, const double *x) { return sqrt(dev_dot(h, n, x, x)); }
// tree-reduced ||x||_2 for quantities that never feed back into x (printing, stopping)
double dev_norm2_tree(amgb200_hier *h, int n, const double *x) {
    const int g = red_grid(h, n);
    LAUNCH(dot_tree_kernel, g, BLOCK, h->stream, n, x, x, h->d_partial);
    LAUNCH(reduce_partials_kernel, 1, BLOCK, h->stream, h->d_partial, g, h->partial_stride, 1, 0, h->d_scal);
    fetch_scalars(h, 1);
    return sqrt(h->h_scal[0]);
}
void dev_copy(amgb200_hier *h, int n, const double *src, double *dst) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
}
void dev_zero(amgb200_hier *h, int n, double *x) { CUDA_CHECK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), h->stream)); }
void dev_axpy(amgb200_hier *h, int n, double a, const double *x, double *y) { LAUNCH(axpy_kernel, std::min(grid_for(n), 1184), BLOCK, h->stream, n, a, x, y); }
void dev_scale(amgb200_hier *h, int n, double a, double *x) { LAUNCH(scale_kernel, std::min(grid_for(n), 1184), BLOCK, h->stream, n, a, x); }
// outer loop: r = b - A u, returns ||r||_2 (tree-reduced: it is only printed and compared with tol)
double dev_true_residual(amgb200_hier *h, const DMat &A, const double *u, const double *b, double *r) {
    spmv(h, A, MODE_RESID, RED_SUMSQ, u, r, b, -1.0);
    fetch_scalars(h, 1);
    h->last_sumsq = h->h_scal[0];
    return sqrt(h->h_scal[0]);
}
// Krylov solvers: r = b - A u and (r, r) summed in index order, because it becomes temp1/temp2 of CG
double krylov_residual(amgb200_hier *h, const DMat &A, const double *u, const double *b, double *r) {
    spmv(h, A, MODE_RESID, RED_NONE, u, r, b, -1.0);
    launch_dot(h, A.nrows, r, r, h->d_scal);
    fetch_scalars(h, 1);
    h->last_sumsq = h->h_scal[0];
    return sqrt(h->h_scal[0]);
}

void ensure_krylov(amgb200_hier *h, size_t len) {
    if (h->kry_len >= len) return;
    if (h->kry) dev_free(h->kry);
    h->kry = dev_alloc<double>(len);
    h->kry_len = len;
}

// CG with the reference's safeguards: amg/Solve/SSS_cycle.cu:15-437, stop_type = STOP_REL_RES.
int coarse_cg(amgb200_hier *h, const DMat &A, const double *b, double *u, double tol, int maxit) {
    const int m = A.nrows;
    const double maxdiff = tol * 1e-4;
    int iter = 0, stag = 1, more_step = 1, iter_best = 0;
    double absres0, absres = BIGF, relres, normu, normr0, absres_best = BIGF;
    double alpha, beta, temp1, temp2, reldiff, infnormu;
    ensure_krylov(h, (size_t)6 * m);
    double *p = h->kry, *r = p + m, *t = r + m, *u_best = t + m;   // z == r (no preconditioner): not materialised
    const int ug = red_grid(h, m);

    // the whole iteration in one cooperative launch (coarse_cg_kernel); the rarely taken safeguard branches make it stop with
    // CG_FALLBACK, and the host-driven loop below redoes the solve from the saved initial guess
    if (h->exact && h->opt.coarse_mode == AMGB200_BETA_FIX && h->pcg_ctas > 0) {
        double *u0 = h->kry + (size_t)5 * m;
        dev_copy(h, m, u, u0);
        CgArgs a;
        a.A = A; a.b = b; a.u = u; a.p = p; a.r = r; a.t = t; a.u_best = u_best; a.tol = tol; a.maxit = maxit; a.beta_fix = 1;
        a.scal = h->d_scal; a.partial = h->d_partial; a.status = reinterpret_cast<int *>(h->d_scal + 7);
        a.barrier = reinterpret_cast<unsigned *>(h->d_scal + 6); a.barrier_gen = 0;
        CUDA_CHECK(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), h->stream));
        const void *kern = A.kind == KIND_SELL ? (const void *)coarse_cg_kernel<0> : (const void *)coarse_cg_kernel<1>;
        // (a few dozen rows -- the coarsest matrices of the 2D and the anisotropic problems: ONE CTA, whose barrier is __syncthreads)
        const int grid = A.nitems <= 128 ? 1 : std::max(1, std::min(h->pcg_ctas, (A.nitems + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK));
        void *args[] = {&a};
        CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(BLOCK), args, 0, h->stream));
        ++g_launches;
        fetch_scalars(h, 8);
        int status;
        memcpy(&status, h->h_scal + 7, sizeof(int));
        if (status != CG_FALLBACK) return status;
        dev_copy(h, m, u0, u);
    }

    absres0 = krylov_residual(h, A, u, b, r);
    normr0 = std::max(SMALLF, absres0);
    relres = absres0 / normr0;
    if (relres < tol) goto done;
    dev_copy(h, m, r, p);
    temp1 = h->last_sumsq;                                          // (z, r) with z = r

    while (iter++ < maxit) {
        // one read-back per iteration: (t, p) stays on the device, the update kernel forms alpha from it (and does nothing when
        // |(t, p)| <= 1e-40, in which case the host leaves the loop below exactly like SSS_cycle.cu:190)
        spmv(h, A, MODE_MXY, RED_NONE, p, t, nullptr, 0.0);          // t = A p
        launch_dot(h, m, t, p, h->d_scal + 4);                       // (t, p)
        LAUNCH(cg_update_kernel, ug, BLOCK, h->stream, m, temp1, (const double *)(h->d_scal + 4), p, t, u, r, h->d_partial, h->partial_stride);
        LAUNCH(reduce_partials_kernel, 1, BLOCK, h->stream, h->d_partial, ug, h->partial_stride, 2, 1, h->d_scal + 1);
        launch_dot(h, m, r, r, h->d_scal);                           // (r, r): absres^2 and the next (z, r)
        fetch_scalars(h, 5);
        temp2 = h->h_scal[4];
        if (fabs(temp2) > SMALLF2) alpha = temp1 / temp2;
        else goto restore;
        const double rr = h->h_scal[0], uu = h->h_scal[1], pp = h->h_scal[2];
        infnormu = h->h_scal[3];
        absres = sqrt(rr);
        relres = absres / normr0;
        if (absres < absres_best - maxdiff) {
            absres_best = absres;
            iter_best = iter;
            dev_copy(h, m, u, u_best);
        }
        if (infnormu <= SMALLF) { iter = ERR_SOLSTAG; break; }
        normu = sqrt(uu);
        reldiff = fabs(alpha) * sqrt(pp) / normu;
        bool fresh_r = false;
        if ((stag <= MAX_STAG) & (reldiff < maxdiff)) {
            absres = krylov_residual(h, A, u, b, r);
            relres = absres / normr0;
            fresh_r = true;
            if (relres < tol) break;
            if (stag >= MAX_STAG) { iter = ERR_STAG; break; }
            dev_zero(h, m, p);
            ++stag;
        }
        if (relres < tol) {
            absres = krylov_residual(h, A, u, b, r);
            relres = absres / normr0;
            fresh_r = true;
            if (relres < tol) break;
            if (more_step >= MAX_RESTART) { iter = ERR_TOLSMALL; break; }
            dev_zero(h, m, p);
            ++more_step;
        }
        absres0 = absres;
        temp2 = fresh_r ? h->last_sumsq : rr;                        // (z, r) with z = r
        if (h->opt.coarse_mode == AMGB200_BETA_FIX) { beta = temp2 / temp1; temp1 = temp2; }
        else beta = temp1 / temp1;
        LAUNCH(axpby_kernel, std::min(grid_for(m), 1184), BLOCK, h->stream, m, 1.0, r, beta, p);
    }
restore:
    if (iter != iter_best) {
        absres_best = krylov_residual(h, A, u_best, b, r);
        if (absres > absres_best + maxdiff) dev_copy(h, m, u_best, u);
    }
done:
    (void)absres0;
    return iter > maxit ? ERR_MAXIT : iter;
}

// GMRES(restart) as in amg/Solve/SSS_cycle.cu:440-817 (stop_type = STOP_REL_RES).  Reached only
// when CG fails (never on the BASELINE matrices in FIX mode; every cycle in AS_COMPILED mode).
int coarse_gmres(amgb200_hier *h, const DMat &A, const double *b, double *x, double tol, int maxit, int restart) {
    const int n = A.nrows;
    const double maxdiff = tol * 1e-4;
    const int restart1 = restart + 1;
    int iter = 0, iter_best = 0, i, j, k;
    double r_norm, gamma, t, normr0, absres = BIGF, relres, absres_best = BIGF;
    ensure_krylov(h, (size_t)(restart1 + 3) * n + (size_t)5 * n);
    double *base = h->kry + (size_t)5 * n;               // keep clear of CG's vectors
    double *r = base, *w = r + n, *x_best = w + n;
    std::vector<double *> p(restart1);
    for (i = 0; i < restart1; ++i) p[i] = x_best + n + (size_t)i * n;
    std::vector<double> rs(restart1, 0.0), c(restart, 0.0), s(restart, 0.0);
    std::vector<std::vector<double>> hh(restart1, std::vector<double>(restart, 0.0));

    r_norm = krylov_residual(h, A, x, b, p[0]);
    normr0 = std::max(SMALLF, r_norm);
    relres = r_norm / normr0;
    if (relres < tol) goto done;

    while (iter < maxit) {
        rs[0] = r_norm;
        t = 1.0 / r_norm;
        dev_scale(h, n, t, p[0]);
        i = 0;
        while (i < restart && iter < maxit) {
            i++; iter++;
            spmv(h, A, MODE_MXY, RED_NONE, p[i - 1], p[i], nullptr, 0.0);
            for (j = 0; j < i; j++) {
                hh[j][i - 1] = dev_dot(h, n, p[j], p[i]);
                dev_axpy(h, n, -hh[j][i - 1], p[j], p[i]);
            }
            t = dev_norm2(h, n, p[i]);
            hh[i][i - 1] = t;
            if (t != 0.0) { t = 1.0 / t; dev_scale(h, n, t, p[i]); }
            for (j = 1; j < i; ++j) {
                t = hh[j - 1][i - 1];
                hh[j - 1][i - 1] = s[j - 1] * hh[j][i - 1] + c[j - 1] * t;
                hh[j][i - 1] = -s[j - 1] * t + c[j - 1] * hh[j][i - 1];
            }
            t = hh[i][i - 1] * hh[i][i - 1];
            t += hh[i - 1][i - 1] * hh[i - 1][i - 1];
            gamma = sqrt(t);
            if (gamma == 0.0) gamma = SMALLF;
            c[i - 1] = hh[i - 1][i - 1] / gamma;
            s[i - 1] = hh[i][i - 1] / gamma;
            rs[i] = -s[i - 1] * rs[i - 1];
            rs[i - 1] = c[i - 1] * rs[i - 1];
            hh[i - 1][i - 1] = s[i - 1] * hh[i][i - 1] + c[i - 1] * hh[i - 1][i - 1];
            absres = r_norm = fabs(rs[i]);
            relres = absres / normr0;
            if (relres <= tol) break;
        }
        rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];
        for (k = i - 2; k >= 0; k--) {
            t = 0.0;
            for (j = k + 1; j < i; j++) t -= hh[k][j] * rs[j];
            t += rs[k];
            rs[k] = t / hh[k][k];
        }
        dev_copy(h, n, p[i - 1], w);
        dev_scale(h, n, rs[i - 1], w);
        for (j = i - 2; j >= 0; j--) dev_axpy(h, n, rs[j], p[j], w);
        dev_axpy(h, n, 1.0, w, x);
        if (absres < absres_best - maxdiff) {
            absres_best = absres;
            iter_best = iter;
            dev_copy(h, n, x, x_best);
        }
        if (relres <= tol) {
            r_norm = krylov_residual(h, A, x, b, r);
            absres = r_norm;
            relres = absres / normr0;
            if (relres <= tol) break;
            dev_copy(h, n, r, p[0]);
            i = 0;
        }
        for (j = i; j > 0; j--) {
            rs[j - 1] = -s[j - 1] * rs[j];
            rs[j] = c[j - 1] * rs[j];
        }
        if (i) dev_axpy(h, n, rs[i] - 1.0, p[i], p[i]);
        for (j = i - 1; j > 0; j--) dev_axpy(h, n, rs[j], p[j], p[i]);
        if (i) {
            dev_axpy(h, n, rs[0] - 1.0, p[0], p[0]);
            dev_axpy(h, n, 1.0, p[i], p[0]);
        }
    }
    if (iter != iter_best) {
        absres_best = krylov_residual(h, A, x_best, b, r);
        if (absres > absres_best + maxdiff) dev_copy(h, n, x_best, x);
    }
done:
    return iter >= maxit ? ERR_MAXIT : iter;
}

// amg/Solve/SSS_cycle.cu:819-846
int coarse_solve(amgb200_hier *h, const DMat &A, const double *b, double *x, double tol, int its[2]) {
    const long long nn = (long long)A.nrows * A.nrows;
    const int maxit = (int)std::max<long long>(250, std::min<long long>(nn, 1000));
    int status = coarse_cg(h, A, b, x, tol, maxit);
    if (its) { its[0] = status; its[1] = 0; }
    if (status < 0) {
        status = coarse_gmres(h, A, b, x, tol, maxit, MAX_RESTART);
        if (its) its[1] = status;
    }
    if (status < 0) printf("### WARNING: Coarse level solver failed to converge!\n");
    return status;
}

// ---- V/W-cycle: amg/Solve/SSS_cycle.cu:848-967 ---------------------------------------------
void cycle_from(amgb200_hier *h, int lstart) {
    const int nl = h->nl;
    int cycle_type = h->pars.cycle_type;
    double tol = h->pars.ctol;
    int visits[64] = {0}, l = lstart;
    if (tol > h->pars.tol) tol = h->pars.tol * 0.1;
    if (cycle_type <= 0) cycle_type = 1;
    for (;;) {
        while (l < nl - 1) {
            Level &lv = h->L[l];
            visits[l]++;
            {
                PhaseTimer pt(h, 0, l);
                if (lv.lo && lv.x_is_zero && h->pars.pre_iter >= 1) {        // first sweep from x = 0: lower triangle only (bit-identical, see upload)
                    smooth_level(h, *lv.lo, 1);
                    smooth(h, l, h->pars.pre_iter - 1);
                } else smooth(h, l, h->pars.pre_iter);
                lv.x_is_zero = false;
            }
            if (lv.rr.valid) {
                PhaseTimer pt(h, 1, l);
                resid_restrict(h, l, true);
                l++;
            } else {
                { PhaseTimer pt(h, 1, l); spmv(h, lv.spmvA(), MODE_RESID, RED_NONE, lv.x, lv.wp, lv.b, -1.0); }
                { PhaseTimer pt(h, 2, l); spmv(h, lv.R.v, MODE_MXY, RED_NONE, lv.wp, h->L[l + 1].b, nullptr, 0.0); }
                l++;
                dev_zero(h, h->L[l].n, h->L[l].x);
            }
            h->L[l].x_is_zero = true;
        }
        { PhaseTimer pt(h, 4); coarse_solve(h, h->L[nl - 1].A.v, h->L[nl - 1].b, h->L[nl - 1].x, tol, nullptr); }
        while (l > lstart) {
            l--;
            Level &lv = h->L[l];
            { PhaseTimer pt(h, 3, l); spmv(h, lv.P.v, MODE_AMXPY, RED_NONE, h->L[l + 1].x, lv.x, nullptr, 1.0); }
            { PhaseTimer pt(h, 0, l); smooth(h, l, h->pars.post_iter, true); }
            if (visits[l] < cycle_type) break;
            visits[l] = 0;
            if (l == lstart) break;
        }
        if (l <= lstart) break;
    }
}
void cycle(amgb200_hier *h) { cycle_from(h, 0); }

void to_schedule(amgb200_hier *h, int l, const double *d_nat, double *d_sched) {
    LAUNCH(gather_kernel, grid_for(h->L[l].n), BLOCK, h->stream, h->L[l].n, h->L[l].d_order, d_nat, d_sched);
}
void to_natural(amgb200_hier *h, int l, const double *d_sched, double *d_nat) {
    LAUNCH(scatter_kernel, grid_for(h->L[l].n), BLOCK, h->stream, h->L[l].n, h->L[l].d_order, d_sched, d_nat);
}

void print_itinfo(int iter, double relres, double absres, double factor) {   // amg/SSS_utils.c:104-133
    if (iter > 0) {
        printf("%6d | %13.6e   | %13.6e  | %10.4lf\n", iter, relres, absres, factor);
    } else {
        printf("-----------------------------------------------------------\n");
        printf("It Num |   ||r||/||b||   |     ||r||      |  Conv. Factor\n");
        printf("-----------------------------------------------------------\n");
        printf("%6d | %13.6e   | %13.6e  |     -.-- \n", iter, relres, absres);
    }
}

void check_level(const amgb200_hier *h, int level) {
    if (level < 0 || level >= h->nl) { fprintf(stderr, "libamgb200: level %d out of range [0,%d)\n", level, h->nl); exit(72); }
}

}  // namespace

// =============================================================================================
// public API
// =============================================================================================
extern "C" {

void amgb200_default_options(amgb200_options *o) {
    memset(o, 0, sizeof(*o));
    o->coarse_mode = AMGB200_BETA_FIX;
    o->verbose = 0;
    o->device = -1;
    o->fast = 0;
    if (getenv("AMGB200_FAST") && atoi(getenv("AMGB200_FAST")) > 0) o->fast = 1;
    const char *e = getenv("AMGB200_COARSE_MODE");
    if (e && (!strcmp(e, "asc") || !strcmp(e, "as_compiled") || !strcmp(e, "1"))) o->coarse_mode = AMGB200_BETA_AS_COMPILED;
}

const char *amgb200_version(void) { return "amg-b200 0.1 (sm_100a)"; }
long long amgb200_launch_count(void) { return g_launches; }

amgb200_hier *amgb200_upload(const amgb200_amg *mg, const amgb200_options *opt_in) {
    amgb200_options opt;
    if (opt_in) opt = *opt_in; else amgb200_default_options(&opt);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        fprintf(stderr, "libamgb200: no usable CUDA device (%s); this library has no CPU fallback\n", cudaGetErrorString(e));
        exit(70);
    }
    if (opt.device >= 0) CUDA_CHECK(cudaSetDevice(opt.device));
    if (mg->pars.smoother != 2) {
        printf("### ERROR: Wrong smoother type %d!\n", mg->pars.smoother);   // SSS_smooth.c:216-218
        exit(-12);
    }
    if (mg->num_levels < 1 || mg->num_levels > 64) { fprintf(stderr, "libamgb200: %d levels (supported: 1..64)\n", mg->num_levels); exit(72); }
    amgb200_hier *h = new amgb200_hier();
    // level-0 worker of the sharded solve: levels 0 and 1 only, and of level 1 nothing but its vectors in its REAL schedule numbering
    // (P_0's columns and the coarse correction that rank 0 stores into x_1 use it)
    const bool worker = opt.level0_worker && mg->num_levels >= 2;
    h->nl = worker ? 2 : mg->num_levels;
    h->pars = mg->pars;
    h->opt = opt;
    h->profile = getenv("AMGB200_PROFILE") && atoi(getenv("AMGB200_PROFILE")) > 0;
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreate(&h->ev0));
    CUDA_CHECK(cudaEventCreate(&h->ev1));
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    {
        // (three attribute queries instead of cudaGetDeviceProperties, which costs milliseconds on every upload)
        int sms = 0, smem_optin = 0, coop = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        h->num_sms = sms;
        h->max_dyn_smem = smem_optin - 1024;
        if (!coop) { fprintf(stderr, "libamgb200: device lacks cooperative launch\n"); exit(70); }
    }
    h->exact = !opt.fast;
    {
        int per_sm = 0;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coarse_cg_kernel<1>, BLOCK, 0));
        h->pcg_ctas = std::min(per_sm, 2) * h->num_sms;
        if (getenv("AMGB200_PCG_CTAS")) h->pcg_ctas = std::max(0, std::min(per_sm * h->num_sms, atoi(getenv("AMGB200_PCG_CTAS"))));
    }
    const double sell_max_mean = getenv("AMGB200_SELL_MAX_MEAN") ? atof(getenv("AMGB200_SELL_MAX_MEAN")) : (h->exact ? 48.0 : 24.0);
    const double cta_max_avg = getenv("AMGB200_CTA_MAX_AVG") ? atof(getenv("AMGB200_CTA_MAX_AVG")) : 6.0;
    const double stream_max_avg = getenv("AMGB200_STREAM_MAX_AVG") ? atof(getenv("AMGB200_STREAM_MAX_AVG")) : 12.0;
    if (getenv("AMGB200_CLUSTER_BLOCK")) h->cluster_block = std::max(32, std::min(BLOCK, atoi(getenv("AMGB200_CLUSTER_BLOCK")) / 32 * 32));
    auto kind_of = [&](const amgb200_mat &M) { return choose_kind(M, sell_max_mean); };
    const double ordered_csr_min = getenv("AMGB200_ORDERED_CSR_MIN") ? atof(getenv("AMGB200_ORDERED_CSR_MIN")) : 24.0;
    const bool no_df = getenv("AMGB200_NO_DF") && atoi(getenv("AMGB200_NO_DF"));
    const bool df_all = getenv("AMGB200_DF_ALL") && atoi(getenv("AMGB200_DF_ALL"));       // also SELL levels the streaming kernels took
    if (getenv("AMGB200_DF_AHEAD")) h->df_ahead = std::max(1, atoi(getenv("AMGB200_DF_AHEAD")));
    const double dfw_min_width = getenv("AMGB200_DFW_MIN_WIDTH") ? atof(getenv("AMGB200_DFW_MIN_WIDTH")) : 150.0;
    const bool use_lower = !(getenv("AMGB200_NO_LOWER") && atoi(getenv("AMGB200_NO_LOWER")));

    const double t0 = now_s();
    const int nl = h->nl;
    h->L.resize(nl);
    const bool dev_fill = !(getenv("AMGB200_HOST_LAYOUT") && atoi(getenv("AMGB200_HOST_LAYOUT")));
    RawPrefetch prefetch;
    if (dev_fill && !(getenv("AMGB200_NO_PREFETCH") && atoi(getenv("AMGB200_NO_PREFETCH")))) {
        std::vector<const amgb200_mat *> mats;                  // in the order the level loop below asks for them
        for (int l = 0; l < nl; ++l) {
            const amgb200_comp &c = mg->cg[l];
            if (kind_of(c.A) == KIND_SELL && !(worker && l == 1)) mats.push_back(&c.A);
            if (l < nl - 1) { if (kind_of(c.P) == KIND_SELL) mats.push_back(&c.P); if (kind_of(c.R) == KIND_SELL && !worker) mats.push_back(&c.R); }
        }
        prefetch.start(mats, dev);
    }
    std::vector<Schedule> sched(nl);
    int maxn = 0;
    // natural-order Gauss-Seidel (SSS_smooth.c:90-137) when cf_order = 0 or a level has no C/F marks (SSS_smooth.c:171-176):
    // forward sweeps before, backward sweeps after the coarse-grid correction -> a second (reversed) schedule per level
    std::vector<Schedule> sched_b(nl);
    std::vector<char> natural(nl, 0);
    for (int l = 0; l + 1 < nl; ++l) natural[l] = !(mg->pars.cf_order && mg->cg[l].cfmark.d);
    omp_set_max_active_levels(2);                                                                   // (build_schedule runs its two passes on two threads)
#pragma omp parallel for schedule(dynamic, 1)
    for (int l = 0; l < nl - 1; ++l) {                                                              // levels are independent
        const amgb200_mat &A = mg->cg[l].A;
        build_schedule(A, natural[l] ? nullptr : mg->cg[l].cfmark.d, sched[l]);
        if (natural[l]) {
            const int n = A.num_rows;                                                               // the matrix seen from the last row backwards
            std::vector<int> rp((size_t)n + 1, 0), ci((size_t)A.num_nnzs);
            for (int i = 0; i < n; ++i) rp[i + 1] = rp[i] + (A.row_ptr[n - i] - A.row_ptr[n - 1 - i]);
            for (int i = 0; i < n; ++i) {
                const int src = A.row_ptr[n - 1 - i];
                for (int q = 0; q < rp[i + 1] - rp[i]; ++q) ci[rp[i] + q] = n - 1 - A.col_idx[src + q];
            }
            amgb200_mat Arev = {n, n, A.num_nnzs, rp.data(), ci.data(), A.val};
            Schedule R;
            build_schedule(Arev, nullptr, R);
            Schedule &B = sched_b[l];
            B = R;
            for (int k = 0; k < n; ++k) B.order[k] = n - 1 - R.order[k];
            for (int i = 0; i < n; ++i) B.pos[i] = R.pos[n - 1 - i];
        }
    }
    for (int l = 0; l < nl; ++l) {
        const amgb200_comp &c = mg->cg[l];
        if (l < nl - 1) {
            if (sched[l].rows_without_diag) {
                fprintf(stderr, "libamgb200: level %d has %d rows without a stored diagonal; the reference's carried-over "
                                "diagonal (SSS_smooth.c:13,30) cannot be reproduced in parallel\n", l, sched[l].rows_without_diag);
                exit(-22);
            }
        } else if (worker && l + 1 < mg->num_levels) {
            build_schedule(c.A, (mg->pars.cf_order && c.cfmark.d) ? c.cfmark.d : nullptr, sched[l]);      // level 1 as rank 0 numbers it
        } else {
            identity_schedule(c.A.num_rows, sched[l]);
        }
        maxn = std::max(maxn, c.A.num_rows);
    }
    h->analysis_s = now_s() - t0;
    DevMatOwner::t_upload() = 0;
    double t_layout = 0;
    std::mutex tl_mutex;
    auto tl_note = [&](const char *what, int l, double dt) {
        std::lock_guard<std::mutex> lk(tl_mutex);
        t_layout += dt;
        if (opt.verbose >= 3) printf("      [layout] level %d %-8s %.1f ms\n", l, what, 1e3 * dt);
    };

    int max_items = 1;
    // wavefront tables and launch strategy of one level's smoother (used for the level itself and, in natural
    // order, for its backward twin)
    auto setup_smoother = [&](Level &lv, const DevLayout &lay, const Schedule &S, const amgb200_mat &Amat) {
    lv.W = S.wf_count[0] + S.wf_count[1];
    lv.wf_count[0] = S.wf_count[0]; lv.wf_count[1] = S.wf_count[1];
    lv.pass_rows[0] = S.pass_rows[0]; lv.pass_rows[1] = S.pass_rows[1];
    lv.ordered = S.wf_count[0] > 1 || S.wf_count[1] > 1;
    const std::vector<int> &wip = lay.wf_item_ptr;
    lv.pass_items[0] = wip[S.wf_count[0]] - wip[0];
    lv.pass_items[1] = wip[lv.W] - wip[S.wf_count[0]];
    std::vector<int> item_wf((size_t)wip[lv.W]);
    lv.max_width = 1;
    for (int w = 0; w < lv.W; ++w) {
        lv.max_width = std::max(lv.max_width, wip[w + 1] - wip[w]);
        for (int it = wip[w]; it < wip[w + 1]; ++it) item_wf[it] = w;
    }
    lv.d_item_wf = dev_upload(item_wf);
    lv.d_wf_item_ptr = dev_upload(wip);
    // the dependency-chain floor of one sweep: per wavefront the longest in-order chain that can only start when the
    // previous wavefront is complete (warp-per-row layouts: the row suffix from its first entry in that wavefront; thread-
    // per-row SELL slices: the whole padded row), one dependent fp64 subtraction (8.1 cycles measured) per term
    lv.chain_terms = 0;
    if (lv.ordered) {
        for (int w = 0; w < lv.W; ++w) {
            int longest = 0;
            for (int it = wip[w]; it < wip[w + 1]; ++it) {
                if (lay.kind == KIND_CSR) longest = std::max(longest, lay.rptr[it + 1] - lay.rptr[it] - (lay.split.empty() ? 0 : lay.split[it]));
                else longest = std::max(longest, (int)((lay.slice_ptr[it + 1] - lay.slice_ptr[it]) / 32));
            }
            lv.chain_terms += longest;
        }
    }
    // 0 parallel passes | 2 one CTA (narrow wavefronts) | 3 one 16-CTA cluster | 4 / 5 streaming CTA / cluster | 6 data-flow grid
    if (!lv.ordered) lv.strategy = 0;
    else lv.strategy = ((double)wip[lv.W] / lv.W <= cta_max_avg) ? 2 : 3;
    {
        const int maxw = lay.kind == KIND_SELL ? CTA_MAX_WARPS_SELL : CTA_MAX_WARPS_CSR;
        const double avg = (double)wip[lv.W] / lv.W;
        (void)avg;
        int G = std::max(1, std::min(lv.max_width, maxw / 2));
        if (getenv("AMGB200_CTA_G")) G = std::max(1, std::min(maxw, atoi(getenv("AMGB200_CTA_G"))));
        G = std::min(G, maxw / 2);                 // the producer/consumer barrier scheme needs >= 2 groups
        lv.cta_G = G;
        lv.cta_D = std::max(2, maxw / G);
        if (getenv("AMGB200_CTA_D")) lv.cta_D = std::max(2, std::min(maxw / G, atoi(getenv("AMGB200_CTA_D"))));
        // the prefix/suffix scheme of the warp-per-row EXACT kernel folds a row's prefix while exactly ONE earlier
        // wavefront is still in flight: two alternating groups
        if (lay.kind == KIND_CSR && h->exact) lv.cta_D = 2;
        const int nwarps = lv.cta_G * lv.cta_D;
        const size_t xb = (size_t)((lv.n + 1) & ~1) * 8, stb = (size_t)nwarps * STAGE * 8;
        lv.x_in_smem = lv.strategy == 2 && xb + stb <= (size_t)h->max_dyn_smem && !(getenv("AMGB200_NO_SMEM_X") && atoi(getenv("AMGB200_NO_SMEM_X")));
        if (lv.strategy == 2 && lay.kind == KIND_CSR && h->exact && !(getenv("AMGB200_NO_PARK") && atoi(getenv("AMGB200_NO_PARK")))) {
            // park the whole suffix of a row (its products) in shared memory before the barrier when it fits
            int cap = std::min(1024, (lay.max_row + 31) / 32 * 32);
            const size_t base = (lv.x_in_smem ? xb : 0) + stb;
            while (cap >= 128 && base + (size_t)nwarps * (cap + 24 + 2 * LATE_CAP) * 8 > (size_t)h->max_dyn_smem) cap -= 128;
            lv.cta_cap = cap >= 128 ? cap : 0;
        }
    }
    if (getenv("AMGB200_GS_STRATEGY") && lv.ordered) lv.strategy = atoi(getenv("AMGB200_GS_STRATEGY")) <= 2 ? 2 : 3;       // (tests: force the CTA / cluster barrier kernels)
    // streaming single-CTA smoother: warp-per-row EXACT levels whose x vector plus a ring of at least two of the
    // largest wavefront blocks fit in shared memory
    if (lv.ordered && lay.kind == KIND_CSR && h->exact && !getenv("AMGB200_GS_STRATEGY") && (double)wip[lv.W] / lv.W <= stream_max_avg) {
        const size_t xb = ((size_t)lv.n * 8 + 127) & ~(size_t)127;
        const long long ring = (long long)h->max_dyn_smem - STREAM_HDR - (long long)xb - 128;
        if (ring >= 4096) {
            StreamLayout SL;
            const double tl = now_s();
            build_stream(lay, SL);
            tl_note("stream", lv.n, now_s() - tl);
            // (a block larger than half the ring is loaded alone -- the loader drains the ring first -- so its wavefront runs without
            // overlap; accepted as long as the MEAN block fits three times: 256^3 level 8, 9 585 rows, 30.7 -> 27.9 ms per sweep against
            // the cluster kernel; AMGB200_STREAM_RELAX=0 restores the strict rule)
            const bool relax = !(getenv("AMGB200_STREAM_RELAX") && !atoi(getenv("AMGB200_STREAM_RELAX")));
            if ((long long)SL.max_block * 2 <= ring || (relax && (long long)SL.max_block <= ring && SL.mean_block * 3 <= ring)) {
                lv.strategy = 4;
                lv.stream_ring = (int)(ring & ~127LL);
                // product warps per group (1, 2 or 4) and row slots of the folding warp (the widest wavefront in one
                // round if possible: unused slots cost nothing)
                // measured (128^3, three wavefronts in flight): with narrow wavefronts the folding warp alone keeps up with the
                // products, and every additional product warp only disturbs the chains of the other wavefronts in flight
                // (level 5: 3.62 -> 3.39 ms per sweep, level 6: 4.50 -> 4.31 with one warp per group instead of four)
                lv.stream_G = (double)wip[lv.W] / lv.W <= 6.0 ? 1 : 4;
                if (getenv("AMGB200_STREAM_G")) { const int g = atoi(getenv("AMGB200_STREAM_G")); lv.stream_G = g >= 4 ? 4 : g >= 2 ? 2 : 1; }
                lv.stream_S = 1;
                while (lv.stream_S < 16 && lv.stream_S < lv.max_width) lv.stream_S *= 2;      // (16 slots + a rare second round beat 32 slots: fewer distinct shared-memory addresses per chain load)
                if (getenv("AMGB200_STREAM_S")) { int v = std::max(1, std::min(32, atoi(getenv("AMGB200_STREAM_S")))); lv.stream_S = 1; while (lv.stream_S < v) lv.stream_S *= 2; }
                // three wavefronts in flight (the late2 lists exist for W >= 4); a ring that cannot hold three of the largest
                // blocks only serialises the affected wavefronts (every block is released in order)
                lv.stream_D = lv.W >= 4 ? 3 : 2;
                if (getenv("AMGB200_STREAM_D")) { const int d = atoi(getenv("AMGB200_STREAM_D")); if (d == 2 || (lv.W >= 4 && d == 3)) lv.stream_D = d; }
                if (lv.stream_D * lv.stream_G > 12) lv.stream_G = 4;
                lv.d_stream = dev_upload(SL.data);
                lv.d_blk_ptr = dev_upload(SL.blk_ptr);
                if (h->opt.verbose >= 2) printf("      streaming CTA smoother: %d groups of %d product warps, %d row slots in the folding warp, ring %d B, wavefront block mean %lld B max %d B, stream %.1f MB\n", lv.stream_D, lv.stream_G, lv.stream_S, lv.stream_ring, SL.mean_block, SL.max_block, SL.data.size() / 1e6);
            }
        }
    }
    // data-flow smoother, warp-per-row form, for the levels with longer rows whose wavefronts are wide enough to feed the whole GPU (the
    // streaming cluster kernel keeps them on 16 SMs): measured threshold AMGB200_DFW_MIN_WIDTH rows per wavefront on average
    if (lv.ordered && h->exact && lay.kind == KIND_CSR && lv.A.valid && lv.strategy != 4 && !getenv("AMGB200_GS_STRATEGY") && !no_df &&
        (double)S.n / std::max(1, lv.W) >= dfw_min_width) {
        const int cap = (lay.max_row + 7) & ~7;
        const size_t smem = (size_t)(DFW_BLOCK / 32) * (cap + 16) * sizeof(double);
        if (smem <= (size_t)h->max_dyn_smem) {
            const double tl = now_s();
            int *d_missing = dev_alloc<int>(1);
            CUDA_CHECK(cudaMemsetAsync(d_missing, 0, sizeof(int), (cudaStream_t)0));
            LAUNCH(dfw_symmetry_kernel, std::max(1, (lv.n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK), BLOCK, (cudaStream_t)0, lv.A.v, d_missing);
            int missing = 0;
            CUDA_CHECK(cudaMemcpy(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost));
            dev_free(d_missing);
            if (missing == 0) {
                lv.strategy = 7;
                lv.dfw_cap = cap;
                lv.d_rec = dev_alloc<XRec>((size_t)lv.n);
                CUDA_CHECK(cudaMemsetAsync(lv.d_rec, 0, (size_t)lv.n * sizeof(XRec), (cudaStream_t)0));
                lv.df_vbase = 0;
                if (first_use((const void *)gs_dataflow_csr_kernel)) CUDA_CHECK(cudaFuncSetAttribute(gs_dataflow_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_dyn_smem));
                int per_sm = 0;
                CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gs_dataflow_csr_kernel, DFW_BLOCK, smem));
                if (getenv("AMGB200_DF_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(getenv("AMGB200_DF_PER_SM"))));
                lv.df_grid = std::max(1, std::min(per_sm * h->num_sms, (lv.n + DFW_BLOCK / 32 - 1) / (DFW_BLOCK / 32)));
                if (lv.d_stream) { dev_free(lv.d_stream); dev_free(lv.d_blk_ptr); lv.d_stream = nullptr; lv.d_blk_ptr = nullptr; }
                if (lv.d_wf_row_ptr) { dev_free(lv.d_wf_row_ptr); lv.d_wf_row_ptr = nullptr; }
                if (h->opt.verbose >= 2) printf("      data-flow smoother (warp per row): %d CTAs x %d threads (%d per SM), staging %d doubles per warp\n", lv.df_grid, DFW_BLOCK, per_sm, cap + 16);
            } else if (h->opt.verbose >= 2) printf("      data-flow smoother not used: %d entries without a mirror entry (non-symmetric pattern)\n", missing);
            tl_note("symmetry", lv.n, now_s() - tl);
        }
    }
    // streaming cluster smoother: the other ordered levels (any layout; the packer works from the host matrix)
    // (not for narrow thread-per-row levels whose x fits in one SM's shared memory: with branch-free chains the single-CTA barrier kernel
    // is faster there -- 2D 256^2 levels 2-5: 8.6 / 4.2 / 2.1 / 1.3 ms per solve against 11.3 / 5.3 / 2.3 / 1.4 on the cluster)
    // (rows of one register chunk only: a second chunk is fetched on the dependency path -- anisotropic 64^3 level 4, 22 entries per row: 998 us
    // per sweep against ~700 on the cluster)
    const bool cta_keeps = lv.strategy == 2 && lay.kind == KIND_SELL && lv.x_in_smem && lay.max_row <= 20 && !(getenv("AMGB200_XC_SMALL") && atoi(getenv("AMGB200_XC_SMALL")));
    if (lv.ordered && h->exact && lv.strategy != 4 && lv.strategy != 7 && !cta_keeps && lv.W >= 4 && !getenv("AMGB200_GS_STRATEGY") && !(getenv("AMGB200_NO_XC") && atoi(getenv("AMGB200_NO_XC")))) {
        ClusterStreamLayout SL;
        const double tl = now_s();
        build_stream_cluster(Amat, S, XC_CTAS, SL, (long long)h->max_dyn_smem - XC_HDR - 128, getenv("AMGB200_XC_LD") ? atoi(getenv("AMGB200_XC_LD")) : 3);
        tl_note("xcluster", lv.n, now_s() - tl);
        const int cap = (SL.max_width + 1) & ~1;        // every CTA holds the whole wavefront (x3)
        const long long avail = (long long)h->max_dyn_smem - XC_HDR - (SL.late_dist + 1LL) * cap * 8 - 128;
        if (SL.filled && (long long)SL.max_block * 2 <= avail) {
            lv.strategy = 5;
            lv.xc_cap = cap;
            lv.xc_NB = SL.late_dist + 1;
            lv.xc_ring = (int)(std::min<long long>(avail, std::max<long long>(8LL * SL.max_block, 65536)) & ~15LL);
            const double mean_local = (double)S.n / lv.W / XC_CTAS;
            lv.xc_S = 1;
            while (lv.xc_S < 32 && lv.xc_S < SL.max_local) lv.xc_S *= 2;
            lv.xc_F = std::max(1, std::min(XC_G, (int)((1.5 * mean_local + 31) / 32)));
            lv.xc_P = (double)Amat.num_nnzs / std::max(1, Amat.num_rows) <= 96.0 ? 8 : 32;      // lanes per row in the product pass
            lv.xc_D = 3;
            if (getenv("AMGB200_XC_D")) lv.xc_D = atoi(getenv("AMGB200_XC_D")) >= 3 ? 3 : 2;
            if (getenv("AMGB200_XC_P")) lv.xc_P = atoi(getenv("AMGB200_XC_P")) >= 32 ? 32 : 8;
            if (getenv("AMGB200_XC_F")) lv.xc_F = std::max(1, std::min(XC_G, atoi(getenv("AMGB200_XC_F"))));
            if (getenv("AMGB200_XC_S")) { int v = std::max(1, std::min(32, atoi(getenv("AMGB200_XC_S")))); lv.xc_S = 1; while (lv.xc_S < v) lv.xc_S *= 2; }
            if (lv.d_stream) { dev_free(lv.d_stream); dev_free(lv.d_blk_ptr); }
            lv.d_stream = dev_upload(SL.data);
            lv.d_blk_ptr = dev_upload(SL.blk_ptr);
            lv.d_wf_row_ptr = dev_upload(S.wf_row_ptr);
            if (h->opt.verbose >= 2) printf("      streaming cluster smoother: %d CTAs, %d folding warps/group x %d row slots, ring %d B, block mean %lld B max %d B, rows per CTA per wavefront <= %d, stream %.1f MB\n",
                                            XC_CTAS, lv.xc_F, lv.xc_S, lv.xc_ring, SL.mean_block, SL.max_block, SL.max_local, SL.data.size() / 1e6);
        }
    }
    // data-flow smoother (whole GPU, no wavefront barriers): thread-per-row levels that the streaming kernels did not take.
    // The matrix of the level is already on the device here (lv.A); its pattern must be structurally symmetric.
    if (lv.ordered && h->exact && lay.kind == KIND_SELL && lv.A.valid && (lv.strategy == 3 || df_all) && !getenv("AMGB200_GS_STRATEGY") && !no_df) {
        const double tl = now_s();
        int *d_row_slice = dev_alloc<int>((size_t)lv.n);
        int *d_missing = dev_alloc<int>(1);
        CUDA_CHECK(cudaMemsetAsync(d_missing, 0, sizeof(int), (cudaStream_t)0));
        const int ns = lay.nitems(), g = std::max(1, (ns + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
        LAUNCH(df_row_slice_kernel, g, BLOCK, (cudaStream_t)0, ns, lv.A.v.slice_row, d_row_slice);
        LAUNCH(df_symmetry_kernel, g, BLOCK, (cudaStream_t)0, lv.A.v, (const int *)d_row_slice, d_missing);
        int missing = 0;
        CUDA_CHECK(cudaMemcpy(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost));
        dev_free(d_row_slice); dev_free(d_missing);
        if (missing == 0) {
            lv.strategy = 6;
            lv.d_rec = dev_alloc<XRec>((size_t)lv.n);
            CUDA_CHECK(cudaMemsetAsync(lv.d_rec, 0, (size_t)lv.n * sizeof(XRec), (cudaStream_t)0));
            lv.df_vbase = 0;
            lv.df_sch = lay.max_row <= 8 ? 8 : lay.max_row <= 20 ? 20 : lay.max_row <= 28 ? 28 : 16;
            const void *kern = df_kernel(lv.df_sch);
            int per_sm = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DF_BLOCK, 0));
            if (getenv("AMGB200_DF_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(getenv("AMGB200_DF_PER_SM"))));
            lv.df_grid = std::max(1, std::min(per_sm * h->num_sms, (ns + DF_BLOCK / 32 - 1) / (DF_BLOCK / 32)));
            if (lv.d_stream) { dev_free(lv.d_stream); dev_free(lv.d_blk_ptr); lv.d_stream = nullptr; lv.d_blk_ptr = nullptr; }
            if (h->opt.verbose >= 2) printf("      data-flow smoother: %d CTAs x %d threads (%d per SM), %d entries per thread in registers\n", lv.df_grid, DF_BLOCK, per_sm, lv.df_sch);
        } else if (h->opt.verbose >= 2) printf("      data-flow smoother not used: %d entries without a mirror entry (non-symmetric pattern)\n", missing);
        tl_note("symmetry", lv.n, now_s() - tl);
    }
    };

    // SELL layouts are permuted and padded on the device from the raw CSR arrays (sell_fill_kernel); the host only builds their
    // O(rows) slice tables.  AMGB200_HOST_LAYOUT=1 restores the host fill (debugging).
    std::vector<void *> temps;
    std::vector<int *> d_pos(nl, nullptr);
    for (int l = 0; l < nl; ++l) {
        h->L[l].d_order = dev_upload(sched[l].order);
        if (dev_fill) { d_pos[l] = dev_upload(sched[l].pos); temps.push_back(d_pos[l]); }
    }
    // ticket list and chunk tables of the fused residual (+) restriction launch (kernels.cuh, resid_restrict_kernel)
    const bool use_fused = !(getenv("AMGB200_NO_FUSED") && atoi(getenv("AMGB200_NO_FUSED")));
    // (a cooperative launch with in-kernel hand-offs loses on small levels: 2D 256^2 level 0, 65 536 rows, 13.6 vs 11.5 us)
    const int rr_min_rows = getenv("AMGB200_RR_MIN_ROWS") ? atoi(getenv("AMGB200_RR_MIN_ROWS")) : 262144;
    const bool rr_all = getenv("AMGB200_RR_ALL") && atoi(getenv("AMGB200_RR_ALL"));
    auto build_fused = [&](Level &lv, const DevLayout &la, const DevLayout &lr, const Schedule &Sf, const Schedule &Sc, const amgb200_mat &Rm) {
        const int nA = la.nitems(), nR = lr.nitems();
        bool f_one_a, f_one_r;
        // chunks: 1/32 of the level, at least 4 ticket blocks each; lag: twice the blocks of A the resident grid holds in flight, in chunks, + 2
        f_one_a = la.max_row <= 8; f_one_r = lr.max_row <= 8;
        int per_sm = 0;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rr_kernel(f_one_a, f_one_r), BLOCK, 0));
        if (getenv("AMGB200_RR_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(getenv("AMGB200_RR_PER_SM"))));
        int nch = std::max(1, std::min(32, nA / (4 * RR_TICKETS)));
        if (getenv("AMGB200_RR_CHUNKS")) nch = std::max(1, std::min(4096, atoi(getenv("AMGB200_RR_CHUNKS"))));
        const int blocks_per_chunk = std::max(1, nA / nch / RR_TICKETS);
        int lag = 2 * ((per_sm * h->num_sms + blocks_per_chunk - 1) / blocks_per_chunk) + 2;   // (measured at 256^3: lag 3 482 us, 5 480, 8 470)
        if (getenv("AMGB200_RR_LAG")) lag = std::max(0, atoi(getenv("AMGB200_RR_LAG")));
        FusedPlan P;
        build_fused_plan(la, lr, Sf, Sc, Rm, nch, lag, RR_TICKETS, P);
        Level::Fused &f = lv.rr;
        f.nblocks = (int)P.block_info.size(); f.nch = P.nch; f.epoch = 0;
        f.one_a = f_one_a; f.one_r = f_one_r;
        f.d_work = dev_upload(P.work); f.d_block_info = dev_upload(P.block_info); f.d_chunk_items = dev_upload(P.chunk_items);
        f.d_cnt = dev_alloc<unsigned>((size_t)P.nch);
        CUDA_CHECK(cudaMemset(f.d_cnt, 0, (size_t)P.nch * sizeof(unsigned)));
        f.grid = std::max(1, std::min(per_sm * h->num_sms, f.nblocks));
        f.valid = true;
        if (h->opt.verbose >= 2) printf("      fused residual+restriction: %d + %d slices in %d chunks (lag %d), %d blocks of %d tickets, %d CTAs (%d per SM)\n", nA, nR, P.nch, lag, f.nblocks, RR_TICKETS, f.grid, per_sm);
    };

    // one matrix: host layout + copy, or slice table + device fill
    auto put_matrix = [&](DevMatOwner &dst, DevLayout &lay, const amgb200_mat &M, const Schedule &rowS, const int *d_row_order, const Schedule *colS,
                          const int *d_col_pos, int kind, const std::vector<int> *breaks, const char *what, int l) {
        const double tl = now_s();
        if (kind == KIND_SELL && dev_fill) {
            build_sell_structure(M, rowS.order.data(), breaks, lay);
            tl_note(what, l, now_s() - tl);
            // (legacy default stream: ordered after the cudaMemcpy's of small pageable arrays, whose DMA may still be in flight
            // when the call returns; the library's own stream is non-blocking and would not wait for them)
            dst.upload_sell(lay, M, d_row_order, d_col_pos, (cudaStream_t)0, temps, &prefetch);
        } else {
            build_layout(M, rowS.order.data(), colS ? colS->pos.data() : nullptr, kind, breaks, lay);
            tl_note(what, l, now_s() - tl);
            dst.upload(lay);
        }
    };

    std::vector<std::thread> lower_threads;
    const bool lower_async = !(getenv("AMGB200_LOWER_SYNC") && atoi(getenv("AMGB200_LOWER_SYNC")));
    for (int l = 0; l < nl; ++l) {
        const amgb200_comp &c = mg->cg[l];
        Level &lv = h->L[l];
        const Schedule &S = sched[l];
        lv.n = c.A.num_rows;
        lv.smoothed = l < nl - 1;
        DevLayout lay;
        const bool is_ordered = l < nl - 1 && (S.wf_count[0] > 1 || S.wf_count[1] > 1);
        // ordered (latency-bound) sweeps: a warp per row keeps the scattered x gathers of one dependency step
        // at ~len L1 wavefronts instead of 32*len for a 32-row slice, so rows longer than ordered_csr_min
        // use the warp-per-row layout there; SpMV/residual (throughput-bound) keep the coalesced SELL layout
        int gs_kind = kind_of(c.A);
        const double mean_len = (double)c.A.num_nnzs / std::max(1, c.A.num_rows);
        if (is_ordered && h->exact && mean_len > ordered_csr_min) {
            // ... unless every row fits the single-chunk thread-per-row kernel (<= 28 entries in registers: 27-point operators) and
            // the level is too large for the single-CTA streaming smoother anyway (measured on the 27-point 64^3 level 0: 2.78 ->
            // 2.34 ms per sweep against the warp-per-row cluster kernel)
            int max_len = 0;
#pragma omp parallel for reduction(max : max_len) schedule(static)
            for (int i = 0; i < c.A.num_rows; ++i) max_len = std::max(max_len, c.A.row_ptr[i + 1] - c.A.row_ptr[i]);
            const bool one_chunk = max_len <= 28 && (size_t)c.A.num_rows * 8 > (size_t)h->max_dyn_smem && kind_of(c.A) == KIND_SELL;
            if (!one_chunk) gs_kind = KIND_CSR;
        }
        const bool vectors_only = worker && l == 1;            // (level-0 worker: level 1 only lends its numbering and its vectors)
        if (!vectors_only) {
            put_matrix(lv.A, lay, c.A, S, lv.d_order, &S, d_pos[l], gs_kind, lv.smoothed ? &S.wf_row_ptr : nullptr, "A", l);
            max_items = std::max(max_items, lay.nitems());
        }
        DevLayout lsp_keep;
        if (gs_kind != kind_of(c.A) && !vectors_only) {
            DevLayout &lsp = lsp_keep;
            put_matrix(lv.Asp, lsp, c.A, S, lv.d_order, &S, d_pos[l], kind_of(c.A), nullptr, "A(spmv)", l);
            max_items = std::max(max_items, lsp.nitems());
        }
        lv.x = dev_alloc<double>(lv.n);
        lv.b = dev_alloc<double>(lv.n + 2);               // (+2: the streaming smoother fetches 16-byte aligned segments of b)
        lv.wp = dev_alloc<double>(lv.n);
        CUDA_CHECK(cudaMemset(lv.x, 0, (size_t)lv.n * sizeof(double)));
        CUDA_CHECK(cudaMemset(lv.b, 0, (size_t)lv.n * sizeof(double)));
        lv.pattern_symmetric = S.pattern_symmetric;
        if (lv.smoothed) {
            setup_smoother(lv, lay, S, c.A);
            // Pre-smoothing on a coarse level starts from x = 0 (SSS_cycle.cu:929).  In its first sweep every column that has not been
            // updated yet contributes a*0 = +-0 and t - (+-0) = t exactly (t is never -0: b comes out of sums that start at +0.0, and a
            // difference that cancels is +0), so those entries can leave the in-order chain: the first sweep is a forward substitution
            // with the lower triangle of the schedule numbering.  Built for the chain-bound levels (streaming kernels) only.
            if (l >= 1 && !natural[l] && use_lower && (lv.strategy == 4 || lv.strategy == 5)) {
              // (on a helper thread per level, next to the layouts of the following levels: host packing + its own uploads)
              auto build_lower = [&, l]() {
                CUDA_CHECK(cudaSetDevice(dev));
                const amgb200_comp &c = mg->cg[l];
                Level &lv = h->L[l];
                const Schedule &S = sched[l];
                const double tl = now_s();
                const int n = c.A.num_rows;
                std::vector<int> rp((size_t)n + 1, 0);
#pragma omp parallel for schedule(static)
                for (int i = 0; i < n; ++i) {
                    int cnt = 0;
                    const int pi = S.pos[i];
                    for (int q = c.A.row_ptr[i]; q < c.A.row_ptr[i + 1]; ++q) cnt += S.pos[c.A.col_idx[q]] <= pi;
                    rp[i + 1] = cnt;
                }
                for (int i = 0; i < n; ++i) rp[i + 1] += rp[i];
                std::vector<int> ci((size_t)rp[n]);
                std::vector<double> va((size_t)rp[n]);
#pragma omp parallel for schedule(static)
                for (int i = 0; i < n; ++i) {
                    int w = rp[i];
                    const int pi = S.pos[i];
                    for (int q = c.A.row_ptr[i]; q < c.A.row_ptr[i + 1]; ++q)
                        if (S.pos[c.A.col_idx[q]] <= pi) { ci[w] = c.A.col_idx[q]; va[w] = c.A.val[q]; ++w; }
                }
                amgb200_mat Alow = {n, n, rp[n], rp.data(), ci.data(), va.data()};
                Level *lo = new Level();
                lo->n = lv.n; lo->smoothed = true;
                lo->x = lv.x; lo->b = lv.b;                                  // shared, not owned
                DevLayout llo;
                put_matrix(lo->A, llo, Alow, S, lv.d_order, &S, d_pos[l], KIND_CSR, &S.wf_row_ptr, "A(lower)", l);
                setup_smoother(*lo, llo, S, Alow);
                if (lo->strategy == 4 || lo->strategy == 5) lv.lo = lo;
                else {                                                       // (would fall back to a barrier kernel: not worth it)
                    lo->A.release();
                    dev_free(lo->d_item_wf); dev_free(lo->d_wf_item_ptr); dev_free(lo->d_stream); dev_free(lo->d_blk_ptr); dev_free(lo->d_wf_row_ptr);
                    dev_free(lo->d_rec); dev_free(lo->d_hint);
                    delete lo;
                }
                tl_note("lower", l, now_s() - tl);
              };
              if (lower_async) lower_threads.emplace_back(build_lower); else build_lower();
            }
            if (natural[l]) {
                // backward sweeps: own schedule, layout and vectors; d_fb maps its numbering into this level's
                lv.natural = true;
                lv.A.v.recip = 1;
                const Schedule &B = sched_b[l];
                Level *bk = new Level();
                lv.bk = bk;
                bk->n = lv.n; bk->smoothed = true; bk->natural = true;
                DevLayout lb;
                {
                    int *d_border = nullptr, *d_bpos = nullptr;
                    if (gs_kind == KIND_SELL && dev_fill) {
                        d_border = dev_upload(B.order); d_bpos = dev_upload(B.pos);
                        temps.push_back(d_border); temps.push_back(d_bpos);
                    }
                    put_matrix(bk->A, lb, c.A, B, d_border, &B, d_bpos, gs_kind, &B.wf_row_ptr, "A(back)", l);
                }
                bk->A.v.recip = 1;
                max_items = std::max(max_items, lb.nitems());
                bk->x = dev_alloc<double>(lv.n);
                bk->b = dev_alloc<double>(lv.n + 2);
                setup_smoother(*bk, lb, B, c.A);
                std::vector<int> fb((size_t)lv.n);
                for (int k = 0; k < lv.n; ++k) fb[k] = S.pos[B.order[k]];
                lv.d_fb = dev_upload(fb);
            }
            // transfers: P_l rows in this level's schedule, columns in the next level's; R_l the other way round
            DevLayout lp, lr;
            put_matrix(lv.P, lp, c.P, S, lv.d_order, &sched[l + 1], d_pos[l + 1], kind_of(c.P), nullptr, "P", l);
            max_items = std::max(max_items, lp.nitems());
            if (!worker) {                                      // (restriction and everything below it run on rank 0)
                put_matrix(lv.R, lr, c.R, sched[l + 1], h->L[l + 1].d_order, &S, d_pos[l], kind_of(c.R), nullptr, "R", l);
                max_items = std::max(max_items, lr.nitems());
            }
            // residual (+) restriction in one launch: both operators thread-per-row (the levels that carry the bytes)
            const DevLayout &la = lv.Asp.valid ? lsp_keep : lay;
            // (measured: the instances for rows longer than one register chunk need 80-114 registers and lose to the separate kernels --
            // 256^3 level 1: 691 vs 586 us -- so only single-chunk operators, i.e. level 0 of the 5-/7-point problems, take it by default)
            const bool rr_short = la.max_row <= 8 && lr.max_row <= 8;
            if (use_fused && !worker && (rr_short || rr_all) && la.kind == KIND_SELL && lr.kind == KIND_SELL && lv.n >= rr_min_rows && la.nitems() >= 1 && lr.nitems() >= 1) {
                const double tl = now_s();
                build_fused(lv, la, lr, S, sched[l + 1], c.R);
                tl_note("fused", l, now_s() - tl);
            }
        }
    }
    for (std::thread &t : lower_threads) t.join();
    h->partial_stride = std::max(1184, (max_items + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    h->d_partial = dev_alloc<double>((size_t)4 * h->partial_stride);
    h->d_scal = dev_alloc<double>(8);
    {
        // pinned read-back slots are recycled across hierarchies (cudaMallocHost / cudaFreeHost cost up to tens of milliseconds each)
        std::lock_guard<std::mutex> lk(global_mutex());
        std::vector<double *> &fl = pinned_scalars();
        if (!fl.empty()) { h->h_scal = fl.back(); fl.pop_back(); }
        else CUDA_CHECK(cudaMallocHost((void **)&h->h_scal, 8 * sizeof(double)));
    }
    if (getenv("AMGB200_DEBUG_TIMING")) h->d_dbg = dev_alloc<long long>(16 * 8 + 8);
    h->d_xnat = dev_alloc<double>(maxn);
    h->d_bnat = dev_alloc<double>(maxn);
    prefetch.finish(temps);
    CUDA_CHECK(cudaDeviceSynchronize());
    for (void *p : temps) dev_free(p);                  // raw CSR copies and numbering tables of the device-side fills
    h->upload_s = now_s() - t0;
    if (opt.verbose >= 2) {
        printf("libamgb200: %d levels resident; schedule analysis %.3f s, layout build %.3f s, cudaMalloc+H2D %.3f s, total %.3f s\n", nl, h->analysis_s, t_layout, DevMatOwner::t_upload(), h->upload_s);
        printf("  mode %s; lvl       rows         nnz  kind  F-rows  wavefronts F/C   P nnz      R nnz   sym strat\n", h->exact ? "EXACT" : "FAST");
        for (int l = 0; l < nl; ++l) {
            const Level &lv = h->L[l];
            printf("  %3d %10d %11lld  %s %8d  %6d/%-6d %9lld %9lld  %d\n", l, lv.n, lv.A.nnz, lv.A.v.kind == KIND_SELL ? "SELL" : "CSR ",
                   lv.pass_rows[0], lv.wf_count[0], lv.wf_count[1], lv.P.valid ? lv.P.nnz : 0LL, lv.R.valid ? lv.R.nnz : 0LL, (int)lv.pattern_symmetric);
            printf("      strategy %d%s  max wavefront width %d items\n", lv.strategy, lv.x_in_smem ? " (x in smem)" : "", lv.max_width);
            if (lv.strategy == 2) printf("      CTA pipeline: %d groups x %d warps, parked suffix capacity %d\n", lv.cta_D, lv.cta_G, lv.cta_cap);
        }
    }
    return h;
}

void amgb200_free(amgb200_hier *h) {
    if (!h) return;
    cudaStreamSynchronize(h->stream);
    for (Level &lv : h->L) {
        lv.A.release(); lv.Asp.release(); lv.P.release(); lv.R.release();
        dev_free(lv.d_order); dev_free(lv.x); dev_free(lv.b); dev_free(lv.wp);
        dev_free(lv.d_item_wf); dev_free(lv.d_wf_item_ptr); dev_free(lv.d_fb); dev_free(lv.d_stream); dev_free(lv.d_blk_ptr); dev_free(lv.d_wf_row_ptr);
        dev_free(lv.d_rec); dev_free(lv.d_hint);
        lv.rr.release();
        if (lv.lo) {
            lv.lo->A.release();
            dev_free(lv.lo->d_item_wf); dev_free(lv.lo->d_wf_item_ptr); dev_free(lv.lo->d_stream); dev_free(lv.lo->d_blk_ptr); dev_free(lv.lo->d_wf_row_ptr);
            delete lv.lo;
        }
        if (lv.bk) {
            dev_free(lv.bk->d_rec); dev_free(lv.bk->d_hint);
            lv.bk->A.release();
            dev_free(lv.bk->x); dev_free(lv.bk->b);
            dev_free(lv.bk->d_item_wf); dev_free(lv.bk->d_wf_item_ptr); dev_free(lv.bk->d_stream); dev_free(lv.bk->d_blk_ptr); dev_free(lv.bk->d_wf_row_ptr);
            delete lv.bk;
        }
    }
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    for (auto &pl : h->peer_plan) { dev_free(pl.d_push); dev_free(pl.d_flag_ptr); dev_free(pl.d_src); }    // (index lists: a few KB, released with the pool)
    dev_free(h->d_peer_flags);
    dev_free(h->d_partial); dev_free(h->d_scal);
    if (h->h_scal) { std::lock_guard<std::mutex> lk(global_mutex()); pinned_scalars().push_back(h->h_scal); }
    dev_free(h->d_xnat); dev_free(h->d_bnat); dev_free(h->kry); dev_free(h->d_dbg);
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

int amgb200_num_levels(const amgb200_hier *h) { return h->nl; }

void amgb200_level_info(const amgb200_hier *h, int level, long long info[8]) {
    check_level(h, level);
    const Level &lv = h->L[level];
    info[0] = lv.n; info[1] = lv.A.nnz; info[2] = lv.wf_count[0]; info[3] = lv.wf_count[1];
    info[4] = lv.A.v.kind; info[5] = lv.pass_rows[0];
    info[6] = lv.P.valid ? lv.P.nnz : 0; info[7] = lv.R.valid ? lv.R.nnz : 0;
}

double amgb200_algorithmic_bytes(const amgb200_hier *h, int level, int op) {
    auto S = [](long long z, long long n) { return 12.0 * z + 4.0 * (n + 1); };
    auto lvl = [&](int l, int o) -> double {
        const Level &lv = h->L[l];
        const double n = lv.n, SA = S(lv.A.nnz, lv.n);
        const double nc = l + 1 < h->nl ? h->L[l + 1].n : 0;
        switch (o) {
            case 0: return SA + 4 * n + 8 * n + 16 * n;                       // one GS sweep (both passes)
            case 1: return SA + 24 * n;                                       // r = b - A x
            case 2: return lv.R.valid ? S(lv.R.nnz, (long long)nc) + 8 * n + 8 * nc : 0;   // b_c = R r
            case 3: return lv.P.valid ? S(lv.P.nnz, lv.n) + 8 * nc + 16 * n : 0;           // x += P e
            case 4: return SA + 16 * n;                                       // y = A x
            case 6: return lv.R.valid ? SA + 16 * n + S(lv.R.nnz, (long long)nc) + 8 * nc + 8 * nc : 0;   // residual (+) restriction (+ x_c = 0), r not counted
            default: return 0;
        }
    };
    if (op != 5) { check_level(h, level); return lvl(level, op); }
    double tot = 0;                                                           // fused op sequence of one V-cycle
    for (int l = 0; l + 1 < h->nl; ++l) {
        const Level &lv = h->L[l];
        const double n = lv.n, nc = h->L[l + 1].n;
        tot += (h->pars.pre_iter + h->pars.post_iter) * lvl(l, 0);
        tot += S(lv.A.nnz, lv.n) + 16 * n + S(lv.R.nnz, (long long)nc) + 8 * nc;   // residual (+) restrict, r not materialised
        tot += 8 * nc;                                                        // x_c = 0
        tot += lvl(l, 3);
    }
    tot += S(h->L[0].A.nnz, h->L[0].n) + 16.0 * h->L[0].n;                    // outer residual (+) norm
    return tot;
}

void amgb200_last_phase_ms(const amgb200_hier *h, double ms[8]) { for (int i = 0; i < 8; ++i) ms[i] = h->phase_ms[i]; }
void amgb200_last_level_ms(const amgb200_hier *h, int level, double ms[4]) {
    check_level(h, level);
    for (int i = 0; i < 4; ++i) ms[i] = h->L[level].prof_ms[i];
}
void amgb200_set_profile(amgb200_hier *h, int on) { h->profile = on != 0; }
void amgb200_upload_seconds(const amgb200_hier *h, double s[2]) { s[0] = h->analysis_s; s[1] = h->upload_s; }
long long amgb200_device_bytes(const amgb200_hier *h) {
    long long tot = 0;
    for (const Level &lv : h->L) {
        for (const DevMatOwner *m : {&lv.A, &lv.Asp, &lv.P, &lv.R}) if (m->valid) tot += m->padded * 12 + (long long)m->v.nitems * 12;
        tot += (long long)lv.n * (3 * 8 + 4);
    }
    return tot;
}
long long amgb200_level_chain_terms(const amgb200_hier *h, int level) {
    check_level(h, level);
    const Level &lv = h->L[level];
    const int sweeps = h->pars.pre_iter + h->pars.post_iter;
    if (lv.lo && sweeps > 0) return (lv.chain_terms * (sweeps - 1) + lv.lo->chain_terms) / sweeps;     // (mean per sweep of a cycle: the first one runs on the lower triangle)
    return lv.chain_terms;
}
int amgb200_level_fused(const amgb200_hier *h, int level) {
    check_level(h, level);
    return h->L[level].rr.valid ? 1 : 0;
}
const char *amgb200_level_kernel(const amgb200_hier *h, int level) {
    check_level(h, level);
    const Level &lv = h->L[level];
    if (!lv.smoothed) return "none";
    static const char *names[8] = {"gs_pass_kernel", "-", "gs_ordered_cta_kernel", "gs_ordered_cluster_kernel", "gs_stream_cta_kernel", "gs_stream_cluster_kernel", "gs_dataflow_kernel", "gs_dataflow_csr_kernel"};
    return names[lv.strategy];
}

// K timed solves from the same initial guess with CUDA events on the library's stream.
// d_x0, d_b, d_x: device arrays (natural numbering); d_x is overwritten each step.
void amgb200_bench_solve(amgb200_hier *h, const double *d_x0, const double *d_b, double *d_x, int warmup, int steps,
                         double *ms_total, amgb200_rtn *last) {
    const size_t bytes = (size_t)h->L[0].n * sizeof(double);
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    amgb200_rtn r = {0, 0, 0};
    for (int i = 0; i < warmup; ++i) {
        CUDA_CHECK(cudaMemcpyAsync(d_x, d_x0, bytes, cudaMemcpyDeviceToDevice, h->stream));
        r = amgb200_solve_device(h, d_x, d_b, nullptr, 0);
    }
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    CUDA_CHECK(cudaEventRecord(a, h->stream));
    for (int i = 0; i < steps; ++i) {
        CUDA_CHECK(cudaMemcpyAsync(d_x, d_x0, bytes, cudaMemcpyDeviceToDevice, h->stream));
        r = amgb200_solve_device(h, d_x, d_b, nullptr, 0);
    }
    CUDA_CHECK(cudaEventRecord(b, h->stream));
    CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_total = ms;
    if (last) *last = r;
}

amgb200_rtn amgb200_solve_device(amgb200_hier *h, double *d_x, const double *d_b, double *res_hist, int hist_cap) {
    amgb200_rtn rtn = {0, 0, 0};
    Level &l0 = h->L[0];
    for (int i = 0; i < 8; ++i) h->phase_ms[i] = 0;
    for (Level &lv : h->L) for (int i = 0; i < 4; ++i) lv.prof_ms[i] = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (h->profile) { CUDA_CHECK(cudaEventCreate(&t0)); CUDA_CHECK(cudaEventCreate(&t1)); CUDA_CHECK(cudaEventRecord(t0, h->stream)); }
    to_schedule(h, 0, d_b, l0.b);
    to_schedule(h, 0, d_x, l0.x);
    const double sumb = dev_norm2_tree(h, l0.n, l0.b);
    double absres0 = sumb;
    if (h->opt.verbose >= 1) print_itinfo(0, 1.0, sumb, 0.0);
    if (fabs(sumb) == 0.) {                                                   // SSS_SOLVE.c:41-46
        CUDA_CHECK(cudaMemsetAsync(d_x, 0, (size_t)l0.n * sizeof(double), h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (t0) { cudaEventDestroy(t0); cudaEventDestroy(t1); }
        return rtn;
    }
    int iter = 0;
    while (++iter <= h->pars.max_it) {
        cycle(h);
        double absres;
        { PhaseTimer pt(h, 5); absres = dev_true_residual(h, l0.spmvA(), l0.x, l0.b, l0.wp); }
        const double relres = absres / sumb;
        if (h->opt.verbose >= 1) print_itinfo(iter, relres, absres, absres / absres0);
        absres0 = absres;
        if (res_hist && iter <= hist_cap) res_hist[iter - 1] = absres;
        rtn.ares = absres; rtn.rres = relres; rtn.nits = iter;
        if (relres < h->pars.tol) break;
    }
    to_natural(h, 0, l0.x, d_x);
    if (h->profile) {
        CUDA_CHECK(cudaEventRecord(t1, h->stream)); CUDA_CHECK(cudaEventSynchronize(t1));
        float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, t0, t1)); h->phase_ms[6] = ms;
        cudaEventDestroy(t0); cudaEventDestroy(t1);
    }
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return rtn;
}

amgb200_rtn amgb200_solve(amgb200_hier *h, double *x, const double *b, double *res_hist, int hist_cap) {
    const int n = h->L[0].n;
    CUDA_CHECK(cudaMemcpyAsync(h->d_xnat, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->d_bnat, b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    amgb200_rtn rtn = amgb200_solve_device(h, h->d_xnat, h->d_bnat, res_hist, hist_cap);
    CUDA_CHECK(cudaMemcpyAsync(x, h->d_xnat, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return rtn;
}

void amgb200_cycle(amgb200_hier *h, double *x, const double *b) {
    Level &l0 = h->L[0];
    const int n = l0.n;
    CUDA_CHECK(cudaMemcpyAsync(h->d_xnat, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->d_bnat, b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    to_schedule(h, 0, h->d_bnat, l0.b);
    to_schedule(h, 0, h->d_xnat, l0.x);
    cycle(h);
    to_natural(h, 0, l0.x, h->d_xnat);
    CUDA_CHECK(cudaMemcpyAsync(x, h->d_xnat, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

void amgb200_level_spmv(amgb200_hier *h, int level, int which, double alpha, const double *x, int beta, double *y) {
    check_level(h, level);
    Level &lv = h->L[level];
    if (which != 0 && !lv.P.valid) { fprintf(stderr, "libamgb200: level %d has no transfer operators\n", level); exit(72); }
    const int in_l = which == 1 ? level + 1 : level, out_l = which == 2 ? level + 1 : level;
    const DMat &M = which == 0 ? lv.spmvA() : which == 1 ? lv.P.v : lv.R.v;
    Level &li = h->L[in_l], &lo = h->L[out_l];
    // scratch: input in li.wp (schedule), output in lo.b ... use level vectors that the hooks own
    double *din = li.wp, *dout = lo.b;
    if (in_l == out_l) dout = lo.b;
    CUDA_CHECK(cudaMemcpyAsync(h->d_xnat, x, (size_t)li.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    to_schedule(h, in_l, h->d_xnat, din);
    if (beta) {
        CUDA_CHECK(cudaMemcpyAsync(h->d_bnat, y, (size_t)lo.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        to_schedule(h, out_l, h->d_bnat, dout);
        spmv(h, M, MODE_AMXPY, RED_NONE, din, dout, nullptr, alpha);
    } else {
        spmv(h, M, MODE_MXY, RED_NONE, din, dout, nullptr, 0.0);
        if (alpha != 1.0) dev_scale(h, lo.n, alpha, dout);
    }
    to_natural(h, out_l, dout, h->d_bnat);
    CUDA_CHECK(cudaMemcpyAsync(y, h->d_bnat, (size_t)lo.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

void amgb200_level_smooth(amgb200_hier *h, int level, int nsweeps, double *x, const double *b) {
    check_level(h, level);
    Level &lv = h->L[level];
    if (!lv.smoothed) { fprintf(stderr, "libamgb200: level %d is the coarsest level (no smoother)\n", level); exit(72); }
    CUDA_CHECK(cudaMemcpyAsync(h->d_xnat, x, (size_t)lv.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->d_bnat, b, (size_t)lv.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    to_schedule(h, level, h->d_xnat, lv.x);
    to_schedule(h, level, h->d_bnat, lv.b);
    smooth(h, level, nsweeps < 0 ? -nsweeps : nsweeps, nsweeps < 0);
    to_natural(h, level, lv.x, h->d_xnat);
    CUDA_CHECK(cudaMemcpyAsync(x, h->d_xnat, (size_t)lv.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

double amgb200_level_residual(amgb200_hier *h, int level, const double *x, const double *b, double *r) {
    check_level(h, level);
    Level &lv = h->L[level];
    CUDA_CHECK(cudaMemcpyAsync(h->d_xnat, x, (size_t)lv.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->d_bnat, b, (size_t)lv.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    to_schedule(h, level, h->d_xnat, lv.x);
    to_schedule(h, level, h->d_bnat, lv.b);
    const double nrm = dev_true_residual(h, lv.spmvA(), lv.x, lv.b, lv.wp);
    if (r) {
        to_natural(h, level, lv.wp, h->d_xnat);
        CUDA_CHECK(cudaMemcpyAsync(r, h->d_xnat, (size_t)lv.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    return nrm;
}

// wp = b - A x and b_{level+1} = R wp from host vectors in natural numbering, the way the cycle computes them (one fused launch
// where the level has one, else two); returns 1 when the fused kernel ran
int amgb200_level_resid_restrict(amgb200_hier *h, int level, const double *x, const double *b, double *r, double *bc) {
    check_level(h, level);
    check_level(h, level + 1);
    Level &lv = h->L[level];
    Level &lc = h->L[level + 1];
    if (!lv.R.valid) { fprintf(stderr, "libamgb200: level %d has no transfer operators\n", level); exit(72); }
    CUDA_CHECK(cudaMemcpyAsync(h->d_xnat, x, (size_t)lv.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->d_bnat, b, (size_t)lv.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    to_schedule(h, level, h->d_xnat, lv.x);
    to_schedule(h, level, h->d_bnat, lv.b);
    if (lv.rr.valid) resid_restrict(h, level, true);
    else {
        spmv(h, lv.spmvA(), MODE_RESID, RED_NONE, lv.x, lv.wp, lv.b, -1.0);
        spmv(h, lv.R.v, MODE_MXY, RED_NONE, lv.wp, lc.b, nullptr, 0.0);
        dev_zero(h, lc.n, lc.x);
    }
    lc.x_is_zero = true;
    to_natural(h, level, lv.wp, h->d_xnat);
    CUDA_CHECK(cudaMemcpyAsync(r, h->d_xnat, (size_t)lv.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    to_natural(h, level + 1, lc.b, h->d_bnat);
    CUDA_CHECK(cudaMemcpyAsync(bc, h->d_bnat, (size_t)lc.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return lv.rr.valid ? 1 : 0;
}

int amgb200_coarse_solve(amgb200_hier *h, double *x, const double *b, double tol, int its[2]) {
    Level &lv = h->L[h->nl - 1];
    CUDA_CHECK(cudaMemcpyAsync(h->d_xnat, x, (size_t)lv.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->d_bnat, b, (size_t)lv.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    to_schedule(h, h->nl - 1, h->d_xnat, lv.x);
    to_schedule(h, h->nl - 1, h->d_bnat, lv.b);
    const int st = coarse_solve(h, lv.A.v, lv.b, lv.x, tol, its);
    to_natural(h, h->nl - 1, lv.x, h->d_xnat);
    CUDA_CHECK(cudaMemcpyAsync(x, h->d_xnat, (size_t)lv.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return st;
}

double amgb200_time_op(amgb200_hier *h, int level, int op, int reps) {
    check_level(h, level);
    Level &lv = h->L[level];
    if (reps < 1) reps = 1;
    if ((op == 0 && !lv.smoothed) || ((op == 2 || op == 3 || op == 6) && !lv.P.valid)) return 0.0;
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    const int nsw = getenv("AMGB200_TIMEOP_SWEEPS") ? std::max(1, atoi(getenv("AMGB200_TIMEOP_SWEEPS"))) : 1;   // (developer probe: sweeps per launch)
    auto run = [&]() {
        switch (op) {
            case 0: smooth(h, level, nsw); break;
            case 1: spmv(h, lv.spmvA(), MODE_RESID, RED_NONE, lv.x, lv.wp, lv.b, -1.0); break;
            case 2: spmv(h, lv.R.v, MODE_MXY, RED_NONE, lv.wp, h->L[level + 1].b, nullptr, 0.0); break;
            case 3: spmv(h, lv.P.v, MODE_AMXPY, RED_NONE, h->L[level + 1].x, lv.x, nullptr, 1.0); break;
            case 6:                                  // what the cycle runs between pre-smoothing and the next level: fused, or three launches
                if (lv.rr.valid) resid_restrict(h, level, true);
                else {
                    spmv(h, lv.spmvA(), MODE_RESID, RED_NONE, lv.x, lv.wp, lv.b, -1.0);
                    spmv(h, lv.R.v, MODE_MXY, RED_NONE, lv.wp, h->L[level + 1].b, nullptr, 0.0);
                    dev_zero(h, h->L[level + 1].n, h->L[level + 1].x);
                }
                break;
            default: spmv(h, lv.spmvA(), MODE_MXY, RED_NONE, lv.x, lv.wp, nullptr, 0.0); break;
        }
    };
    run(); run();                                    // warm-up
    CUDA_CHECK(cudaEventRecord(a, h->stream));
    for (int i = 0; i < reps; ++i) run();
    CUDA_CHECK(cudaEventRecord(b, h->stream));
    CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms / reps / (op == 0 ? nsw : 1);
}


// ---- multi-GPU building blocks (amg_b200/distributed.py drives them; one process per GPU) ----
void amgb200_set_stream(amgb200_hier *h, void *stream) {
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
}
void *amgb200_level_vec(amgb200_hier *h, int level, int which) {
    check_level(h, level);
    Level &lv = h->L[level];
    return which == 0 ? (void *)lv.x : which == 1 ? (void *)lv.b : (void *)lv.wp;
}
// level vector (0 x, 1 b, 2 wp) in NATURAL numbering to a host array of n_level doubles
void amgb200_level_download(amgb200_hier *h, int level, int which, double *host) {
    check_level(h, level);
    Level &lv = h->L[level];
    const double *src = which == 0 ? lv.x : which == 1 ? lv.b : lv.wp;
    to_natural(h, level, src, h->d_bnat);
    CUDA_CHECK(cudaMemcpyAsync(host, h->d_bnat, (size_t)lv.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
}
void amgb200_level_order(const amgb200_hier *h, int level, int *order_host) {
    check_level(h, level);
    CUDA_CHECK(cudaMemcpy(order_host, h->L[level].d_order, (size_t)h->L[level].n * sizeof(int), cudaMemcpyDeviceToHost));
}
// info: [0] n, [1] rows in the F pass, [2] items (slices/rows) in the F pass, [3] items in the C pass,
//       [4] layout kind of A, [5] 1 if both passes are fully parallel (shardable), [6] rows per item (32 | 1), [7] items of P
void amgb200_l0_shape(const amgb200_hier *h, long long info[8]) {
    const Level &lv = h->L[0];
    info[0] = lv.n; info[1] = lv.pass_rows[0]; info[2] = lv.pass_items[0]; info[3] = lv.pass_items[1];
    info[4] = lv.A.v.kind; info[5] = (lv.smoothed && !lv.ordered) ? 1 : 0; info[6] = lv.A.v.kind == KIND_SELL ? 32 : 1;
    info[7] = lv.P.valid ? lv.P.v.nitems : 0;
    // (the item ranges of amgb200_l0_prolong are derived from row ranges with info[6]: only meaningful when P uses A's layout kind)
    if (lv.P.valid && lv.P.v.kind != lv.A.v.kind) info[5] = 0;
}
// one Gauss-Seidel pass (0 = F, 1 = C) over the pass-relative item range [item0, item1) of level 0
void amgb200_l0_gs_pass(amgb200_hier *h, int pass, int item0, int item1) {
    Level &lv = h->L[0];
    if (!lv.smoothed || lv.ordered) { fprintf(stderr, "libamgb200: level 0 is not two-colour; it cannot be sharded\n"); exit(73); }
    if (pass < 0 || pass > 1 || item0 < 0 || item1 > lv.pass_items[pass]) { fprintf(stderr, "libamgb200: item range [%d,%d) outside pass %d (%d items)\n", item0, item1, pass, pass >= 0 && pass <= 1 ? lv.pass_items[pass] : 0); exit(72); }
    const int base = pass ? lv.pass_items[0] : 0;
    const int cnt = item1 - item0;
    if (cnt <= 0) return;
    const int grid = (cnt + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    if (lv.A.v.kind == KIND_SELL && lv.A.v.max_row <= 8) LAUNCH((gs_pass_kernel<0, true, true>), grid, BLOCK, h->stream, lv.A.v, lv.b, lv.x, base + item0, base + item1);
    else if (lv.A.v.kind == KIND_SELL) LAUNCH((gs_pass_kernel<0, true>), grid, BLOCK, h->stream, lv.A.v, lv.b, lv.x, base + item0, base + item1);
    else if (h->exact) LAUNCH((gs_pass_kernel<1, true>), grid, BLOCK, h->stream, lv.A.v, lv.b, lv.x, base + item0, base + item1);
    else LAUNCH((gs_pass_kernel<1, false>), grid, BLOCK, h->stream, lv.A.v, lv.b, lv.x, base + item0, base + item1);
}
// wp = b - A x on the absolute item range [item0, item1) of level 0's smoother layout
void amgb200_l0_residual(amgb200_hier *h, int item0, int item1) {
    Level &lv = h->L[0];
    if (item0 < 0 || item1 > lv.A.v.nitems) { fprintf(stderr, "libamgb200: item range [%d,%d) outside level 0 (%d items)\n", item0, item1, lv.A.v.nitems); exit(72); }
    if (item1 <= item0) return;
    h->item0 = item0; h->item1 = item1;
    spmv(h, lv.A.v, MODE_RESID, RED_NONE, lv.x, lv.wp, lv.b, -1.0);
    h->item0 = h->item1 = -1;
}
// x_0 += P_0 x_1 on the item range [item0, item1) of P_0 (32 schedule rows per item for SELL)
void amgb200_l0_prolong(amgb200_hier *h, int item0, int item1) {
    Level &lv = h->L[0];
    if (!lv.P.valid) { fprintf(stderr, "libamgb200: level 0 has no transfer operators\n"); exit(72); }
    if (item0 < 0 || item1 > lv.P.v.nitems) { fprintf(stderr, "libamgb200: item range [%d,%d) outside P_0 (%d items)\n", item0, item1, lv.P.v.nitems); exit(72); }
    if (item1 <= item0) return;
    h->item0 = item0; h->item1 = item1;
    spmv(h, lv.P.v, MODE_AMXPY, RED_NONE, h->L[1].x, lv.x, nullptr, 1.0);
    h->item0 = h->item1 = -1;
}
// b_{level+1} = R_level wp_level ; x_{level+1} = 0   (SSS_cycle.cu:921,929)
void amgb200_restrict_from(amgb200_hier *h, int level) {
    check_level(h, level);
    check_level(h, level + 1);
    Level &lv = h->L[level];
    if (!lv.R.valid) { fprintf(stderr, "libamgb200: level %d has no transfer operators\n", level); exit(72); }
    spmv(h, lv.R.v, MODE_MXY, RED_NONE, lv.wp, h->L[level + 1].b, nullptr, 0.0);
    dev_zero(h, h->L[level + 1].n, h->L[level + 1].x);
    h->L[level + 1].x_is_zero = true;
}
// the cycle on levels >= level (b_level and x_level already set), V-cycle only when level > 0
void amgb200_cycle_from(amgb200_hier *h, int level) {
    check_level(h, level);
    if (level == h->nl - 1) {
        double tol = h->pars.ctol;
        if (tol > h->pars.tol) tol = h->pars.tol * 0.1;
        coarse_solve(h, h->L[level].A.v, h->L[level].b, h->L[level].x, tol, nullptr);
    } else cycle_from(h, level);
}
void amgb200_vec_to_schedule(amgb200_hier *h, int level, const double *d_nat, double *d_sched) { check_level(h, level); to_schedule(h, level, d_nat, d_sched); }
void amgb200_vec_to_natural(amgb200_hier *h, int level, const double *d_sched, double *d_nat) { check_level(h, level); to_natural(h, level, d_sched, d_nat); }
// test hook (not part of the documented ABI): number of operand pairs for which the precomputed-reciprocal quotient of the
// ordered smoothers differs from __ddiv_rn
__attribute__((visibility("default"))) long long amgb200_debug_quotient_check(long long n, unsigned long long seed, int mode) {
    unsigned long long *d_bad = nullptr, bad = 0;
    CUDA_CHECK(cudaMalloc(&d_bad, sizeof(bad)));
    CUDA_CHECK(cudaMemset(d_bad, 0, sizeof(bad)));
    quotient_check_kernel<<<1184, BLOCK>>>(n, seed, mode, d_bad);
    ++g_launches;
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaFree(d_bad));
    return (long long)bad;
}
// ---- peer-memory exchange (NVLink / NVSwitch P2P through CUDA IPC; one process per GPU) ----
// which: 0 x, 1 b, 2 wp of `level`; 3: this rank's flag words (level ignored).  handle: 64 bytes (cudaIpcMemHandle_t)
void amgb200_ipc_export(amgb200_hier *h, int level, int which, unsigned char *handle) {
    void *p = nullptr;
    if (which == 3) {
        if (!h->d_peer_flags) {
            h->d_peer_flags = dev_alloc<unsigned>((size_t)amgb200_hier::PEER_MAX_PLANS * amgb200_hier::PEER_MAX_RANKS);
            CUDA_CHECK(cudaMemset(h->d_peer_flags, 0, sizeof(unsigned) * amgb200_hier::PEER_MAX_PLANS * amgb200_hier::PEER_MAX_RANKS));
        }
        p = h->d_peer_flags;
    } else p = amgb200_level_vec(h, level, which);
    cudaIpcMemHandle_t hd;
    CUDA_CHECK(cudaIpcGetMemHandle(&hd, p));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t");
    memcpy(handle, &hd, 64);
}
void *amgb200_ipc_open(amgb200_hier *h, const unsigned char *handle) {
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    void *p = nullptr;
    CUDA_CHECK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(p);
    return p;
}
// Plan `plan` (0..7) of this rank: npush transfers  my_vec[idx] -> peer_vec[idx]  (idx[i] == nullptr: the contiguous range
// [range0[i], range1[i]); idx arrays are host arrays, copied), after which the epoch of the run is stored into flag_slot[j] (device
// addresses inside the PEERS' flag arrays: peer_flags + plan * 64 + my rank), j < nflag; then this rank waits until the flag words
// [plan][src[i]], i < nsrc, of its own array have reached the epoch.  Every rank must run its plans in the same global order.
void amgb200_peer_plan(amgb200_hier *h, int plan, int npush, const double *my_vec, void *const *peer_vec, const int *const *idx, const int *count,
                       const int *range0, int nflag, void *const *flag_slot, int nsrc, const int *src) {
    if (plan < 0 || plan >= amgb200_hier::PEER_MAX_PLANS) { fprintf(stderr, "libamgb200: peer plan %d out of range\n", plan); exit(72); }
    amgb200_hier::PeerPlan &pl = h->peer_plan[plan];
    std::vector<PeerPush> pp((size_t)npush);
    pl.max_count = 1;
    for (int i = 0; i < npush; ++i) {
        pp[i].count = count[i];
        pl.max_count = std::max(pl.max_count, count[i]);
        if (idx[i]) {
            std::vector<int> v(idx[i], idx[i] + count[i]);
            pp[i].idx = dev_upload(v);
            pp[i].src = my_vec; pp[i].dst = (double *)peer_vec[i];
        } else {
            pp[i].idx = nullptr;
            pp[i].src = my_vec + range0[i]; pp[i].dst = (double *)peer_vec[i] + range0[i];
        }
    }
    pl.npush = npush;
    pl.d_push = npush ? dev_upload(pp) : nullptr;
    std::vector<unsigned *> fl((size_t)nflag);
    for (int i = 0; i < nflag; ++i) fl[i] = (unsigned *)flag_slot[i];
    pl.nflag = nflag;
    pl.d_flag_ptr = nflag ? dev_upload(fl) : nullptr;
    std::vector<int> sr((size_t)nsrc);
    for (int i = 0; i < nsrc; ++i) sr[i] = plan * amgb200_hier::PEER_MAX_RANKS + src[i];
    pl.nsrc = nsrc;
    pl.d_src = nsrc ? dev_upload(sr) : nullptr;
    pl.epoch = 0;
}
// one exchange = start (store my entries into the peers' vectors, then raise the epoch in their flag words) + wait (for the peers' epochs in
// my flag words).  Kernels launched between the two run while the peers' stores are in flight: the interior rows of a level-0 pass, which
// read no ghost entry, overlap the halo exchange (amg_b200/distributed.py: smooth).
void amgb200_peer_start(amgb200_hier *h, int plan) {
    if (plan < 0 || plan >= amgb200_hier::PEER_MAX_PLANS) { fprintf(stderr, "libamgb200: peer plan %d out of range\n", plan); exit(72); }
    amgb200_hier::PeerPlan &pl = h->peer_plan[plan];
    ++pl.epoch;
    if (pl.npush) {
        const int gx = std::max(1, std::min(64, (pl.max_count + BLOCK - 1) / BLOCK));
        peer_push_kernel<<<dim3(gx, pl.npush), BLOCK, 0, h->stream>>>(pl.d_push);
        ++g_launches;
    }
    if (pl.nflag) { peer_flag_kernel<<<1, 64, 0, h->stream>>>(pl.d_flag_ptr, pl.nflag, pl.epoch); ++g_launches; }
    CUDA_CHECK(cudaGetLastError());
}
void amgb200_peer_wait(amgb200_hier *h, int plan) {
    if (plan < 0 || plan >= amgb200_hier::PEER_MAX_PLANS) { fprintf(stderr, "libamgb200: peer plan %d out of range\n", plan); exit(72); }
    amgb200_hier::PeerPlan &pl = h->peer_plan[plan];
    if (pl.nsrc) { peer_wait_kernel<<<1, 64, 0, h->stream>>>(h->d_peer_flags, pl.d_src, pl.nsrc, pl.epoch); ++g_launches; }
    CUDA_CHECK(cudaGetLastError());
}
void amgb200_peer_run(amgb200_hier *h, int plan) {
    amgb200_peer_start(h, plan);
    amgb200_peer_wait(h, plan);
}
void amgb200_sync(amgb200_hier *h) { CUDA_CHECK(cudaStreamSynchronize(h->stream)); }

}  // extern "C"
