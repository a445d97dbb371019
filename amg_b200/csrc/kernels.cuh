// sm_100a kernels of the AMG solve phase.  All vectors and matrices are in the level's
// schedule numbering (analysis.h); fp64 values, int32 indices.
//
// Arithmetic contract (EXACT mode, the default).  A 1-ulp difference in x is amplified by
// |x|/|r| in the residual, so the north-star bar "residual history within 1e-10 relative" can
// only be met when every operation that feeds x is rounded exactly like the reference's CPU code
// (gcc, baseline x86-64: no FMA, strictly sequential sums in CSR storage order).  Therefore:
//   * every product a_k*x_j is rounded on its own (__dmul_rn, never contracted),
//   * every row sum is accumulated in storage order by one dependent chain of __dadd_rn/__dsub_rn
//     -- thread-per-row for short rows (SELL layout, coalesced across rows), warp-per-row for long
//     rows, where the 32 lanes fetch and multiply 32 entries in parallel and the chain consumes the
//     products in order through warp shuffles (adding +0.0 for padding is exact),
//   * dot products that feed the Krylov coefficients are summed left to right by one thread.
// Norms that only steer printing / stopping use deterministic tree reductions.
// FAST mode (opt-in) replaces the in-order chain of the warp-per-row kernels by a shuffle tree:
// ~1e-16 relative per row, which the |x|/|r| amplification turns into ~1e-7 at the last V-cycle.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace amgb200 {

struct DMat {                      // device view of a DevLayout
    int kind, nrows, ncols, nitems, max_row;
    int recip;                     // Gauss-Seidel update x = t * (1/d) (natural-order smoother, SSS_smooth.c:112-118) instead of t / d
    const int *slice_row;          // SELL
    const long long *slice_ptr;    // SELL
    const int *rptr;               // CSR
    const int *split;              // CSR, ordered levels: per row, index of the first entry that reads the previous wavefront (or NULL)
    const unsigned *late;          // CSR, ordered levels: one bit per entry, set when it reads the previous wavefront (or NULL)
    const int *col;
    const double *val;
};

enum { MODE_MXY = 0, MODE_AMXPY = 1, MODE_RESID = 2 };   // y = Ax | y += alpha*Ax | y = b - Ax
enum { RED_NONE = 0, RED_SUMSQ = 1 };                    // fused tree reduction over the output rows

constexpr int WARPS_PER_BLOCK = 8;
constexpr int BLOCK = 32 * WARPS_PER_BLOCK;
constexpr double GS_TINY = 1e-20;                         // SMALLFLOAT, amg/SSS_main.h:34
constexpr unsigned FULL = 0xffffffffu;

// COH: x is being updated by other SMs during this launch -> read it at L2 (L1 is not coherent)
template <bool COH>
__device__ __forceinline__ double ld_x(const double *p) { return COH ? __ldcg(p) : *p; }

// x_i from the accumulated t and the diagonal d: t / d (SSS_smooth.c:32) or t * (1/d) (SSS_smooth.c:112-118)
__device__ __forceinline__ double gs_quotient(double t, double d, int recip) {
    return recip ? __dmul_rn(t, __ddiv_rn(1.0, d)) : __ddiv_rn(t, d);
}

// The same quotient with everything that depends only on d taken off the dependency path: y = RN(1/d) (an IEEE division,
// done while the row still waits for its inputs) and `dsafe` (d in a range where nothing below can over/underflow).
// t/d is then q2 of  q0 = RN(t y), r0 = t - q0 d (exact, FMA), q1 = RN(q0 + r0 y), r1 = t - q1 d, q2 = RN(q1 + r1 y):
// q1 is within 1 ulp of t/d (|t/d - (q0 + r0 y)| = |t/d - q0| |1 - d y| <= 1.5 ulp * 2^-53), and one Markstein correction
// step from a faithful q with y = RN(1/d) and an exact remainder yields the correctly rounded quotient (Markstein 1990;
// Muller et al., Handbook of Floating-Point Arithmetic, "Newton-Raphson-based division with an FMA") -- the same
// last step __ddiv_rn itself ends with.  5 dependent fp64 operations (~40 cycles) instead of MUFU.RCP64H + 9 (~200 measured).
// Outside the guarded range (|t| tiny/huge/zero/non-finite) it falls back to __ddiv_rn.  Checked bit for bit against
// __ddiv_rn on the device by tests/test_gpu_parity.py::test_quotient_fast_path (amgb200_debug_quotient_check).
__device__ __forceinline__ bool gs_quotient_dsafe(double d) {
    const double ad = fabs(d);
    const long long m = __double_as_longlong(d) & 0x000fffffffffffffLL;
    return ad > 0x1p-200 && ad < 0x1p200 && m != 0x000fffffffffffffLL;
}
// (a separate, never-inlined function: inlined, ptxas if-converts the guard and runs the whole IEEE division next to the
// fast path on every call)
__device__ __noinline__ double gs_quotient_slow(double t, double d) { return __ddiv_rn(t, d); }
__device__ __forceinline__ double gs_quotient_pre(double t, double d, double y, bool dsafe, int recip, bool need = true) {
    if (recip) return __dmul_rn(t, y);                                         // SSS_smooth.c:112-118: x = t * (1/d)
    const double q0 = __dmul_rn(t, y);
    const double r0 = __fma_rn(-q0, d, t);
    const double q1 = __fma_rn(r0, y, q0);
    const double r1 = __fma_rn(-q1, d, t);
    double q2 = __fma_rn(r1, y, q1);
    const double at = fabs(t);
    if (need && !(dsafe && at > 0x1p-700 && at < 0x1p700)) q2 = gs_quotient_slow(t, d);
    return q2;
}

// where the x vector of a level lives during a launch
template <bool COH>
struct GlobalX {                       // global memory (or this CTA's shared memory: generic addressing)
    double *x;
    __device__ __forceinline__ double ld(int j) const { return ld_x<COH>(x + j); }
    __device__ __forceinline__ void st(int k, double v) const { x[k] = v; }
};
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
// block-level deterministic tree sum of one value per thread; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *smem /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < (int)(blockDim.x >> 5) ? smem[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}
__device__ __forceinline__ double block_max(double v, double *smem) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < (int)(blockDim.x >> 5) ? smem[lane] : 0.0;
        v = warp_max(v);
    }
    return v;
}

// ==========================================================================================
// Thread-per-row work item: one SELL slice (<= 32 rows), lane r owns row slice_row[s]+r.
// prologue() only touches the (static) matrix, so it may run before a dependency wait.
// ==========================================================================================
template <int SCH>
struct SellItem {
    const int *cp;
    const double *vp;
    int width, k, r1, recip;
    int j[SCH];
    double a[SCH];
    double bk;
    __device__ __forceinline__ void load_chunk(int e0) {
#pragma unroll
        for (int u = 0; u < SCH; ++u) {
            if (e0 + u < width) { j[u] = cp[(size_t)(e0 + u) * 32]; a[u] = vp[(size_t)(e0 + u) * 32]; }
            else { j[u] = -1; a[u] = 0.0; }
        }
    }
    // stage 1: the slice descriptor (two dependent-load levels below the wavefront table)
    struct Desc { int r0, r1; long long p0, p1; };
    __device__ __forceinline__ static Desc load_desc(const DMat &A, int s) {
        Desc d;
        d.r0 = A.slice_row[s]; d.r1 = A.slice_row[s + 1];
        d.p0 = A.slice_ptr[s]; d.p1 = A.slice_ptr[s + 1];
        return d;
    }
    // stage 2: first chunk of matrix entries and the right-hand side
    __device__ __forceinline__ void load_entries(const DMat &A, const Desc &d, int lane, const double *__restrict__ b) {
        r1 = d.r1;
        recip = A.recip;
        width = (int)((d.p1 - d.p0) >> 5);
        k = d.r0 + lane;
        cp = A.col + d.p0 + lane;
        vp = A.val + d.p0 + lane;
        load_chunk(0);
        bk = (b && k < r1) ? b[k] : 0.0;
    }
    __device__ __forceinline__ void prologue(const DMat &A, int s, int lane, const double *__restrict__ b = nullptr) {
        load_entries(A, load_desc(A, s), lane, b);
    }
};

// Gauss-Seidel row update (amg/Solve/SSS_smooth.c:18-33): t = b_i - sum_{j != i} a_ij x_j in storage
// order; x_i = t / a_ii when |a_ii| > 1e-20
// rows no longer than one chunk (max row length <= SCH, e.g. level 0 of the 5-/7-point problems): no loop,
// no next-chunk registers -> ~40 registers, 6 blocks/SM for the HBM-bound kernels
// Two forms of the same arithmetic.  LAT = false (throughput-bound kernels: gs_pass, spmv, resid_restrict): a per-entry `if`, which skips the
// work of padding slots and measured 5-10 % more HBM throughput.  LAT = true (latency-bound ordered sweeps): branch-free -- ptxas turns every
// per-entry `if` into its own divergence region (BSSY / BRA / DMUL / DADD / BSYNC: ~125 cycles per entry ON the dependency chain, 2 700 cycles
// for a 20-slot row against 145 for the division; tools/ubench5.cu).  There all products are formed first (independent, pipelined), padding
// and the diagonal slot get the neutral element -- t - (+0.0) = t and t + (-0.0) = t bit for bit, for every t including the signed zeros --
// and the chain is SCH dependent DSUB / DADD and nothing else (8.1 cycles each).  Both forms give the same bits.
template <bool COH, int SCH, bool LAT = false>
__device__ __forceinline__ void gs_finish_sell_one(SellItem<SCH> &it, double *x) {
    double t = it.bk, d = 0.0;
    if constexpr (LAT) {
        double pr[SCH];
#pragma unroll
        for (int u = 0; u < SCH; ++u) {
            const bool off = it.j[u] >= 0 && it.j[u] != it.k;
            const double xv = off ? ld_x<COH>(x + it.j[u]) : 0.0;
            pr[u] = off ? __dmul_rn(it.a[u], xv) : 0.0;
            d = it.j[u] == it.k ? it.a[u] : d;
        }
#pragma unroll
        for (int u = 0; u < SCH; ++u) t = __dsub_rn(t, pr[u]);
    } else {
        double xv[SCH];
#pragma unroll
        for (int u = 0; u < SCH; ++u) xv[u] = (it.j[u] >= 0 && it.j[u] != it.k) ? ld_x<COH>(x + it.j[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < SCH; ++u) {
            if (it.j[u] == it.k) d = it.a[u];
            else if (it.j[u] >= 0) t = __dsub_rn(t, __dmul_rn(it.a[u], xv[u]));
        }
    }
    if (it.k < it.r1 && fabs(d) > GS_TINY) x[it.k] = gs_quotient(t, d, it.recip);
}
template <int SCH, bool COH = false, bool LAT = false>
__device__ __forceinline__ double spmv_finish_sell_one(SellItem<SCH> &it, const double *__restrict__ x) {
    double t = 0.0;
    if constexpr (LAT) {
        double pr[SCH];
#pragma unroll
        for (int u = 0; u < SCH; ++u) {
            const bool on = it.j[u] >= 0;
            const double xv = on ? ld_x<COH>(x + it.j[u]) : 0.0;
            pr[u] = on ? __dmul_rn(it.a[u], xv) : -0.0;
        }
#pragma unroll
        for (int u = 0; u < SCH; ++u) t = __dadd_rn(t, pr[u]);
    } else {
        double xv[SCH];
#pragma unroll
        for (int u = 0; u < SCH; ++u) xv[u] = it.j[u] >= 0 ? ld_x<COH>(x + it.j[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < SCH; ++u)
            if (it.j[u] >= 0) t = __dadd_rn(t, __dmul_rn(it.a[u], xv[u]));
    }
    return t;
}

template <bool COH, int SCH, bool LAT = false>
__device__ __forceinline__ void gs_finish_sell(SellItem<SCH> &it, double *x) {
    const bool active = it.k < it.r1;
    double t = it.bk, d = 0.0;
    for (int e0 = 0; e0 < it.width; e0 += SCH) {
        double xv[SCH];
#pragma unroll
        for (int u = 0; u < SCH; ++u) xv[u] = (it.j[u] >= 0 && it.j[u] != it.k) ? ld_x<COH>(x + it.j[u]) : 0.0;
        int jn[SCH];
        double an[SCH];
        const bool more = e0 + SCH < it.width;
        if (more) {                                   // next chunk's matrix entries fly during this chunk's chain
#pragma unroll
            for (int u = 0; u < SCH; ++u) {
                if (e0 + SCH + u < it.width) { jn[u] = it.cp[(size_t)(e0 + SCH + u) * 32]; an[u] = it.vp[(size_t)(e0 + SCH + u) * 32]; }
                else { jn[u] = -1; an[u] = 0.0; }
            }
        }
        if constexpr (LAT) {
#pragma unroll
            for (int u = 0; u < SCH; ++u) {
                const bool off = it.j[u] >= 0 && it.j[u] != it.k;
                d = it.j[u] == it.k ? it.a[u] : d;
                xv[u] = off ? __dmul_rn(it.a[u], xv[u]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < SCH; ++u) t = __dsub_rn(t, xv[u]);
        } else {
#pragma unroll
            for (int u = 0; u < SCH; ++u) {
                if (it.j[u] == it.k) d = it.a[u];
                else if (it.j[u] >= 0) t = __dsub_rn(t, __dmul_rn(it.a[u], xv[u]));
            }
        }
        if (more) {
#pragma unroll
            for (int u = 0; u < SCH; ++u) { it.j[u] = jn[u]; it.a[u] = an[u]; }
        }
    }
    if (active && fabs(d) > GS_TINY) x[it.k] = gs_quotient(t, d, it.recip);
}

// row sum t = sum_k a_k x_{j_k} from 0.0 in storage order (amg/SSS_utils.c:169-177, :190-200)
template <int SCH, bool COH = false, bool LAT = false>
__device__ __forceinline__ double spmv_finish_sell(SellItem<SCH> &it, const double *__restrict__ x) {
    double t = 0.0;
    for (int e0 = 0; e0 < it.width; e0 += SCH) {
        double xv[SCH];
#pragma unroll
        for (int u = 0; u < SCH; ++u) xv[u] = it.j[u] >= 0 ? ld_x<COH>(x + it.j[u]) : 0.0;
        int jn[SCH];
        double an[SCH];
        const bool more = e0 + SCH < it.width;
        if (more) {
#pragma unroll
            for (int u = 0; u < SCH; ++u) {
                if (e0 + SCH + u < it.width) { jn[u] = it.cp[(size_t)(e0 + SCH + u) * 32]; an[u] = it.vp[(size_t)(e0 + SCH + u) * 32]; }
                else { jn[u] = -1; an[u] = 0.0; }
            }
        }
        if constexpr (LAT) {
#pragma unroll
            for (int u = 0; u < SCH; ++u) xv[u] = it.j[u] >= 0 ? __dmul_rn(it.a[u], xv[u]) : -0.0;
#pragma unroll
            for (int u = 0; u < SCH; ++u) t = __dadd_rn(t, xv[u]);
        } else {
#pragma unroll
            for (int u = 0; u < SCH; ++u)
                if (it.j[u] >= 0) t = __dadd_rn(t, __dmul_rn(it.a[u], xv[u]));
        }
        if (more) {
#pragma unroll
            for (int u = 0; u < SCH; ++u) { it.j[u] = jn[u]; it.a[u] = an[u]; }
        }
    }
    return t;
}

// ==========================================================================================
// Warp-per-row work item (CSR).  The 32 lanes fetch and multiply 128 entries per step ("super
// chunk", 4 per lane, coalesced); the separately rounded products are staged in 1 KB of shared
// memory per warp and folded into the accumulator IN STORAGE ORDER by a dependent chain of
// DADD/DSUB fed with broadcast LDS.128 (8.4 cycles per term measured on B200 = the latency of a
// dependent fp64 add; shuffle-fed chains degrade to 16 cycles when several warps share an SM).
// Pipeline: col/val of super chunk s+1 and the x gather of s+1 are in flight during the chain of s.
// prologue() touches only the static matrix (and b) and may run before a dependency wait.
// ==========================================================================================
constexpr int SUPER = 128;
constexpr int STAGE = SUPER + 16;   // per-warp product staging area (the chain reads up to 16 slots ahead)
struct CsrItem {
    int k, p0, p1, ps;         // row, its entry range, and the first entry that reads the previous wavefront (p0 <= ps <= p1)
    int pb, pe;                // range currently being streamed
    int j[4], j1[4];           // super chunks 0 and 1 of that range
    double a[4], a1[4];
    unsigned lw[4], lw1[4];    // (two-phase rows) the late-flag words of those entries
    double bk, dl;             // right-hand side; diagonal entry if this lane has seen it
#ifdef AMGB200_TIMING
    long long tg = 0, tp = 0, tc = 0, nch = 0;
#endif
    __device__ __forceinline__ void load_super(const DMat &A, int base, int lane, int (&jj)[4], double (&aa)[4]) const {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = base + u * 32 + lane;
            if (p < pe) { jj[u] = A.col[p]; aa[u] = A.val[p]; } else { jj[u] = -1; aa[u] = 0.0; }
        }
    }
    __device__ __forceinline__ void load_late(const DMat &A, int base, int lane, unsigned (&ll)[4]) const {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = base + u * 32 + lane;
            ll[u] = p < pe ? A.late[p >> 5] : 0u;
        }
    }
    // the late-flag words of the first two super chunks of the suffix [ps, p1), requested early ...
    unsigned lwp[4], lwp1[4];
    __device__ __forceinline__ void suffix_prefetch_late(const DMat &A, int lane) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = ps + u * 32 + lane, q = p + SUPER;
            lwp[u] = p < p1 ? A.late[p >> 5] : 0u;
            lwp1[u] = q < p1 ? A.late[q >> 5] : 0u;
        }
    }
    // ... and adopted once the suffix range has been begun
    __device__ __forceinline__ void adopt_late() {
#pragma unroll
        for (int u = 0; u < 4; ++u) { lw[u] = lwp[u]; lw1[u] = lwp1[u]; }
    }
    struct Desc { int row, p0, p1, ps; };
    __device__ __forceinline__ static Desc load_desc(const DMat &A, int row) {
        Desc d;
        d.row = row; d.p0 = A.rptr[row]; d.p1 = A.rptr[row + 1];
        d.ps = A.split ? d.p0 + A.split[row] : d.p0;
        return d;
    }
    // start streaming the entries [beg, end): the first two super chunks of col/val
    __device__ __forceinline__ void begin_range(const DMat &A, int beg, int end, int lane) {
        pb = beg; pe = end;
        load_super(A, beg, lane, j, a);
        load_super(A, beg + SUPER, lane, j1, a1);
    }
    __device__ __forceinline__ void load_entries(const DMat &A, const Desc &d, int lane, const double *__restrict__ b, bool prefix_only = false) {
        k = d.row; p0 = d.p0; p1 = d.p1; ps = d.ps; dl = 0.0;
        begin_range(A, p0, prefix_only ? ps : p1, lane);
        bk = b ? b[d.row] : 0.0;
    }
    __device__ __forceinline__ void prologue(const DMat &A, int row, int lane, const double *__restrict__ b) {
        load_entries(A, load_desc(A, row), lane, b);
    }
};

// pull the col/val (and late-flag) lines of the rows [i0, i1) towards this SM's L1: the warps of a group do this for
// the wavefront after their next one, so that the dependent loads of fetch() and the chunk stream of the rows hit
// L1 (~70 cycles) instead of L2 (~450) when their turn comes
__device__ __forceinline__ void csr_prefetch_rows(const DMat &A, int i0, int i1, int lane, int worker, int nworkers) {
    for (int row = i0 + worker; row < i1; row += nworkers) {
        const int p0 = A.rptr[row], p1 = A.rptr[row + 1];
        for (int p = p0 + lane * 16; p < p1; p += 32 * 16) {          // one touch per 128-byte line of val, every other line of col
            asm volatile("prefetch.global.L1 [%0];" ::"l"(A.val + p));
            if ((lane & 1) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.col + p));
        }
    }
}

// fold `cnt` staged products into t in order: 8-term blocks, ping-pong registers; the LDS.128 of the next
// block are issued before the chain of the current one, and the __syncwarp between them keeps ptxas from
// sinking the loads next to their uses (which would expose ~30 cycles of shared-memory latency every few
// terms).  Four blocks per loop trip: one taken branch per 32 terms.  Slots >= cnt hold +0.0 (exact no-ops);
// reads run up to 16 slots past the staged chunk (STAGE pad).
__device__ __forceinline__ double2 lds_v2f64(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
template <bool GS>
__device__ __forceinline__ double chain_fold(double t, const double2 *sp2, int cnt) {
#if defined(AMGB200_ABLATE) && AMGB200_ABLATE == 1
    return t + sp2[0].x * 1e-300;      // timing experiment only: no chain
#endif
#define AMGB200_FOLD(v)                                                        \
    _Pragma("unroll") for (int u = 0; u < 4; ++u) {                            \
        if (GS) { t = __dsub_rn(t, v[u].x); t = __dsub_rn(t, v[u].y); }        \
        else { t = __dadd_rn(t, v[u].x); t = __dadd_rn(t, v[u].y); }           \
    }
    // volatile-asm shared loads stay in program order and ahead of the chain; a __syncwarp fence here turns into a
    // BRA.DIV per block wherever the compiler cannot prove the warp converged (+1.9 cycles per term, tools/ubench2.cu)
#define AMGB200_LOAD(v, q)                                                     \
    _Pragma("unroll") for (int u = 0; u < 4; ++u) v[u] = lds_v2f64(sp + 8u * (unsigned)(q) + 16u * u);
    const unsigned sp = (unsigned)__cvta_generic_to_shared(sp2);
    double2 va[4], vb[4];
    int q = 0;
    AMGB200_LOAD(va, 0)
#pragma unroll 1
    for (; q + 32 <= cnt; q += 32) {
        AMGB200_LOAD(vb, q + 8)  AMGB200_FOLD(va)
        AMGB200_LOAD(va, q + 16) AMGB200_FOLD(vb)
        AMGB200_LOAD(vb, q + 24) AMGB200_FOLD(va)
        AMGB200_LOAD(va, q + 32) AMGB200_FOLD(vb)
    }
#pragma unroll 1
    for (; q < cnt; q += 8) {
        AMGB200_LOAD(vb, q + 8)  AMGB200_FOLD(va)
#pragma unroll
        for (int u = 0; u < 4; ++u) va[u] = vb[u];
    }
#undef AMGB200_FOLD
#undef AMGB200_LOAD
    return t;
}

// EXACT in-order accumulation over the range begun with begin_range().  GS: products are subtracted and the
// diagonal entry is skipped (remembered in it.dl); otherwise products are added.  Padding and the skipped
// diagonal contribute +0.0, which leaves t bit-unchanged.  Result valid in all lanes.
template <class XA, bool GS>
__device__ __forceinline__ double csr_chain_run_x(const DMat &A, CsrItem &it, const XA &xa, double t, int lane, double *sprod) {
    if (it.pb >= it.pe) return t;
    double xc[4];
#ifdef AMGB200_TIMING
    long long k0 = clock64();
#endif
#pragma unroll
    for (int u = 0; u < 4; ++u) xc[u] = (it.j[u] >= 0 && !(GS && it.j[u] == it.k)) ? xa.ld(it.j[u]) : 0.0;
#ifdef AMGB200_TIMING
    { double sink = xc[0] + xc[1] + xc[2] + xc[3]; if (sink == 1.2345e300) it.k = -1; it.tg += clock64() - k0; }
#endif
    for (int base = it.pb; base < it.pe; base += SUPER) {
#ifdef AMGB200_TIMING
        long long k2 = clock64();
#endif
        // stage 1: col/val of super chunk s+2 ; stage 2: x gather of s+1 (its col arrived an iteration ago)
        int jn[4];
        double an[4], xn[4];
        const bool more = base + SUPER < it.pe;
        if (base + 2 * SUPER < it.pe) it.load_super(A, base + 2 * SUPER, lane, jn, an);
        else {
#pragma unroll
            for (int u = 0; u < 4; ++u) { jn[u] = -1; an[u] = 0.0; }
        }
        if (more) {
#pragma unroll
            for (int u = 0; u < 4; ++u) xn[u] = (it.j1[u] >= 0 && !(GS && it.j1[u] == it.k)) ? xa.ld(it.j1[u]) : 0.0;
        }
        // stage 3: products of super chunk s into the staging buffer, then the in-order chain
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double prod = 0.0;
            if (it.j[u] >= 0) {
                if (GS && it.j[u] == it.k) it.dl = it.a[u];
                else prod = __dmul_rn(it.a[u], xc[u]);
            }
            sprod[u * 32 + lane] = prod;
        }
        __syncwarp();
#ifdef AMGB200_TIMING
        long long k3 = clock64(); it.tp += k3 - k2; ++it.nch;
#endif
        t = chain_fold<GS>(t, reinterpret_cast<const double2 *>(sprod), min(SUPER, it.pe - base));
#ifdef AMGB200_TIMING
        { if (t == 1.2345e300) it.k = -1; it.tc += clock64() - k3; }
#endif
        __syncwarp();
        if (more) {
#pragma unroll
            for (int u = 0; u < 4; ++u) { it.j[u] = it.j1[u]; it.a[u] = it.a1[u]; xc[u] = xn[u]; it.j1[u] = jn[u]; it.a1[u] = an[u]; }
        }
    }
    return t;
}
template <bool COH, bool GS>
__device__ __forceinline__ double csr_chain_run(const DMat &A, CsrItem &it, const double *x, double t, int lane, double *sprod) {
    return csr_chain_run_x<GlobalX<COH>, GS>(A, it, GlobalX<COH>{const_cast<double *>(x)}, t, lane, sprod);
}
__device__ __forceinline__ double csr_diag(const CsrItem &it) {   // exactly one lane saw the diagonal entry: broadcast it
    const unsigned m = __ballot_sync(FULL, it.dl != 0.0);
    return m ? __shfl_sync(FULL, it.dl, __ffs(m) - 1) : 0.0;
}
// Pre-barrier half of the two-phase Gauss-Seidel row: compute the (separately rounded) products of the suffix
// entries [it.pb, it.pe) with the x values that are already final and park them, in storage order, in `big`.
// Entries that read the wavefront still in flight ("late", flagged by the host analysis) get a +0.0 placeholder
// and are appended to the warp's late list {slot, column, value}; after the barrier only those few products are
// recomputed before the in-order chain runs over `big`.  Returns the number of late entries (warp-uniform),
// or -1 if the list overflowed (the caller then falls back to streaming the suffix after the barrier).
constexpr int LATE_CAP = 32;
// one super chunk of the staging loop: `cur` holds the chunk at `base`, the loads of the chunk two ahead go into
// the free buffer.  The three buffers change roles at compile time (see csr_stage_suffix): a register that a load
// in flight will write is never moved, so nothing here waits for the loads this step issues.
template <class XA>
__device__ __forceinline__ void csr_stage_step(const DMat &A, CsrItem &it, const XA &xa, int lane, double *big, int *late_slot,
                                               int *late_col, double *late_val, int base, int &nlate,
                                               const int (&cj)[4], const double (&ca)[4], const unsigned (&cl)[4],
                                               int (&fj)[4], double (&fa)[4], unsigned (&fl)[4]) {
    it.load_super(A, base + 2 * SUPER, lane, fj, fa);
    it.load_late(A, base + 2 * SUPER, lane, fl);
    // all four gathers in flight together, no divergent region between them; the (rare) late entries are
    // collected afterwards
    bool late[4];
    double xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int p = base + u * 32 + lane;
        const int j = cj[u];
        late[u] = j >= 0 && ((cl[u] >> (p & 31)) & 1u);
        const bool use = j >= 0 && j != it.k && !late[u];
        xv[u] = use ? xa.ld(j) : 0.0;
        if (j >= 0 && j == it.k) it.dl = ca[u];
    }
    unsigned any = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int p = base + u * 32 + lane;
        const int j = cj[u];
        const bool use = j >= 0 && j != it.k && !late[u];
        if (p < it.pe) big[p - it.pb] = use ? __dmul_rn(ca[u], xv[u]) : 0.0;
        any |= late[u] ? 1u << u : 0u;
    }
    if (__any_sync(FULL, any != 0u)) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned m = __ballot_sync(FULL, late[u]);
            if (m) {
                const int slot = nlate + __popc(m & ((1u << lane) - 1u));
                if (late[u] && slot < LATE_CAP) { late_slot[slot] = base + u * 32 + lane - it.pb; late_col[slot] = cj[u]; late_val[slot] = ca[u]; }
                nlate += __popc(m);
            }
        }
    }
}
template <class XA>
__device__ __forceinline__ int csr_stage_suffix(const DMat &A, CsrItem &it, const XA &xa, int lane, double *big,
                                               int *late_slot, int *late_col, double *late_val) {
    int nlate = 0;
    const int len = it.pe - it.pb;
    if (len > 0) {
        int j2[4];
        double a2[4];
        unsigned lw2[4];
        int base = it.pb;
#pragma unroll 1
        for (;;) {
            csr_stage_step(A, it, xa, lane, big, late_slot, late_col, late_val, base, nlate, it.j, it.a, it.lw, j2, a2, lw2);
            base += SUPER; if (base >= it.pe) break;
            csr_stage_step(A, it, xa, lane, big, late_slot, late_col, late_val, base, nlate, it.j1, it.a1, it.lw1, it.j, it.a, it.lw);
            base += SUPER; if (base >= it.pe) break;
            csr_stage_step(A, it, xa, lane, big, late_slot, late_col, late_val, base, nlate, j2, a2, lw2, it.j1, it.a1, it.lw1);
            base += SUPER; if (base >= it.pe) break;
        }
    }
    if (lane < 24) big[len + lane] = 0.0;          // the chain reads in blocks of 8 and up to 16 slots ahead
    __syncwarp();
    return nlate <= LATE_CAP ? nlate : -1;
}

template <bool COH, bool GS>
__device__ __forceinline__ double csr_row_exact(const DMat &A, CsrItem &it, const double *x, double t, double &d, int lane, double *sprod) {
    t = csr_chain_run<COH, GS>(A, it, x, t, lane, sprod);
    d = csr_diag(it);
    return t;
}

// FAST variant: per-lane partial sums combined by a shuffle tree (not the reference's rounding order)
template <bool COH, bool GS>
__device__ __forceinline__ double csr_row_fast(const DMat &A, CsrItem &it, const double *x, double t, double &d, int lane) {
    double dl = 0.0, acc = 0.0;
    for (int base = it.pb; base < it.pe; base += SUPER) {
        if (base != it.pb) it.load_super(A, base, lane, it.j, it.a);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (it.j[u] >= 0) {
                if (GS && it.j[u] == it.k) dl = it.a[u];
                else acc = __dadd_rn(acc, __dmul_rn(it.a[u], ld_x<COH>(x + it.j[u])));
            }
        }
    }
    acc = warp_sum(acc);
    t = GS ? __dsub_rn(t, acc) : __dadd_rn(t, acc);
    d = warp_sum(dl);
    return t;
}

template <bool COH, bool EXACT>
__device__ __forceinline__ void gs_finish_csr(const DMat &A, CsrItem &it, double *x, int lane, double *sprod) {
    double d;
    const double t = EXACT ? csr_row_exact<COH, true>(A, it, x, it.bk, d, lane, sprod) : csr_row_fast<COH, true>(A, it, x, it.bk, d, lane);
    if (lane == 0 && fabs(d) > GS_TINY) x[it.k] = gs_quotient(t, d, A.recip);
}

// ------------------------------------------------------------------------------------------
// Gauss-Seidel kernels
// ------------------------------------------------------------------------------------------
// one fully parallel pass (a pass whose dependency DAG has depth 1): items [item0, item1)
template <int KIND, bool EXACT, bool ONE = false>
__global__ void __launch_bounds__(BLOCK) gs_pass_kernel(DMat A, const double *__restrict__ b, double *x, int item0, int item1) {
    __shared__ double sprod[KIND == 1 ? WARPS_PER_BLOCK * STAGE : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int it = item0 + blockIdx.x * WARPS_PER_BLOCK + warp;
    if (it >= item1) return;
    if constexpr (KIND == 0) {
        SellItem<8> w;
        w.prologue(A, it, lane, b);
        if constexpr (ONE) gs_finish_sell_one<false>(w, x); else gs_finish_sell<false>(w, x);
    } else {
        CsrItem w;
        w.prologue(A, it, lane, b);
        gs_finish_csr<false, EXACT>(A, w, x, lane, sprod + warp * STAGE);
    }
}

// Ordered sweeps inside ONE thread block (levels whose wavefronts are narrow).  The block's G*D
// warps form D groups of G warps; group (g mod D) owns wavefront g: its warps take the wavefront's
// items round-robin.  A group fetches the matrix entries (row pointers, col, val, b) of its next
// wavefront D steps ahead, so only the x gathers, the in-order chain and one __syncthreads per
// wavefront sit on the dependency path.  No inter-SM traffic.
// XS: the whole x vector lives in shared memory for the duration of the launch (levels with
// n*8 bytes <= ~200 KB): the gathers cost ~30 cycles instead of an L2 round trip.
// Dynamic shared memory: [x (n doubles, XS only)] [nwarps * 128 doubles of product staging]
constexpr int CTA_MAX_WARPS_SELL = 8, CTA_MAX_WARPS_CSR = 16;
template <int KIND, bool EXACT, bool XS>
__global__ void __launch_bounds__(KIND == 0 ? 32 * CTA_MAX_WARPS_SELL : 32 * CTA_MAX_WARPS_CSR) gs_ordered_cta_kernel(
    DMat A, const double *__restrict__ b, double *xg, const int *__restrict__ wf_item_ptr, int W, int nsweeps, int G, int D,
    int cap, long long *dbg) {
    extern __shared__ double dyn_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp % D, r = warp / D;      // row r of both groups on neighbouring warps = different SM sub-partitions (fp64 pipe, see the cluster kernel)
    const int n = A.nrows;
    double *x = XS ? dyn_smem : xg;
    // per warp: [STAGE streaming slots][cap + 24 parked suffix products][LATE_CAP values][LATE_CAP slots + LATE_CAP columns]
    const int per_warp = STAGE + (cap ? cap + 24 + LATE_CAP + LATE_CAP : 0);
    double *sprod = dyn_smem + (XS ? ((n + 1) & ~1) : 0) + warp * per_warp;
    double *big = sprod + STAGE;
    double *late_val = big + cap + 24;
    int *late_slot = reinterpret_cast<int *>(late_val + LATE_CAP);
    int *late_col = late_slot + LATE_CAP;
    int nlate = -1;                               // >= 0: the suffix of my row is parked in `big`
    if (XS) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) dyn_smem[i] = xg[i];
        __syncthreads();
    }
    const int totalw = W * nsweeps;
    SellItem<20> ws;
    CsrItem wc;
    int i0 = 0, i1 = 0;
    bool have = false;
    int my_g = grp, my_wl = grp;                  // next wavefront of this warp's group (global index, index within a sweep)
    while (my_wl >= W) my_wl -= W;
    // Two-phase rows (warp-per-row, EXACT): the entries of a row that come before its first entry reading the
    // wavefront in flight ("prefix", ~30 % of a row on Galerkin operators) only touch x values that are already
    // final, so the waiting group folds them into the accumulator BEFORE the barrier; only the suffix chain
    // is on the dependency path.  Storage order is preserved: prefix then suffix.
    constexpr bool TWO_PHASE = KIND == 1 && EXACT;
    double t_acc = 0.0;
    auto fetch = [&]() {
        i0 = wf_item_ptr[my_wl]; i1 = wf_item_ptr[my_wl + 1];
        have = i0 + r < i1;
        if (have) {
            if constexpr (KIND == 0) ws.prologue(A, i0 + r, lane, b);
            else wc.load_entries(A, CsrItem::load_desc(A, i0 + r), lane, b, TWO_PHASE);
        }
    };
    // Wavefront g is produced by group g mod D and consumed (waited for) by group (g+1) mod D through the
    // named barrier 1 + (g mod 8): the producers only *arrive* (non-blocking) and go on to prefetch their
    // next wavefront, so the prefetch latency never delays the consumers.  D >= 2.  Barrier 9 + grp keeps
    // the warps of one group together before they start reading x for their next wavefront's prefixes.
    const int pair_threads = 2 * G * 32;
    if (my_g < totalw) fetch();
#ifdef AMGB200_TIMING
    long long tm_prefix = 0, tm_wait = 0, tm_suffix = 0, tm_post = 0, tm_items = 0, tm_sg = 0, tm_sp = 0, tm_sc = 0, tm_sn = 0;
#endif
#ifdef AMGB200_TIMELINE
    long long tl[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define TL_FORCE(v) { if ((v) == -12345) t_acc = 0; }
#define TL_MARK(i) { volatile double *vs_ = sprod; if (vs_[0] == 1.2345e300) t_acc = 0; const long long c_ = clock64(); tl[i] += c_ - tl_prev; tl_prev = c_; }
    long long tl_prev = clock64();
#else
#define TL_MARK(i)
#define TL_FORCE(v)
#endif
    for (; my_g < totalw; my_g += D) {
#ifdef AMGB200_TIMING
        const long long q0 = clock64();
#endif
        if constexpr (TWO_PHASE) {
            if (have) {
                const bool park = cap && wc.p1 - wc.ps <= cap && A.late;
                if (park) wc.suffix_prefetch_late(A, lane);                               // late words of the suffix fly during the prefix chain
                TL_FORCE(wc.j[0]) TL_MARK(10)
                t_acc = csr_chain_run<false, true>(A, wc, x, wc.bk, lane, sprod);     // prefix [p0, ps)
                TL_MARK(0)
                nlate = -1;
                if (park) {
                    wc.begin_range(A, wc.ps, wc.p1, lane);
                    wc.adopt_late();
                    TL_MARK(11) TL_FORCE(wc.j[0]) TL_MARK(12)
                    nlate = csr_stage_suffix(A, wc, GlobalX<false>{x}, lane, big, late_slot, late_col, late_val);
                }
                if (nlate < 0) wc.begin_range(A, wc.ps, wc.p1, lane);                 // stream the suffix after the barrier
                TL_MARK(1)
            }
        }
#ifdef AMGB200_TIMING
        if (t_acc == 1.2345e300) ++tm_items;
        const long long q1 = clock64();
#endif
        if (my_g > 0) asm volatile("bar.sync %0, %1;" ::"r"(1 + ((my_g - 1) & 7)), "r"(pair_threads) : "memory");
        TL_MARK(2)
#ifdef AMGB200_TIMING
        { volatile double *vs = sprod; if (vs[0] == 1.2345e300) ++tm_items; }
        const long long q2 = clock64();
        if (have) ++tm_items;
#endif
        if (have) {
            if constexpr (KIND == 0) gs_finish_sell<false, 20, true>(ws, x);
            else if constexpr (TWO_PHASE) {
                double t;
                if (nlate >= 0) {                 // patch the few products that needed the wavefront just completed, then fold
                    if (lane < nlate) big[late_slot[lane]] = __dmul_rn(late_val[lane], x[late_col[lane]]);
                    __syncwarp();
                    TL_MARK(3)
                    t = chain_fold<true>(t_acc, reinterpret_cast<const double2 *>(big), wc.p1 - wc.ps);
                } else t = csr_chain_run<false, true>(A, wc, x, t_acc, lane, sprod);        // suffix [ps, p1), streamed
#ifdef AMGB200_TIMELINE
                if (t == 1.2345e300) t_acc = 1;
#endif
                TL_MARK(4)
                const double d = csr_diag(wc);
                if (lane == 0 && fabs(d) > GS_TINY) x[wc.k] = gs_quotient(t, d, A.recip);
                TL_MARK(5)
            } else gs_finish_csr<false, EXACT>(A, wc, x, lane, sprod);
            for (int it = i0 + r + G; it < i1; it += G) {              // wavefront wider than the group
                if constexpr (KIND == 0) { ws.prologue(A, it, lane, b); gs_finish_sell<false, 20, true>(ws, x); }
                else { wc.prologue(A, it, lane, b); gs_finish_csr<false, EXACT>(A, wc, x, lane, sprod); }
            }
        }
        // bar.arrive orders this thread's prior shared-memory stores before the consumers' bar.sync (PTX ISA, bar:
        // producer/consumer example); x in global memory additionally needs the CTA-scope fence
        if (!XS) __threadfence_block();
        if (my_g + 1 < totalw) asm volatile("bar.arrive %0, %1;" ::"r"(1 + (my_g & 7)), "r"(pair_threads) : "memory");
        TL_MARK(6)
#ifdef AMGB200_TIMING
        const long long q3 = clock64();
#endif
        my_wl += D;
        while (my_wl >= W) my_wl -= W;
        if (my_g + D < totalw) {
            if constexpr (TWO_PHASE) asm volatile("bar.sync %0, %1;" ::"r"(9 + grp), "r"(G * 32) : "memory");   // whole group done with my_g
            TL_MARK(7)
            fetch();
            TL_MARK(8)
            if constexpr (KIND == 1) {
                if (my_g + 2 * D < totalw) {                          // warm L1 for the wavefront after that one
                    int wl2 = my_wl + D;
                    while (wl2 >= W) wl2 -= W;
                    csr_prefetch_rows(A, wf_item_ptr[wl2], wf_item_ptr[wl2 + 1], lane, r, G);
                }
            }
            TL_MARK(9)
        }
#ifdef AMGB200_TIMING
        { volatile double *vs = sprod; if (vs[0] == 1.2345e300) ++tm_items; }
        const long long q4 = clock64();
        tm_prefix += q1 - q0; tm_wait += q2 - q1; tm_suffix += q3 - q2; tm_post += q4 - q3;
#endif
    }
#ifdef AMGB200_TIMELINE
    if (dbg && lane == 0 && warp < 2) for (int i = 0; i < 14; ++i) dbg[(warp ? 16 : 0) + i] = tl[i];
#endif
#ifdef AMGB200_TIMING
    if (dbg && lane == 0) { dbg[warp * 8 + 0] = tm_prefix; dbg[warp * 8 + 1] = tm_wait; dbg[warp * 8 + 2] = tm_suffix; dbg[warp * 8 + 3] = tm_post; dbg[warp * 8 + 4] = tm_items; dbg[warp * 8 + 5] = tm_sg; dbg[warp * 8 + 6] = tm_sp; dbg[warp * 8 + 7] = tm_sc; if (warp == 0) dbg[16 * 8] = tm_sn; }
#endif
    __syncthreads();
    if (XS) for (int i = threadIdx.x; i < n; i += blockDim.x) xg[i] = dyn_smem[i];
}

// ------------------------------------------------------------------------------------------
// Streaming single-CTA ordered sweeps (warp-per-row EXACT rows, x in shared memory).
//
// The matrix of the level is pre-packed on the host into one self-contained block per wavefront
// (analysis.h, StreamLayout).  A LOADER warp walks the static schedule ahead of everybody else and pulls each
// block -- plus the right-hand-side entries of its rows -- from L2 into a byte ring in shared memory with
// cp.async.bulk (one elected thread, completion on an mbarrier): no consumer ever waits for an L2 round trip.
// 2*G CONSUMER warps in two alternating groups walk the wavefronts: while group A finishes wavefront g-1,
// group B (wavefront g) multiplies all its rows' entries by x in place (the values become the separately
// rounded products, in storage order) and folds each row's prefix -- the entries before the first one that
// reads wavefront g-1 -- into the accumulator.  After the named barrier that announces g-1, B only has to
// recompute the few "late" products, fold the suffix, divide and store: that chain is the whole critical path.
// Dynamic shared memory: [mbarriers full[NS], empty[NS] | ring offsets | 64 B of zeros : 256 B][x : n doubles][ring]
// ------------------------------------------------------------------------------------------
constexpr int STREAM_NS = 8;                  // wavefront blocks in flight (ring descriptors)
#ifndef AMGB200_PROD_U
#define AMGB200_PROD_U 4
#endif
constexpr int PROD_U = AMGB200_PROD_U;         // entries per lane per round of the product pass (3x as many shared-memory requests in flight per warp;
                                              // measured 2/4/8: 4 disturbs the folding warps of the other wavefronts least: level 4 2.58 -> 2.49 ms per sweep)
constexpr int STREAM_HDR = 384;               // bytes of barriers / descriptors in front of x
constexpr int STREAM_MAX_WARPS = 13;          // D groups of G warps + the loader (D*G <= 12: up to 157 registers per thread)
constexpr int STREAM_MAX_G = 4;               // product warps per group (1, 2 or 4)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// busy polling (test_wait never suspends the thread): lowest wake-up latency, for the waits on the dependency path
__device__ __forceinline__ void mbar_wait_spin(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_wait_sleep(unsigned bar, unsigned parity) {
    unsigned ok;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(400);
    }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_s32(unsigned a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
struct StreamLateDev { double val; int pos; int col; };

// chain_fold for a warp whose lanes are split into row slots: every lane folds the `cnt` (multiple of 8) products at
// sp2 of ITS slot's row; slots with fewer terms than the warp-wide maximum `maxc` idle on exact no-ops (t - (+0.0)).
// One warp-wide DSUB then advances up to 32/sub rows at once: the fp64 pipe (4 cycles per warp instruction per
// sub-partition) is shared by a quarter as many chaining warps.  Nothing beyond a slot's own cnt is read.
// sp / zeros are shared-space addresses.  The loads are volatile asm in program order, two blocks (16 terms) ahead of the
// chain; no __syncwarp in the loop (inside a branch the compiler cannot prove uniform it becomes a BRA.DIV per block,
// measured +1.9 cycles per term).
__device__ __forceinline__ double chain_fold_slots(double t, unsigned sp, int cnt, int maxc, unsigned zeros) {
#define AMGB200_FOLD(v)                                                        \
    _Pragma("unroll") for (int u = 0; u < 4; ++u) { t = __dsub_rn(t, v[u].x); t = __dsub_rn(t, v[u].y); }
    // blocks past this slot's own cnt are read from a 64-byte block of zeros: no branch, no select on the chain
#define AMGB200_LOAD(v, q)                                                     \
    {                                                                          \
        const unsigned src_ = (q) < cnt ? sp + 8u * (unsigned)(q) : zeros;     \
        _Pragma("unroll") for (int u = 0; u < 4; ++u) v[u] = lds_v2f64(src_ + 16u * u); \
    }
    double2 v0[4], v1[4], v2[4];
    int q = 0;
    AMGB200_LOAD(v0, 0)
    AMGB200_LOAD(v1, 8)
#pragma unroll 1
    for (; q + 24 <= maxc; q += 24) {
        AMGB200_LOAD(v2, q + 16) AMGB200_FOLD(v0)
        AMGB200_LOAD(v0, q + 24) AMGB200_FOLD(v1)
        AMGB200_LOAD(v1, q + 32) AMGB200_FOLD(v2)
    }
    // at most two 8-term blocks are left, and both are already in registers (v0 = block q, v1 = block q + 8): no loads, no moves
    if (q < maxc) { AMGB200_FOLD(v0) }
    if (q + 8 < maxc) { AMGB200_FOLD(v1) }
#undef AMGB200_FOLD
#undef AMGB200_LOAD
    return t;
}

__global__ void __launch_bounds__(32 * STREAM_MAX_WARPS) gs_stream_cta_kernel(
    const unsigned char *__restrict__ stream, const int *__restrict__ blk_ptr, const int *__restrict__ wf_row_ptr,
    const double *__restrict__ b, double *xg, int n, int W, int nsweeps, int G, int S, int D, int ring_bytes, int recip, long long *dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);      // full[0..NS), empty[NS..2NS); done[8] at +256
    volatile int *stage_off = reinterpret_cast<volatile int *>(smem_raw + 192);
    const unsigned zeros_a = smem_u32(smem_raw + 128);
    const unsigned done0 = smem_u32(smem_raw + 256);
    double *x = reinterpret_cast<double *>(smem_raw + STREAM_HDR);
    unsigned char *ring = smem_raw + STREAM_HDR + (((size_t)n * 8 + 127) & ~(size_t)127);       // (128-byte aligned: see analysis.cpp, bank alignment)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int totalw = W * nsweeps;
    const unsigned x_a = smem_u32(x);
#ifdef AMGB200_TIMELINE
    long long tl[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tl_prev = clock64();
#if AMGB200_TIMELINE == 2          // only two clock reads per wavefront: [8] = done(g-1) observed -> done(g) announced, [3] = everything else
#define SL_MARK(i) { if ((i) == 3 || (i) == 8) { const long long c_ = clock64(); tl[i] += c_ - tl_prev; tl_prev = c_; } }
#elif AMGB200_TIMELINE == 3        // post path in three pieces: [4] late patch, [6] suffix chain + quotient + store, [8] announce; [3] = everything else
#define SL_MARK(i) { if ((i) == 3 || (i) == 4 || (i) == 6 || (i) == 8) { const long long c_ = clock64(); tl[i] += c_ - tl_prev; tl_prev = c_; } }
#else
#define SL_MARK(i) { const long long c_ = clock64(); tl[i] += c_ - tl_prev; tl_prev = c_; }
#endif
#else
#define SL_MARK(i)
#endif
    if (threadIdx.x == 0) {
        for (int s = 0; s < STREAM_NS; ++s) { mbar_init(smem_u32(bars + s), 1); mbar_init(smem_u32(bars + STREAM_NS + s), 1); mbar_init(done0 + 8u * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 8) reinterpret_cast<double *>(smem_raw + 128)[threadIdx.x] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = xg[i];
    __syncthreads();
    if (warp < D * G) {
        // warps [d*G, (d+1)*G) = group d; group d walks the wavefronts g = d, d+D, ...  All G warps of a group multiply (one
        // row at a time, rows dealt round-robin); ONE warp per group folds: its lanes are split into S row slots, so a single
        // warp-wide DSUB advances every row of the wavefront, and the folding warps sit on different SM sub-partitions (a
        // warp-wide DADD occupies a sub-partition's fp64 pipe for 4 cycles: two chains on one pipe slow each other down,
        // measured 8.6 -> 11 cycles per term).  G is 1, 2 or 4 (8 with two groups).
        //
        // D wavefronts are in flight.  x_k of wavefront g is announced on the mbarrier done[g & 7].  With D = 3 the product
        // pass of wavefront g starts when g-3 is complete: entries whose column lies in wavefront g-2 ("late2", host-flagged) are
        // redone after done(g-2), before the prefix fold; entries in g-1 ("late") after done(g-1), before the suffix fold.
        // The pre-barrier work of a wavefront (products + prefix fold) thereby has TWO post-barrier periods to hide in.
        const int grp = warp / G, r = warp - grp * G;
        const bool folder = r == (G > 1 ? grp % G : 0);
        const int sub = 32 / S, slot = lane / sub;
        const int lis = lane - slot * sub;       // lane in slot
        const bool leader = lis == 0;
        for (int g = grp; g < totalw; g += D) {
            const int s = g & (STREAM_NS - 1);
            SL_MARK(7)
            mbar_wait(smem_u32(bars + s), (g / STREAM_NS) & 1);
            SL_MARK(0)
            unsigned char *blk = ring + stage_off[s];
            const int4 hd = *reinterpret_cast<const int4 *>(blk);                   // nrows, first row & ~1, rhs count, block bytes
            const int *rec_off = reinterpret_cast<const int *>(blk + 16);
            const double *bseg = reinterpret_cast<const double *>(blk + hd.w);
            // ---- products of every entry (late / late2 ones are redone below) ...
#if defined(AMGB200_ABLATE) && (AMGB200_ABLATE == 13 || AMGB200_ABLATE == 16)
            for (int ri = r; ri < 0; ri += G) {
#else
            for (int ri = r; ri < hd.x; ri += G) {                                   // all 32 lanes on one row at a time
#endif
                unsigned char *rec = blk + rec_off[ri];
                const int len_pad = reinterpret_cast<const int *>(rec)[2];
                // explicit shared-space accesses in program order (volatile asm): 8 column loads, then the 16 value /
                // x loads they feed, then the products -- ptxas otherwise sinks every value load next to its multiply
                const unsigned val_a = smem_u32(rec + 32), col_a = val_a + 8u * (unsigned)len_pad;
                for (int p0 = 0; p0 < len_pad; p0 += 32 * PROD_U) {
                    int j[PROD_U];
                    double v[PROD_U], xv[PROD_U];
#pragma unroll
                    for (int u = 0; u < PROD_U; ++u) { const int p = p0 + u * 32 + lane; j[u] = lds_s32(col_a + 4u * (unsigned)(p < len_pad ? p : 0)); if (p >= len_pad) j[u] = -1; }
#pragma unroll
                    for (int u = 0; u < PROD_U; ++u) {
                        const int p = p0 + u * 32 + lane;
                        v[u] = lds_f64(val_a + 8u * (unsigned)(p < len_pad ? p : 0));
#if defined(AMGB200_ABLATE) && AMGB200_ABLATE == 17
                        xv[u] = lds_f64(x_a + 8u * (unsigned)(j[u] >= 0 ? lane : 0));       // (timing only: conflict-free stand-in for the x gather)
#else
                        xv[u] = lds_f64(x_a + 8u * (unsigned)max(j[u], 0));
#endif
                    }
#pragma unroll
                    for (int u = 0; u < PROD_U; ++u) { const int p = p0 + u * 32 + lane; if (j[u] >= 0) sts_f64(val_a + 8u * (unsigned)p, __dmul_rn(v[u], xv[u])); }
                }
            }
            SL_MARK(1)
            if (G > 1) asm volatile("bar.sync %0, %1;" ::"r"(9 + grp), "r"(G * 32) : "memory");      // all products of the wavefront are in place
            else __syncwarp();
            if (folder) {
                // ---- after done(g-2): the late2 products of every row
                if (D > 2) {
                    if (g >= 2) mbar_wait(done0 + 8u * (unsigned)((g - 2) & 7), ((g - 2) >> 3) & 1);
                    SL_MARK(9)
                    for (int base = 0; base < hd.x; base += S) {
                        const int ri = base + slot;
                        if (ri < hd.x) {
                            unsigned char *rec = blk + rec_off[ri];
                            const int4 rh = *reinterpret_cast<const int4 *>(rec);
                            const unsigned val_a = smem_u32(rec + 32), l2_a = val_a + 12u * (unsigned)rh.z + 16u * (unsigned)(rh.w & 0xffff);
                            const int n2 = rh.w >> 16;
                            for (int i = lis; i < n2; i += sub) {
                                const double lv = lds_f64(l2_a + 16u * (unsigned)i);
                                const int lp = lds_s32(l2_a + 16u * (unsigned)i + 8u), lc = lds_s32(l2_a + 16u * (unsigned)i + 12u);
                                sts_f64(val_a + 8u * (unsigned)lp, __dmul_rn(lv, lds_f64(x_a + 8u * (unsigned)lc)));
                            }
                        }
                    }
                    __syncwarp();
                    SL_MARK(10)
                }
                // ---- the prefix chains.  Everything the post-barrier half of the FIRST round needs is kept in
                // registers (c_*): its dependent path is x of the late entries -> product -> store -> suffix chain ->
                // divide -> store.
                unsigned c_suf = 0, c_late = 0, c_val = 0;       // shared-space addresses: suffix products, late list, products
                int c_cnt = 0, c_maxc = 0, c_nlate = 0;
                double c_t = 0.0, c_d = 0.0, c_y = 0.0;
                bool c_dsafe = false, c_store = false;
                unsigned c_xaddr = 0;
                StreamLateDev c_e = {0.0, 0, 0}, c_e2 = {0.0, 0, 0};
                for (int base = 0; base < hd.x; base += S) {
                    const int ri = base + slot;
                    const bool mine = ri < hd.x;
                    unsigned char *rec = blk + (mine ? rec_off[ri] : rec_off[0]);
                    const int4 rh = *reinterpret_cast<const int4 *>(rec);               // row, prefix_pad, len_pad, nlate | nlate2 << 16
#if defined(AMGB200_ABLATE) && (AMGB200_ABLATE == 14 || AMGB200_ABLATE == 16)
                    const int cnt = 0;
#else
                    const int cnt = mine ? rh.y : 0;
#endif
                    const int maxc = __reduce_max_sync(FULL, cnt);
                    const double t = chain_fold_slots(mine ? bseg[rh.x - hd.y] : 0.0, smem_u32(rec + 32), cnt, maxc, zeros_a);
                    if (base == 0) {
                        c_t = t; c_d = reinterpret_cast<const double *>(rec)[2];
                        c_y = __ddiv_rn(1.0, c_d); c_dsafe = gs_quotient_dsafe(c_d);
                        c_store = mine && leader && fabs(c_d) > GS_TINY;
                        c_xaddr = x_a + 8u * (unsigned)rh.x;
                        asm volatile("" : "+r"(c_xaddr));            // keep the address in a register (ptxas otherwise rebuilds it from SR_CgaCtaId after the chain)
                        c_val = smem_u32(rec + 32);
                        c_suf = c_val + 8u * (unsigned)rh.y;
#if defined(AMGB200_ABLATE) && AMGB200_ABLATE == 15
                        c_cnt = 0;
#else
                        c_cnt = mine ? rh.z - rh.y : 0;
#endif
                        c_maxc = __reduce_max_sync(FULL, c_cnt);
#if defined(AMGB200_ABLATE) && AMGB200_ABLATE == 12
                        c_nlate = 0;
#else
                        c_nlate = mine ? (rh.w & 0xffff) : 0;
#endif
                        c_late = c_val + 12u * (unsigned)rh.z;
                        if (lis < c_nlate) c_e = *reinterpret_cast<const StreamLateDev *>(rec + 32 + 12 * rh.z + 16 * lis);
                        if (lis + sub < c_nlate) c_e2 = *reinterpret_cast<const StreamLateDev *>(rec + 32 + 12 * rh.z + 16 * (lis + sub));
                    } else if (mine && leader) reinterpret_cast<double *>(rec)[3] = t;
#ifdef AMGB200_TIMELINE
                    if (t == 1.2345e300) tl[11] = 1;
#endif
                }
                SL_MARK(2)
                __syncwarp();
                // hand-off from the folding warp of wavefront g-1: a named barrier between exactly these two warps (bar.arrive /
                // bar.sync, ~130 cycles faster than polling the mbarrier, which the wavefront g+1 still uses for its late2 wait)
                if (g > 0) asm volatile("bar.sync %0, %1;" ::"r"(1 + ((g - 1) & 7)), "r"(64) : "memory");
                SL_MARK(3)
                // ---- after done(g-1): late products (the first two per lane from registers), suffix chains, x_k
                {
                    {
                        const double x1 = lds_f64(x_a + 8u * (unsigned)c_e.col), x2 = lds_f64(x_a + 8u * (unsigned)c_e2.col);
                        if (lis < c_nlate) sts_f64(c_val + 8u * (unsigned)c_e.pos, __dmul_rn(c_e.val, x1));
                        if (lis + sub < c_nlate) sts_f64(c_val + 8u * (unsigned)c_e2.pos, __dmul_rn(c_e2.val, x2));
                    }
                    for (int i = lis + 2 * sub; i < c_nlate; i += sub) {
                        const double lv = lds_f64(c_late + 16u * (unsigned)i);
                        const int lp = lds_s32(c_late + 16u * (unsigned)i + 8u), lc = lds_s32(c_late + 16u * (unsigned)i + 12u);
                        sts_f64(c_val + 8u * (unsigned)lp, __dmul_rn(lv, lds_f64(x_a + 8u * (unsigned)lc)));
                    }
                    __syncwarp();
                    SL_MARK(4)
                    const double t = chain_fold_slots(c_t, c_suf, c_cnt, c_maxc, zeros_a);
#if defined(AMGB200_TIMELINE) && AMGB200_TIMELINE == 1
                    if (t == 1.2345e300) tl[11] = 1;
                    SL_MARK(5)
#endif
#if defined(AMGB200_TIMELINE)
                    tl[5] += c_maxc;
#endif
#if defined(AMGB200_ABLATE) && AMGB200_ABLATE == 11
                    if (c_store) sts_f64(c_xaddr, __dmul_rn(t, c_d));
#else
                    {
                        const double xn = gs_quotient_pre(t, c_d, c_y, c_dsafe, recip, c_store);
                        if (c_store) sts_f64(c_xaddr, xn);
                    }
#endif
                    SL_MARK(6)
                }
                for (int base = S; base < hd.x; base += S) {                             // wavefront wider than S rows: further rounds
                    const int ri = base + slot;
                    const bool mine = ri < hd.x;
                    unsigned char *rec = blk + (mine ? rec_off[ri] : rec_off[0]);
                    const int4 rh = *reinterpret_cast<const int4 *>(rec);
                    double *val = reinterpret_cast<double *>(rec + 32);
                    if (mine) {
                        const StreamLateDev *lt = reinterpret_cast<const StreamLateDev *>(reinterpret_cast<const int *>(val + rh.z) + rh.z);
                        for (int i = lis; i < (rh.w & 0xffff); i += sub) { const StreamLateDev e = lt[i]; val[e.pos] = __dmul_rn(e.val, lds_f64(x_a + 8u * (unsigned)e.col)); }
                    }
                    __syncwarp();
                    const double2 dt = *reinterpret_cast<const double2 *>(rec + 16);    // diag, prefix accumulator
                    const int cnt = mine ? rh.z - rh.y : 0;
                    const int maxc = __reduce_max_sync(FULL, cnt);
                    const double t = chain_fold_slots(dt.y, smem_u32(val + rh.y), cnt, maxc, zeros_a);
                    if (mine && leader && fabs(dt.x) > GS_TINY) sts_f64(x_a + 8u * (unsigned)rh.x, gs_quotient(t, dt.x, recip));
                }
                // the leaders' x stores are ordered before lane 0's arrive by the warp barrier (memory ordering among its
                // participants); mbarrier.arrive has release, the waiters' try_wait acquire semantics at CTA scope
                // bar.arrive orders this thread's prior shared-memory stores before the consumer's bar.sync (PTX ISA, bar:
                // producer/consumer example)
                if (g + 1 < totalw) asm volatile("bar.arrive %0, %1;" ::"r"(1 + (g & 7)), "r"(64) : "memory");
                __syncwarp();
                if (lane == 0 && g + 2 < totalw) mbar_arrive(done0 + 8u * (unsigned)(g & 7));
                SL_MARK(8)
            }
            // every warp's generic writes to the block precede its reuse by the async proxy: fence, group barrier, release
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (G > 1) asm volatile("bar.sync %0, %1;" ::"r"(9 + grp), "r"(G * 32) : "memory");   // wavefront g is complete
            else __syncwarp();
            if (folder && lane == 0) mbar_arrive(smem_u32(bars + STREAM_NS + s));
        }
#ifdef AMGB200_TIMELINE
        if (dbg && lane == 0 && folder && grp < 2) for (int i = 0; i < 12; ++i) dbg[grp * 16 + i] = tl[i];
#endif
    } else if (warp == D * G && lane == 0) {
        // ---- loader: in-order ring allocation; a block is released when its group's folding warp has arrived on empty[s]
        int head = 0, tail = 0, inflight = 0, g_old = 0, wl = 0;
        for (int g = 0; g < totalw; ++g) {
            const int s = g & (STREAM_NS - 1);
            const int o0 = blk_ptr[wl], o1 = blk_ptr[wl + 1];
            const int i0a = wf_row_ptr[wl] & ~1, bcnt = (wf_row_ptr[wl + 1] - i0a + 1) & ~1;
            const int bytes = (o1 - o0) * 16, need = (bytes + bcnt * 8 + 127) & ~127;     // blocks stay 128-byte aligned in the ring
            for (;;) {
                if (inflight == 0) { head = tail = 0; break; }
                if (inflight < STREAM_NS) {
                    if (head >= tail) {                       // (head == tail only when nothing is in flight: strict tests below)
                        if (head + need <= ring_bytes) break;
                        if (need < tail) { head = 0; break; }
                    } else if (head + need < tail) break;
                }
                // (polite wait: the loader is several wavefronts ahead, and a tight try_wait loop would steal issue slots
                // from the folding warp that shares its SM sub-partition)
                mbar_wait_sleep(smem_u32(bars + STREAM_NS + (g_old & (STREAM_NS - 1))), (g_old / STREAM_NS) & 1);
                ++g_old; --inflight;
                tail = inflight ? stage_off[g_old & (STREAM_NS - 1)] : head;
            }
            stage_off[s] = head;
            const unsigned full = smem_u32(bars + s);
            mbar_arrive_expect_tx(full, (unsigned)(bytes + bcnt * 8));
            bulk_g2s(smem_u32(ring + head), stream + (size_t)o0 * 16, (unsigned)bytes, full);
            bulk_g2s(smem_u32(ring + head + bytes), b + i0a, (unsigned)(bcnt * 8), full);
            head += need; ++inflight;
            if (++wl == W) wl = 0;
        }
    }
    __syncwarp();
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) xg[i] = x[i];
}

// ------------------------------------------------------------------------------------------
// Streaming ordered sweeps on a 16-CTA cluster (levels too wide / too large for one SM).
//
// Same division of labour as gs_stream_cta_kernel -- loader warp with a bulk-async ring, product lanes, folding warps with
// row slots, two alternating groups -- in every CTA of the cluster; row i of a wavefront belongs to CTA i % 16.  What
// crosses SMs never goes through a release fence on the dependency path:
//   * a finished x_k is PUSHED into the exchange buffers of all 16 CTAs with st.async (8 bytes, completion counted on the
//     receiving CTA's mbarrier WF[g%3] by complete_tx): data and notification travel together.  Every CTA therefore holds
//     the x of the last three wavefronts in its own shared memory, and entries at wavefront distance 1 or 2 ("late",
//     flagged by the host) are ordinary local shared-memory reads after the wait;
//   * x in GLOBAL memory only serves entries at distance >= 3: a PUBLISHER warp per CTA copies its share of each complete
//     wavefront from the exchange buffer to global memory, fences at GPU scope (the ~1 us that used to sit between two
//     wavefronts) and arrives on GV[g%4] of every CTA; the product pass of wavefront g waits for GV(g-3), long complete.
// Dynamic shared memory: [full[4] empty[4] | WF[3] | GV[4] | ring offsets | zeros : 256 B][3 x xb_cap doubles][ring]
// ------------------------------------------------------------------------------------------
constexpr int CLUSTER_CTAS = 16;
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
constexpr int XC_CTAS = CLUSTER_CTAS;
constexpr int XC_G = 4;                        // product warps per group
constexpr int XC_MAX_D = 3;                    // consumer groups (wavefronts in flight)
constexpr int XC_WARPS = XC_MAX_D * XC_G + 2;  // + loader + publisher
constexpr int XC_NS = 8;                       // ring descriptors
constexpr int XC_HDR = 384;                    // barriers, descriptors and the block of zeros in front of the exchange buffers
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(unsigned raddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// 8 bytes into another CTA's shared memory; the receiving CTA's mbarrier counts them (complete_tx) when they have landed
__device__ __forceinline__ void st_async_f64(unsigned raddr, double v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr), "l"(__double_as_longlong(v)), "r"(rbar) : "memory");
}
// x of one late entry: from the exchange buffer of the wavefront that produced it (distance d = 1 or 2 before wavefront g),
// or -- for the first wavefronts of a launch, whose predecessors belong to the previous launch -- from global memory
template <int NB>
__device__ __forceinline__ double xc_late_x(int g, int src, int col, unsigned xb_a, int xb_cap, const double *x) {
    const int d = (src & 3) + 1;
    double v;
    if (g - d < 0) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(x + col) : "memory");
    else v = lds_f64(xb_a + 8u * ((unsigned)((g - d) % NB) * (unsigned)xb_cap + ((unsigned)src >> 2)));
    return v;
}

// NB = exchange buffers per CTA = late distance + 1 (3 or 4; compile-time: it is the modulus of every buffer / barrier index)
// NB = exchange buffers per CTA = late distance + 1 (3 or 4; compile-time: it is the modulus of every buffer / barrier index)
template <int NB>
__global__ void __launch_bounds__(32 * XC_WARPS) gs_stream_cluster_kernel(
    const unsigned char *__restrict__ stream, const int *__restrict__ blk_ptr, const int *__restrict__ wf_row_ptr,
    const double *__restrict__ b, double *x, int W, int nsweeps, int F, int S, int P, int D, int ring_bytes, int xb_cap, int recip, long long *dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);      // full[0..8) empty[8..16) WF[16..20) GV[20..28)
    volatile int *stage_off = reinterpret_cast<volatile int *>(smem_raw + 224);
    const unsigned zeros_a = smem_u32(smem_raw + 256);
    double *xb = reinterpret_cast<double *>(smem_raw + XC_HDR);
    const unsigned xb_a = smem_u32(xb);
    unsigned char *ring = smem_raw + XC_HDR + (size_t)NB * xb_cap * 8;       // NB = late distance + 1 exchange buffers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int totalw = W * nsweeps;
    const int C = XC_CTAS, G = XC_G;
    const unsigned cta = cluster_ctarank();
#ifdef AMGB200_TIMELINE
    long long tl[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tl_prev = clock64();
#define XC_MARK(i) { const long long c_ = clock64(); tl[i] += c_ - tl_prev; tl_prev = c_; }
#else
#define XC_MARK(i)
#endif
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + XC_NS), wf0 = smem_u32(bars + 2 * XC_NS), gv0 = smem_u32(bars + 2 * XC_NS + 4);
    if (threadIdx.x == 0) {
        for (int s = 0; s < XC_NS; ++s) { mbar_init(full0 + 8u * s, 1); mbar_init(empty0 + 8u * s, 1); }
        for (int s = 0; s < 8; ++s) mbar_init(gv0 + 8u * s, C);
        for (int s = 0; s < 4; ++s) mbar_init(wf0 + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 8) reinterpret_cast<double *>(smem_raw + 256)[threadIdx.x] = 0.0;
    __syncthreads();
    cluster_arrive(); cluster_wait();            // every CTA's barriers exist before anybody signals them remotely
    if (warp < D * G) {
        // D groups of G warps; group d walks the wavefronts g = d, d+D, ...  The product pass of wavefront g needs global x
        // through wavefront g-3 only (GV), so up to three wavefronts can be in flight.
        const int grp = warp / G, r = warp - grp * G;
        const int f = (r - grp % G + G) % G;     // folding warps of a group: f < F (the groups' first folders on different sub-partitions)
        const bool folder = f < F;
        const int sub = 32 / S, slot = lane / sub;
        const int lis = lane - slot * sub;       // lane in slot
        const int FS = F * S;
        const int RP = 32 / P, lr = lane % P, unit = r * RP + lane / P;      // product pass: P lanes per row, G*RP rows in flight per CTA
        for (int g = grp; g < totalw; g += D) {
            const int s = g & (XC_NS - 1);
            XC_MARK(9)
            mbar_wait(full0 + 8u * s, (g / XC_NS) & 1);
            XC_MARK(0)
            unsigned char *blk = ring + stage_off[s];
            const int4 hd = *reinterpret_cast<const int4 *>(blk);                   // rows of this CTA, first row of the wavefront, its width, block bytes
            const int2 hf = *reinterpret_cast<const int2 *>(blk + 16);              // late entries of the block, byte offset of their list
            const int *rec_off = reinterpret_cast<const int *>(blk + 32);
            // right-hand side of my first-round row: requested now, used after the product pass
            double b0 = 0.0;
            if (folder && slot * F + f < hd.x) b0 = __ldg(b + hd.y + (int)cta + (slot * F + f) * C);
            // global x visible through wavefront g-NB (this also means: the previous use of WF[g % NB], by wavefront g-NB, is over)
            if (g >= NB) mbar_wait_cluster(gv0 + 8u * ((g - NB) & 7), ((g - NB) >> 3) & 1);
            // this CTA expects the whole wavefront g (width x 8 bytes) in its exchange buffer g % 3: one arrival per phase
            if (f == 0 && lane == 0) mbar_arrive_expect_tx(wf0 + 8u * (unsigned)(g % NB), (unsigned)hd.z * 8u);
            XC_MARK(1)
            // ---- before the wavefront barrier: products of every entry (late ones are redone below).  P lanes per row, all
            // row groups of the warp walk in lockstep (a diverged warp would serialise the L2 round trips of its row groups)
            for (int rbase = 0; rbase < hd.x; rbase += G * RP) {
                const int ri = rbase + unit;
                unsigned val_a = 0, col_a = 0;
                int len_pad = 0;
                if (ri < hd.x) {
                    unsigned char *rec = blk + rec_off[ri];
                    len_pad = reinterpret_cast<const int *>(rec)[2];
                    val_a = smem_u32(rec + 32); col_a = val_a + 8u * (unsigned)len_pad;
                }
                const int maxlen = __reduce_max_sync(FULL, len_pad);
                for (int p0 = 0; p0 < maxlen; p0 += 8 * P) {
                    int j[8];
                    double v[8], xv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int p = p0 + u * P + lr; j[u] = -1; if (p < len_pad) j[u] = lds_s32(col_a + 4u * (unsigned)p); }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int p = p0 + u * P + lr;
                        v[u] = 0.0; xv[u] = 0.0;
                        if (j[u] >= 0) {
                            v[u] = lds_f64(val_a + 8u * (unsigned)p);
                            asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(xv[u]) : "l"(x + j[u]) : "memory");
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int p = p0 + u * P + lr; if (j[u] >= 0) sts_f64(val_a + 8u * (unsigned)p, __dmul_rn(v[u], xv[u])); }
                }
            }
            XC_MARK(2)
            asm volatile("bar.sync %0, %1;" ::"r"(9 + grp), "r"(G * 32) : "memory");      // all products of my CTA's rows are in place
            XC_MARK(3)
            // ... and the prefix chains; the state of the first round stays in registers (c_*)
            bool c_mine = false;
            unsigned c_suf = 0;
            int c_cnt = 0, c_maxc = 0, c_row = 0, c_li = 0;
            double c_t = 0.0, c_d = 0.0, c_y = 0.0;
            bool c_dsafe = false;
            if (folder) {
                for (int base = 0; base < hd.x; base += FS) {
                    const int ri = base + slot * F + f;
                    const bool mine = ri < hd.x;
                    unsigned char *rec = blk + (mine ? rec_off[ri] : rec_off[0]);
                    const int4 rh = *reinterpret_cast<const int4 *>(rec);               // row, prefix_pad, len_pad, nlate
                    const int cnt = mine ? rh.y : 0;
                    const int maxc = __reduce_max_sync(FULL, cnt);
                    const double t0 = base == 0 ? b0 : (mine ? __ldg(b + rh.x) : 0.0);
                    const double t = chain_fold_slots(t0, smem_u32(rec + 32), cnt, maxc, zeros_a);
                    if (base == 0) {
                        c_mine = mine; c_row = rh.x; c_t = t; c_d = reinterpret_cast<const double *>(rec)[2]; c_li = ri;
                        c_y = __ddiv_rn(1.0, c_d); c_dsafe = gs_quotient_dsafe(c_d);
                        c_suf = smem_u32(rec + 32) + 8u * (unsigned)rh.y;
                        c_cnt = mine ? rh.z - rh.y : 0;
                        c_maxc = __reduce_max_sync(FULL, c_cnt);
                    } else if (mine && lis == 0) reinterpret_cast<double *>(rec)[3] = t;
                }
                __syncwarp();
            }
            XC_MARK(4)
            // the (static) descriptors of my first two late entries are fetched before the wait: after it only x -> product -> store remain
            const unsigned flat_a = smem_u32(blk + hf.y), blk_a = smem_u32(blk);
            const int i0 = r * 32 + lane, i1 = i0 + G * 32;
            double pv0 = 0.0, pv1 = 0.0;
            int pd0 = 0, pd1 = 0, ps0 = 0, ps1 = 0, pc0 = 0, pc1 = 0;
            if (i0 < hf.x) { const unsigned a = flat_a + 24u * (unsigned)i0; pv0 = lds_f64(a); pd0 = lds_s32(a + 8u); pc0 = lds_s32(a + 12u); ps0 = lds_s32(a + 16u); }
            if (i1 < hf.x) { const unsigned a = flat_a + 24u * (unsigned)i1; pv1 = lds_f64(a); pd1 = lds_s32(a + 8u); pc1 = lds_s32(a + 12u); ps1 = lds_s32(a + 16u); }
#pragma unroll
            for (int d = NB - 1; d >= 1; --d)
                if (g >= d) mbar_wait_cluster(wf0 + 8u * (unsigned)((g - d) % NB), ((g - d) / NB) & 1);
            XC_MARK(5)
            // ---- after the barrier: ALL late products of the block (x from the exchange buffers), shared evenly by the lanes of the
            // group; then the suffix chains, and x_k pushed to every CTA
            {
                if (i0 < hf.x) sts_f64(blk_a + (unsigned)pd0, __dmul_rn(pv0, xc_late_x<NB>(g, ps0, pc0, xb_a, xb_cap, x)));
                if (i1 < hf.x) sts_f64(blk_a + (unsigned)pd1, __dmul_rn(pv1, xc_late_x<NB>(g, ps1, pc1, xb_a, xb_cap, x)));
                for (int j0 = i0 + 2 * G * 32; j0 < hf.x; j0 += G * 32) {
                    const unsigned a = flat_a + 24u * (unsigned)j0;
                    const double v = lds_f64(a);
                    const int dd = lds_s32(a + 8u);
                    sts_f64(blk_a + (unsigned)dd, __dmul_rn(v, xc_late_x<NB>(g, lds_s32(a + 16u), lds_s32(a + 12u), xb_a, xb_cap, x)));
                }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(9 + grp), "r"(G * 32) : "memory");      // all late products are in place
            XC_MARK(6)
            if (folder) {
                const unsigned xb_g = xb_a + 8u * (unsigned)(g % NB) * (unsigned)xb_cap, wf_g = wf0 + 8u * (unsigned)(g % NB);
                for (int base = 0; base < hd.x; base += FS) {
                    bool mine; unsigned suf_a; int cnt, maxc, row, li; double t, dg;
                    if (base == 0) { mine = c_mine; suf_a = c_suf; cnt = c_cnt; maxc = c_maxc; row = c_row; li = c_li; t = c_t; dg = c_d; }
                    else {
                        li = base + slot * F + f;
                        mine = li < hd.x;
                        unsigned char *rec = blk + (mine ? rec_off[li] : rec_off[0]);
                        const int4 rh = *reinterpret_cast<const int4 *>(rec);
                        suf_a = smem_u32(rec + 32) + 8u * (unsigned)rh.y;
                        cnt = mine ? rh.z - rh.y : 0; maxc = __reduce_max_sync(FULL, cnt); row = rh.x;
                        const double2 dt = *reinterpret_cast<const double2 *>(rec + 16);
                        dg = dt.x; t = dt.y;
                    }
                    t = chain_fold_slots(t, suf_a, cnt, maxc, zeros_a);
#ifdef AMGB200_TIMELINE
                    if (t == 1.2345e300) tl[11] = 1;
                    XC_MARK(7)
#endif
                    double xn = 0.0;
                    if (mine && lis == 0) {
#if defined(AMGB200_ABLATE) && AMGB200_ABLATE == 21
                        if (fabs(dg) > GS_TINY) xn = __dmul_rn(t, c_y);
#else
                        if (fabs(dg) > GS_TINY) xn = base == 0 ? gs_quotient_pre(t, dg, c_y, c_dsafe, recip) : gs_quotient(t, dg, recip);
#endif
                        else asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(xn) : "l"(x + row) : "memory");     // row without a diagonal: x_k unchanged (never speculated)
                    }
                    XC_MARK(8)
                    // the 16 destinations of every finished row, spread over all lanes of the warp: item = (slot, destination)
                    const int nslots = min(S, (hd.x - base - f + F - 1) / F);            // slots of this round that hold a row
                    for (int it0 = 0; it0 < nslots * C; it0 += 32) {
                        const int it = it0 + lane, sl = min(it / C, S - 1), c2 = it % C;
                        const double xv = __shfl_sync(FULL, xn, sl * sub);
                        const unsigned idx = (unsigned)((base + sl * F + f) * C) + cta;   // index of the row within its wavefront
                        if (it < nslots * C) st_async_f64(mapa_u32(xb_g + 8u * idx, (unsigned)c2), xv, mapa_u32(wf_g, (unsigned)c2));
                    }
                }
                XC_MARK(10)
            }
            // every warp's generic writes to the block precede its reuse by the async proxy: fence, group barrier, release
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, %1;" ::"r"(9 + grp), "r"(G * 32) : "memory");
            if (f == 0 && lane == 0) mbar_arrive(empty0 + 8u * s);
        }
#ifdef AMGB200_TIMELINE
        if (dbg && lane == 0 && f == 0 && cta == 0 && grp < 2) for (int i = 0; i < 12; ++i) dbg[grp * 16 + i] = tl[i];
#endif
    } else if (warp == D * G) {
        if (lane == 0) {
            // ---- loader (see gs_stream_cta_kernel); block of (wavefront wl, this CTA)
            int head = 0, tail = 0, inflight = 0, g_old = 0, wl = 0;
            for (int g = 0; g < totalw; ++g) {
                const int s = g & (XC_NS - 1);
                const int o0 = blk_ptr[wl * C + (int)cta], o1 = blk_ptr[wl * C + (int)cta + 1];
                const int need = (o1 - o0) * 16;
                for (;;) {
                    if (inflight == 0) { head = tail = 0; break; }
                    if (inflight < XC_NS) {
                        if (head >= tail) {
                            if (head + need <= ring_bytes) break;
                            if (need < tail) { head = 0; break; }
                        } else if (head + need < tail) break;
                    }
                    mbar_wait_sleep(empty0 + 8u * (g_old & (XC_NS - 1)), (g_old / XC_NS) & 1);
                    ++g_old; --inflight;
                    tail = inflight ? stage_off[g_old & (XC_NS - 1)] : head;
                }
                stage_off[s] = head;
                mbar_arrive_expect_tx(full0 + 8u * s, (unsigned)need);
                bulk_g2s(smem_u32(ring + head), stream + (size_t)o0 * 16, (unsigned)need, full0 + 8u * s);
                head += need; ++inflight;
                if (++wl == W) wl = 0;
            }
        }
    } else if (warp == D * G + 1) {
        // ---- publisher: my share of each complete wavefront, exchange buffer -> global x, GPU-scope fence, GV(g) on every CTA
        int wl = 0;
        for (int g = 0; g < totalw; ++g) {
            const int start = wf_row_ptr[wl], width = wf_row_ptr[wl + 1] - start;
            mbar_wait_cluster(wf0 + 8u * (unsigned)(g % NB), (g / NB) & 1);
            const double *xb_g = xb + (size_t)(g % NB) * xb_cap;
            for (int i = (int)cta + C * lane; i < width; i += C * 32) x[start + i] = xb_g[i];
            __threadfence();
            __syncwarp();
            if (lane < C) mbar_arrive_remote(mapa_u32(gv0 + 8u * (unsigned)(g & 7), (unsigned)lane));
            if (++wl == W) wl = 0;
        }
    }
    __syncwarp();
    cluster_arrive(); cluster_wait();            // nobody leaves while its exchange buffers / barriers may still be addressed remotely
}

// Ordered sweeps inside ONE thread-block cluster (16 CTAs = 16 SMs on one die): the wavefronts of the
// level are walked in order, the items of a wavefront are spread over all warps of the cluster, and
// consecutive wavefronts are separated by the hardware cluster barrier (barrier.cluster, ~0.2 us)
// instead of a round trip through L2 atomics (~2-5 us for a grid-wide counter).  The barrier is split:
// after its last store a warp *arrives*, then prefetches the matrix entries of its next item, then
// *waits*.  x lives in global memory and is read at L2 (ld.cg) after the acquire.

constexpr int CLUSTER_WARPS_SELL = 8, CLUSTER_WARPS_CSR = 16;
// ONE (SELL only): every row fits in one register chunk (max row length <= 20): no next-chunk registers, which leaves room
// for a second register-resident item -- my item of the NEXT wavefront, requested before this wavefront's rows are finished
// SCH: entries per thread kept in registers (20, or 28 with ONE: whole rows of 27-point operators in one L2 round trip)
template <int KIND, bool EXACT, bool ONE = false, int SCH = 20>
__global__ void __launch_bounds__(KIND == 0 ? 32 * CLUSTER_WARPS_SELL : 32 * CLUSTER_WARPS_CSR) gs_ordered_cluster_kernel(
    DMat A, const double *__restrict__ b, double *x, const int *__restrict__ wf_item_ptr, int W, int nsweeps, long long *dbg) {
    __shared__ double sprod[KIND == 1 ? CLUSTER_WARPS_CSR * STAGE : 1];
    using Item = typename std::conditional<KIND == 0, SellItem<SCH>, CsrItem>::type;
    using Desc = typename Item::Desc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    // rows of a wavefront go round-robin over the 16 SMs first, then over the warps of an SM: a warp-wide DADD
    // occupies an SM sub-partition's fp64 pipe for 4 cycles (measured, tools/ubench2.cu), so more than two
    // chaining warps per sub-partition slow every chain down
    const int gw = warp * CLUSTER_CTAS + (int)cluster_ctarank();
    const int TW = CLUSTER_CTAS * nw;
    const int totalw = W * nsweeps;
    double *sp = sprod + (KIND == 1 ? warp * STAGE : 0);
    auto finish = [&](Item &w) {
        if constexpr (KIND == 0 && ONE) gs_finish_sell_one<true, SCH, true>(w, x);
        else if constexpr (KIND == 0) gs_finish_sell<true, SCH, true>(w, x);
        else gs_finish_csr<true, EXACT>(A, w, x, lane, sp);
    };
    constexpr bool DB = KIND == 0 && ONE && SCH <= 20;
    // Look-ahead pipeline over the (static) schedule, one stage per wavefront step, so that no dependent
    // load chain (wavefront table -> item descriptor -> matrix entries) is exposed between two barriers:
    //   (a0,a1) item range of wavefront g     cur : entries of my item in g     (loaded during step g-1)
    //   (b0,b1) item range of wavefront g+1   dn  : descriptor of my item in g+1 (loaded during step g-1)
    //   (c0,c1) item range of wavefront g+2
    int wl2 = 2 % W;
    int a0 = wf_item_ptr[0], a1 = wf_item_ptr[1];
    int b0 = wf_item_ptr[1 % W], b1 = wf_item_ptr[1 % W + 1];
    int c0 = wf_item_ptr[wl2], c1 = wf_item_ptr[wl2 + 1];
    Item cur, nxt;
    Desc dn = {};
    bool have = a0 + gw < a1, have_n = totalw > 1 && b0 + gw < b1;
    if (have) cur.prologue(A, a0 + gw, lane, b);
    if (have_n) dn = Item::load_desc(A, b0 + gw);
#ifdef AMGB200_TIMING
    long long tf = 0, tf2 = 0, ta = 0, tpre = 0, tw = 0, nit = 0;
#endif
    // one wavefront step; `c` holds my item of wavefront g, `n` receives my item of wavefront g+1 (the loop below alternates
    // two register-resident items instead of copying one into the other)
    auto step = [&](Item &c, Item &n, int g) {
#ifdef AMGB200_TIMING
        long long k0 = clock64(), k1 = k0;
#endif
        // the matrix entries of my item in wavefront g+1 are requested BEFORE this wavefront's rows are finished (their
        // descriptor arrived during step g-1): they travel during the x gathers of finish() and the release fence of the
        // barrier.  Requested after the arrive they used to sit on the critical path (measured 1 825 cycles per wavefront
        // on level 1 of 128^3: memory instructions behind barrier.cluster.arrive.release wait for its fence).
        if constexpr (DB) { if (have_n) n.load_entries(A, dn, lane, b); }
        if (have) {
            finish(c);
#ifdef AMGB200_TIMING
            ++nit; k1 = clock64();
#endif
            for (int it = a0 + gw + TW; it < a1; it += TW) { c.prologue(A, it, lane, b); finish(c); }   // wider than the cluster
        }
#ifdef AMGB200_TIMING
        long long k2 = clock64(); tf += k1 - k0; tf2 += k2 - k1;
#endif
        if (g + 1 < totalw) {
            cluster_arrive();
#ifdef AMGB200_TIMING
            long long k3 = clock64(); ta += k3 - k2;
#endif
            if constexpr (!DB) { if (have_n) n.load_entries(A, dn, lane, b); }   // descriptor arrived during the previous step
            a0 = b0; a1 = b1; have = have_n;
            b0 = c0; b1 = c1; have_n = g + 2 < totalw && b0 + gw < b1;
            if (have_n) dn = Item::load_desc(A, b0 + gw);
            if (++wl2 == W) wl2 = 0;
            c0 = wf_item_ptr[wl2]; c1 = wf_item_ptr[wl2 + 1];
#ifdef AMGB200_TIMING
            long long k4 = clock64(); tpre += k4 - k3;
#endif
            cluster_wait();
#ifdef AMGB200_TIMING
            tw += clock64() - k4;
#endif
        }
    };
    if constexpr (DB) {
        for (int g = 0; g < totalw; g += 2) {
            step(cur, nxt, g);
            if (g + 1 < totalw) step(nxt, cur, g + 1);
        }
    } else {
        for (int g = 0; g < totalw; ++g) step(cur, cur, g);        // one item: its entries are requested after the arrive
    }
#ifdef AMGB200_TIMING
    if (dbg && lane == 0 && (gw < 4 || gw == TW - 1)) {
        const int s = gw < 4 ? gw : 4;
        dbg[s * 8 + 0] = tf; dbg[s * 8 + 1] = tf2; dbg[s * 8 + 2] = ta; dbg[s * 8 + 3] = tpre; dbg[s * 8 + 4] = tw; dbg[s * 8 + 5] = nit;
    }
#endif
}

// ==========================================================================================
// Data-flow ("sync-free") ordered Gauss-Seidel for WIDE levels: the whole GPU, no wavefront barriers.
//
// The reference's sweep (amg/Solve/SSS_smooth.c:16-48) orders the updates of one call totally: (sweep, pass, row).  In
// schedule numbering, position k of sweep s needs x_j of sweep s for j < k and x_j of sweep s-1 for j > k (rows of one
// wavefront are never adjacent, wavefronts are numbered in execution order).
//
// Every x_k travels as a 16-byte record {lo32(x), version, hi32(x), version}, version = launch base + number of updates of
// row k.  A consumer gathers the records of its columns and polls the ones that are not there yet until both version words
// are equal and >= the version it needs: value and readiness arrive in ONE L2 round trip -- no flag, no fence, no barrier,
// no atomic on the dependency path (tools/ubench3.cu: 250-400 ns per hop against ~1 000 ns for "store x, fence, store flag /
// poll flag, load x").  Each 8-byte half carries its own version word, so only 8-byte single-copy atomicity is assumed (what
// the NCCL LL protocol assumes).  With a structurally symmetric matrix the version found is exactly the one needed: the
// producer's next update depends on the consumer's own result (checked at upload; other levels keep the barrier kernels).
// In the first sweep of a launch the not-yet-updated neighbours (j > k) are read from the plain x vector (they cannot change
// before row k is done); versions grow monotonically across launches, so the records never need to be reset.
// The records (16 bytes per row) stay L2-resident on all but the largest levels.  (A push variant -- one private operand slot per
// matrix entry, written by the producer through a mirror table, read coalesced and polled by its owner only -- was built and
// measured: no line is polled by several SMs, but 16 bytes written and read per ENTRY make it DRAM-bound on large levels and
// it lost everywhere: 2.33 vs 1.45 ms per sweep on level 1 of 128^3, 3.28 vs 2.63 on level 0 of the 27-point 96^3.)
//
// Persistent cooperative grid, thread per row (a slice never straddles a wavefront), slices dealt round-robin over all
// warps in schedule order.  A warp only waits for rows at earlier positions of the total order and takes its items in
// increasing order, so the earliest unfinished item is always runnable: no deadlock as long as the grid is co-resident.
// Measured facts that shaped the loop (tools/ubench4.cu, -DAMGB200_DF_TIMING): volatile loads whose results are examined one
// by one serialise into one L2 round trip each (19 operands: 19 000 cycles), so operands are requested in batches; a warp
// keeps only ~64 sectors in flight, so the first round over all operands costs several memory latencies and runs BEFORE the
// warp goes to sleep on its hint; warps that are far ahead sleep on a per-wavefront hint word (written, without any ordering
// guarantee, by the row that closes a wavefront) instead of flooding L2 with polls; correctness never depends on the hint.
// ==========================================================================================
struct __align__(16) XRec { unsigned lo, v0, hi, v1; };
constexpr int DF_BLOCK = 128;
constexpr int DF_BATCH = 10;          // operand loads in flight per thread
#ifndef DF_MINB
#define DF_MINB 1
#endif
constexpr int DF_HINT_STRIDE = 32;    // one hint word per 128-byte line: the words of consecutive wavefronts live in different L2 slices
#ifndef DF_GATE_NS
#define DF_GATE_NS 400
#endif

__device__ __forceinline__ void st_xrec(XRec *p, double x, unsigned ver) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(ver), "r"((unsigned)(b >> 32)), "r"(ver) : "memory");
}
__device__ __forceinline__ bool ld_xrec(const XRec *p, unsigned need, double &x) {
    unsigned lo, v0, hi, v1;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(v0), "=r"(hi), "=r"(v1) : "l"(p) : "memory");
    x = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    return v0 == v1 && (int)(v0 - need) >= 0;
}

// the same load split in two, so that a batch of loads can be ISSUED back to back before the first result is examined
// (volatile asm statements stay in program order: a load whose result is checked right away serialises the batch into one
// L2 round trip per entry -- measured 19 000 cycles for the 19 operands of a row)
__device__ __forceinline__ void ld_xrec_raw(const XRec *p, uint4 &r) {
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
}
__device__ __forceinline__ bool xrec_ok(const uint4 &r, unsigned need, double &x) {
    x = __longlong_as_double((long long)(((unsigned long long)r.z << 32) | r.x));
    return r.y == r.w && (int)(r.y - need) >= 0;
}

// upload: slice of every row, then the number of entries (k, j) without a stored mirror entry (j, k)
__global__ void __launch_bounds__(BLOCK) df_row_slice_kernel(int nslices, const int *__restrict__ slice_row, int *row_slice) {
    const int s = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int k = slice_row[s] + lane;
    if (k < slice_row[s + 1]) row_slice[k] = s;
}
__global__ void __launch_bounds__(BLOCK) df_symmetry_kernel(DMat A, const int *__restrict__ row_slice, int *missing) {
    const int s = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= A.nitems) return;
    const int r0 = A.slice_row[s], r1 = A.slice_row[s + 1];
    const long long p0 = A.slice_ptr[s];
    const int width = (int)((A.slice_ptr[s + 1] - p0) >> 5);
    const int k = r0 + lane;
    int bad = 0;
    for (int u = 0; u < width; ++u) {
        const int j = A.col[p0 + (long long)u * 32 + lane];
        if (k < r1 && j >= 0 && j != k) {
            const int s2 = row_slice[j];
            const long long q0 = A.slice_ptr[s2] + (j - A.slice_row[s2]);
            const int w2 = (int)((A.slice_ptr[s2 + 1] - A.slice_ptr[s2]) >> 5);
            bool found = false;
            for (int e = 0; e < w2 && !found; ++e) found = A.col[q0 + (long long)e * 32] == k;
            bad += !found;
        }
    }
    if (bad) atomicAdd(missing, bad);
}

__device__ __noinline__ void tacc_sink(long long *p) { if (p) p[127] = 1; }
template <int SCH>
__global__ void __launch_bounds__(DF_BLOCK, DF_MINB) gs_dataflow_kernel(DMat A, const double *__restrict__ b, double *x, XRec *rec,
                                                               const int *__restrict__ item_wf, const int *__restrict__ wf_item_ptr, unsigned *hint,
                                                               int W, int nsweeps, int ahead, unsigned vbase, long long *dbg) {
#ifdef AMGB200_DF_TIMING
    long long tg = 0, tf = 0, tp_ = 0, tt = 0, nr = 0, ni = 0, ts = 0;
#define DF_CLK(v) const long long v = clock64();
#else
#define DF_CLK(v)
#endif
    const int lane = threadIdx.x & 31;
    const int TW = gridDim.x * (DF_BLOCK / 32);
    const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;            // consecutive items on different SMs
    const int nitems = A.nitems;
    const long long total = (long long)nitems * nsweeps;
    int s = gw / nitems, q = gw - s * nitems;
    for (long long t = gw; t < total; t += TW) {
        // ---- static data of the item: slice descriptor, first chunk of matrix entries, right-hand side
        const int r0 = A.slice_row[q], r1 = A.slice_row[q + 1];
        const long long p0 = A.slice_ptr[q];
        const int width = (int)((A.slice_ptr[q + 1] - p0) >> 5);
        const int k = r0 + lane;
        const bool active = k < r1;
        const int *cp = A.col + p0 + lane;
        const double *vp = A.val + p0 + lane;
        int j[SCH];
        double a[SCH];
#pragma unroll
        for (int u = 0; u < SCH; ++u) {
            if (u < width) { j[u] = cp[(size_t)u * 32]; a[u] = vp[(size_t)u * 32]; }
            else { j[u] = -1; a[u] = 0.0; }
        }
        const double bk = active ? b[k] : 0.0;
        const int wf = item_wf[q];
        const int gwf = s * W + wf;
        const bool closes = q + 1 == wf_item_ptr[wf + 1];                  // last item of its wavefront: writes the hint
        const unsigned vold = vbase + (unsigned)s, vnew = vold + 1u;
        double tacc = bk, d = 0.0, yrec = 0.0;
        bool dfound = false, dsafe = false;
        for (int e0 = 0; e0 < width; e0 += SCH) {
            DF_CLK(c0)
            if (e0 > 0) {
#pragma unroll
                for (int u = 0; u < SCH; ++u) {
                    if (e0 + u < width) { j[u] = cp[(size_t)(e0 + u) * 32]; a[u] = vp[(size_t)(e0 + u) * 32]; }
                    else { j[u] = -1; a[u] = 0.0; }
                }
            }
            // ---- first round: the operands of the chunk are requested in batches of DF_BATCH loads in flight (volatile loads
            // whose results are examined one by one serialise into one L2 round trip per entry); the products of the ready
            // ones replace a[u]
            unsigned pend = 0;
#pragma unroll
            for (int u0 = 0; u0 < SCH; u0 += DF_BATCH) {
                uint4 r[DF_BATCH];
                double xo[DF_BATCH];
#pragma unroll
                for (int v = 0; v < DF_BATCH; ++v) {
                    const int u = u0 + v;
                    if (u < SCH && j[u] >= 0 && j[u] != k) {
                        if (s == 0 && j[u] > k) xo[v] = __ldcg(x + j[u]);                      // not updated yet in this launch
                        else ld_xrec_raw(rec + j[u], r[v]);
                    }
                }
#pragma unroll
                for (int v = 0; v < DF_BATCH; ++v) {
                    const int u = u0 + v;
                    if (u < SCH) {
                        if (j[u] == k) { d = a[u]; a[u] = 0.0; }                               // (+0.0: neutral in the chain below)
                        else if (j[u] >= 0) {
                            double xv;
                            if (s == 0 && j[u] > k) a[u] = __dmul_rn(a[u], xo[v]);
                            else if (xrec_ok(r[v], j[u] < k ? vnew : vold, xv)) a[u] = __dmul_rn(a[u], xv);
                            else pend |= 1u << u;
                        }
                    }
                }
            }
            // the reciprocal of the diagonal is an IEEE division done while the row still waits (gs_quotient_pre)
            if (!dfound && d != 0.0) { dfound = true; yrec = __ddiv_rn(1.0, d); dsafe = gs_quotient_dsafe(d); }
            // ---- hint: sleep until the wavefront `ahead` steps back has been closed.  The first round above ran BEFORE this wait:
            // it costs several memory latencies (a warp keeps only ~64 sectors in flight, tools/ubench4.cu); what is left for the
            // dependency path is one L2 round trip per polling round
            DF_CLK(c1)
            if (e0 == 0 && gwf >= ahead && __any_sync(FULL, pend != 0)) {
                if (lane == 0) {
                    unsigned f;
                    for (;;) {
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(hint + (size_t)(gwf - ahead) * DF_HINT_STRIDE) : "memory");
                        if ((int)(f - vbase) > 0) break;
                        __nanosleep(DF_GATE_NS);
                    }
                }
                __syncwarp();
            }
            DF_CLK(c2)
            // ---- polling rounds: all operands still missing are requested together, batch by batch
            while (pend) {
#ifdef AMGB200_DF_TIMING
                ++nr;
#endif
#pragma unroll
                for (int u0 = 0; u0 < SCH; u0 += DF_BATCH) {
                    if ((pend >> u0) & ((1u << DF_BATCH) - 1u)) {
                        uint4 r[DF_BATCH];
#pragma unroll
                        for (int v = 0; v < DF_BATCH; ++v)
                            if (u0 + v < SCH && (pend & (1u << (u0 + v)))) ld_xrec_raw(rec + j[u0 + v], r[v]);
#pragma unroll
                        for (int v = 0; v < DF_BATCH; ++v) {
                            const int u = u0 + v;
                            if (u < SCH && (pend & (1u << u))) {
                                double xv;
                                if (xrec_ok(r[v], j[u] < k ? vnew : vold, xv)) { a[u] = __dmul_rn(a[u], xv); pend &= ~(1u << u); }
                            }
                        }
                    }
                }
            }
            DF_CLK(c3)
            // ---- in-order chain over the separately rounded products: branch-free (padding and the diagonal slot hold +0.0, and
            // t - (+0.0) = t bit for bit); a per-entry `if` costs a divergence region of ~125 cycles per entry on the dependency path
            // (tools/ubench5.cu)
#pragma unroll
            for (int u = 0; u < SCH; ++u) tacc = __dsub_rn(tacc, a[u]);
#ifdef AMGB200_DF_TIMING
            tf += c1 - c0; tg += c2 - c1; tp_ += c3 - c2;
#endif
        }
        DF_CLK(c4)
        if (active) {
            double xn;
            if (fabs(d) > GS_TINY) xn = gs_quotient_pre(tacc, d, yrec, dsafe, A.recip);
            else xn = __ldcg(x + k);                                       // row left as it is (SSS_smooth.c:32): only its version moves on
            st_xrec(rec + k, xn, vnew);
            if (s == nsweeps - 1 || fabs(d) <= GS_TINY) __stcg(x + k, xn);
        }
        if (closes) {
            __syncwarp();
            if (lane == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(hint + (size_t)gwf * DF_HINT_STRIDE), "r"(vbase + 1u) : "memory");
        }
#ifdef AMGB200_DF_TIMING
        { const long long c5 = clock64(); tt += c5 - c4; ++ni; }
#endif
        q += TW;
        while (q >= nitems) { q -= nitems; ++s; }
    }
#ifdef AMGB200_DF_TIMING
    if (dbg && lane == 0 && gw < 8) { long long *o = dbg + gw * 8; o[0] = ts; o[1] = tg; o[2] = tf; o[3] = tp_; o[4] = tt; o[5] = nr; o[6] = ni; }
#endif
#undef DF_CLK
}

// ---- the same data-flow scheme for levels with longer rows: a WARP per row (CSR layout in schedule numbering) ------------------
// The 32 lanes request the records of the row's columns (8 loads in flight per lane), the separately rounded products go to a
// per-warp staging area in shared memory in storage order, the in-order chain (chain_fold, 8.4 cycles per term) runs over them:
// the part in front of the first operand that is still missing BEFORE the warp sleeps on its hint, the rest after the polling
// rounds.  A product that is still missing is marked in the staging area by a reserved NaN bit pattern.
constexpr int DFW_BLOCK = 256;
constexpr unsigned long long DFW_MISSING = 0x7ff8dead0badc0deULL;
__device__ __forceinline__ bool dfw_missing(double v) { return (unsigned long long)__double_as_longlong(v) == DFW_MISSING; }
// upload: number of entries (k, j) of a CSR matrix without a stored mirror entry (j, k)
__global__ void __launch_bounds__(BLOCK) dfw_symmetry_kernel(DMat A, int *missing) {
    const int k = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= A.nrows) return;
    int bad = 0;
    for (int p = A.rptr[k] + lane; p < A.rptr[k + 1]; p += 32) {
        const int j = A.col[p];
        if (j == k) continue;
        bool found = false;
        for (int q = A.rptr[j]; q < A.rptr[j + 1] && !found; ++q) found = A.col[q] == k;
        bad += !found;
    }
    bad = __reduce_add_sync(FULL, bad);
    if (lane == 0 && bad) atomicAdd(missing, bad);
}

// dynamic shared memory: (blockDim.x / 32) staging areas of `cap` + 16 doubles (cap = longest row rounded up to 8)
__global__ void __launch_bounds__(DFW_BLOCK, 2) gs_dataflow_csr_kernel(DMat A, const double *__restrict__ b, double *x, XRec *rec,
                                                                    const int *__restrict__ wf_row_ptr, const int *__restrict__ row_wf,
                                                                    unsigned *hint, int W, int nsweeps, int ahead, unsigned vbase, int cap) {
    extern __shared__ __align__(16) double dfw_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *sp = dfw_smem + (size_t)wib * (cap + 16);
    const unsigned sp_a = (unsigned)__cvta_generic_to_shared(sp);
    const int TW = gridDim.x * (DFW_BLOCK / 32);
    const int gw = wib * gridDim.x + blockIdx.x;                          // consecutive rows on different SMs
    const int n = A.nrows;
    const long long total = (long long)n * nsweeps;
    int s = gw / n, k = gw - s * n;
    for (long long t = gw; t < total; t += TW) {
        const int p0 = A.rptr[k], len = A.rptr[k + 1] - p0, len8 = (len + 7) & ~7;
        const int wf = row_wf[k];
        const int gwf = s * W + wf;
        const bool closes = k + 1 == wf_row_ptr[wf + 1];
        const unsigned vold = vbase + (unsigned)s, vnew = vold + 1u;
        const double bk = b[k];
        // ---- first round: records of all columns, 8 loads in flight per lane; products (or the missing mark) to the staging area
        double dl = 0.0;
        int first_missing = len8;
        for (int e0 = 0; e0 < len8 + 16; e0 += 256) {
            int j[8];
            double a[8], xo[8];
            uint4 r[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * 32 + lane;
                j[u] = -1; a[u] = 0.0;
                if (e < len) { j[u] = A.col[p0 + e]; a[u] = A.val[p0 + e]; }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (j[u] >= 0 && j[u] != k) {
                    if (s == 0 && j[u] > k) xo[u] = __ldcg(x + j[u]);          // not updated yet in this launch
                    else ld_xrec_raw(rec + j[u], r[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * 32 + lane;
                if (e < len8 + 16) {
                    double prod = 0.0;                                          // padding and the diagonal: exact no-ops of the chain
                    if (j[u] == k) dl = a[u];
                    else if (j[u] >= 0) {
                        double xv;
                        if (s == 0 && j[u] > k) prod = __dmul_rn(a[u], xo[u]);
                        else if (xrec_ok(r[u], j[u] < k ? vnew : vold, xv)) prod = __dmul_rn(a[u], xv);
                        else { prod = __longlong_as_double((long long)DFW_MISSING); first_missing = min(first_missing, e); }
                    }
                    sp[e] = prod;
                }
            }
        }
        // diagonal (exactly one lane saw it), its reciprocal off the dependency path
        const unsigned dm = __ballot_sync(FULL, dl != 0.0);
        const double d = dm ? __shfl_sync(FULL, dl, __ffs(dm) - 1) : 0.0;
        const double yrec = __ddiv_rn(1.0, d);
        const bool dsafe = gs_quotient_dsafe(d);
        first_missing = __reduce_min_sync(FULL, first_missing);
        __syncwarp();
        // ---- the chain in front of the first missing operand runs before the wait
        const int pre = first_missing & ~7;
        double tacc = chain_fold<true>(bk, reinterpret_cast<const double2 *>(sp), pre);
        if (first_missing < len8) {
            if (gwf >= ahead) {
                if (lane == 0) {
                    unsigned f;
                    for (;;) {
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(hint + (size_t)(gwf - ahead) * DF_HINT_STRIDE) : "memory");
                        if ((int)(f - vbase) > 0) break;
                        __nanosleep(DF_GATE_NS);
                    }
                }
                __syncwarp();
            }
            // ---- polling rounds: every lane re-requests the records of its missing entries (matrix entries come from L1 / L2 again)
            bool more = true;
            while (more) {
                bool mine = false;
                for (int e0 = pre; e0 < len; e0 += 256) {
                    int j[8];
                    uint4 r[8];
                    bool miss[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int e = e0 + u * 32 + lane;
                        miss[u] = e < len && dfw_missing(sp[e]);
                        if (miss[u]) { j[u] = A.col[p0 + e]; ld_xrec_raw(rec + j[u], r[u]); }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (miss[u]) {
                            const int e = e0 + u * 32 + lane;
                            double xv;
                            if (xrec_ok(r[u], j[u] < k ? vnew : vold, xv)) sp[e] = __dmul_rn(A.val[p0 + e], xv);
                            else mine = true;
                        }
                    }
                }
                more = __any_sync(FULL, mine);
            }
            __syncwarp();
            tacc = chain_fold<true>(tacc, reinterpret_cast<const double2 *>(sp + pre), len8 - pre);
        }
        if (lane == 0) {
            double xn;
            if (fabs(d) > GS_TINY) xn = gs_quotient_pre(tacc, d, yrec, dsafe, A.recip);
            else xn = __ldcg(x + k);                                       // row left as it is (SSS_smooth.c:32): only its version moves on
            st_xrec(rec + k, xn, vnew);
            if (s == nsweeps - 1 || fabs(d) <= GS_TINY) __stcg(x + k, xn);
            if (closes) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(hint + (size_t)gwf * DF_HINT_STRIDE), "r"(vbase + 1u) : "memory");
        }
        __syncwarp();
        k += TW;
        while (k >= n) { k -= n; ++s; }
    }
    (void)sp_a;
}

// ------------------------------------------------------------------------------------------
// SpMV family (amg/SSS_utils.c:161-201)
// ------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ double spmv_store(double t, double alpha, const double *__restrict__ b, double *y, int k) {
    double out;
    if (MODE == MODE_MXY) out = t;
    else if (MODE == MODE_AMXPY) out = __dadd_rn(y[k], __dmul_rn(t, alpha));
    else out = __dadd_rn(b[k], __dmul_rn(t, -1.0));
    y[k] = out;
    return out;
}

// (the thread-per-row prolongation-add with rows of <= 8 entries is gather-latency bound: 32 registers give 64 warps per SM, measured
// 251 -> 238 us on level 0 of 256^3; the same bound costs the restriction 187 -> 213 us, so only this instance gets it)
template <int KIND, int MODE, int RED, bool EXACT, bool ONE = false>
__global__ void __launch_bounds__(BLOCK, (KIND == 0 && ONE && MODE == MODE_AMXPY) ? 8 : 1) spmv_kernel(DMat A, const double *__restrict__ x, double *y, const double *__restrict__ b,
                                                     double alpha, double *partial, int item0, int item1) {
    __shared__ double red[32];
    __shared__ double sprod[KIND == 1 ? WARPS_PER_BLOCK * STAGE : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int it = item0 + blockIdx.x * WARPS_PER_BLOCK + warp;
    double contrib = 0.0;
    if (it < item1) {
        if constexpr (KIND == 0) {
            SellItem<8> w;
            w.prologue(A, it, lane);
            const double t = ONE ? spmv_finish_sell_one(w, x) : spmv_finish_sell(w, x);
            if (w.k < w.r1) {
                const double out = spmv_store<MODE>(t, alpha, b, y, w.k);
                if (RED == RED_SUMSQ) contrib = __dmul_rn(out, out);
            }
        } else {
            CsrItem w;
            w.prologue(A, it, lane, nullptr);
            double d;
            const double t = EXACT ? csr_row_exact<false, false>(A, w, x, 0.0, d, lane, sprod + warp * STAGE)
                                   : csr_row_fast<false, false>(A, w, x, 0.0, d, lane);
            if (lane == 0) {
                const double out = spmv_store<MODE>(t, alpha, b, y, it);
                if (RED == RED_SUMSQ) contrib = __dmul_rn(out, out);
            }
        }
    }
    if (RED != RED_NONE) {
        const double s = block_sum(contrib, red);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------
// Residual (+) restriction in ONE launch (amg/Solve/SSS_cycle.cu:916-921: wp = b - A x ; b_{l+1} = R wp ; :929 x_{l+1} = 0).
// Exactness fixes the arithmetic: every coarse entry sums its w_u * r_{i_u} in R's storage order in one thread, so the fine
// residuals must exist as numbers before a coarse row starts; what fusion can remove is the second trip of r through HBM and
// the launch boundary.  The work of both operators is therefore dealt to ONE persistent grid as a single ticket list in which
// locality, not the schedule numbering, decides the order: the slices of A are binned into `nch` chunks by the NATURAL index of
// their first row (the schedule numbering is wavefront-major, natural indices follow the grid), the slices of R are keyed by the
// last chunk whose residual rows they read, and the list is  A(chunk 0), A(chunk 1), R(last chunk 0), A(chunk 2), R(last chunk 1) ...
// cut into blocks of RR_TICKETS tickets that are all slices of A of ONE chunk or all slices of R (padded with no-ops).  A CTA takes
// blocks round-robin in increasing order: after a block of A one thread publishes it (__syncthreads, fence, one atomic on the
// chunk's completion counter -- the other warps are already in the next block); before a block of R one warp polls the counters of
// the chunks the block reads (complete in the steady state because of the lag, which covers the blocks in flight), then the rows
// gather r while its lines are still in L2 (a chunk is 1/32 of the vector: 4 MB at 256^3).  The chunk range of a block covers
// every chunk that writes into a 128-byte line the block touches, so the lines are final and the gathers may use L1.  Every block only waits for blocks with smaller numbers and
// the grid is co-resident (cooperative launch), so the earliest unfinished block is always runnable.  Counters are never reset:
// launch `epoch` completes chunk c at epoch * items(c) (mod 2^32).
// ------------------------------------------------------------------------------------------
constexpr int RR_PER_WARP = 4;
constexpr int RR_TICKETS = WARPS_PER_BLOCK * RR_PER_WARP;
constexpr int RR_NOP = 0x7fffffff;
#ifndef AMGB200_RR_COH
#define AMGB200_RR_COH 0
#endif
constexpr bool RR_COH = AMGB200_RR_COH != 0;   // 1: gather r at L2 (developer build); 0: through L1 -- every line a slice of R touches is final before it starts (analysis.cpp)
struct RRPlan {
    const int *work;               // tickets: >= 0 slice of A (residual), < 0: ~slice of R (restriction), RR_NOP: padding
    int nblocks;                   // blocks of RR_TICKETS tickets
    const int *block_info;         // per block: slices of A -> chunk | count << 16 ; slices of R -> 0x80000000 | first chunk | last chunk << 16 (bits 16..30)
    const unsigned *chunk_items;   // slices of A per chunk
    unsigned *cnt;                 // [nch]: completed slices per chunk
    int nch;
    unsigned epoch;
};
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <bool ONE_A, bool ONE_R>
__global__ void __launch_bounds__(BLOCK) resid_restrict_kernel(DMat A, DMat R, RRPlan pl, const double *__restrict__ x, const double *__restrict__ b,
                                                               double *r, double *bc, double *xc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int blk = blockIdx.x; blk < pl.nblocks; blk += gridDim.x) {
        const unsigned info = (unsigned)pl.block_info[blk];
        const int *wk = pl.work + (size_t)blk * RR_TICKETS;
        if (!(info >> 31)) {
#pragma unroll 1
            for (int u = 0; u < RR_PER_WARP; ++u) {
                const int w = wk[u * WARPS_PER_BLOCK + warp];
                if (w == RR_NOP) continue;
                SellItem<8> it;
                it.prologue(A, w, lane);
                const double s = ONE_A ? spmv_finish_sell_one(it, x) : spmv_finish_sell(it, x);
                if (it.k < it.r1) __stcg(r + it.k, __dadd_rn(b[it.k], __dmul_rn(s, -1.0)));
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(pl.cnt + (info & 0xffffu), info >> 16);
            }
        } else {
            const unsigned lo = info & 0xffffu, hi = (info >> 16) & 0x7fffu;
            if (warp == 0) {
                for (unsigned c0 = lo; c0 <= hi; c0 += 32) {
                    const unsigned c = c0 + lane;
                    for (;;) {
                        const bool ok = c > hi || (int)(ld_acquire_u32(pl.cnt + c) - pl.epoch * pl.chunk_items[c]) >= 0;
                        if (__all_sync(FULL, ok)) break;
                        __nanosleep(256);
                    }
                }
            }
            __syncthreads();
#pragma unroll 1
            for (int u = 0; u < RR_PER_WARP; ++u) {
                const int w = wk[u * WARPS_PER_BLOCK + warp];
                if (w == RR_NOP) continue;
                SellItem<8> it;
                it.prologue(R, ~w, lane);
                const double acc = ONE_R ? spmv_finish_sell_one<8, RR_COH>(it, r) : spmv_finish_sell<8, RR_COH>(it, r);
                if (it.k < it.r1) { bc[it.k] = acc; if (xc) xc[it.k] = 0.0; }
            }
        }
    }
}

// out[q] = sum_{i<nparts} partial[q*stride + i] for q < nsum ; out[q] = max(...) for nsum <= q < nsum+nmax
__global__ void __launch_bounds__(BLOCK) reduce_partials_kernel(const double *__restrict__ partial, int nparts, int stride,
                                                                int nsum, int nmax, double *out) {
    __shared__ double red[32];
    for (int q = 0; q < nsum + nmax; ++q) {
        const double *p = partial + (size_t)q * stride;
        double v = 0.0;
        if (q < nsum) {
            for (int i = threadIdx.x; i < nparts; i += blockDim.x) v = __dadd_rn(v, p[i]);
            v = block_sum(v, red);
        } else {
            for (int i = threadIdx.x; i < nparts; i += blockDim.x) v = fmax(v, p[i]);
            v = block_max(v, red);
        }
        if (threadIdx.x == 0) out[q] = v;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// BLAS-1 pieces of the coarsest-level Krylov solvers (amg/SSS_utils.c:138-260)
// ------------------------------------------------------------------------------------------
// out[0] = sum_i x_i*y_i accumulated strictly left to right from 0.0 (SSS_blas_array_dot, :206-214),
// bit-identical to the CPU loop.  One block: all threads stage the separately rounded products of a
// tile in shared memory, thread 0 folds them in index order.  seq == 0 -> tree sum instead (FAST).
constexpr int DOT_TILE = 4096;
__global__ void __launch_bounds__(1024) dot_seq_kernel(int n, const double *__restrict__ x, const double *__restrict__ y, double *out, int seq) {
    __shared__ double prod[DOT_TILE];
    __shared__ double red[32];
    double acc = 0.0;
    if (!seq) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc = __dadd_rn(acc, __dmul_rn(x[i], y[i]));
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) out[0] = acc;
        return;
    }
    for (int base = 0; base < n; base += DOT_TILE) {
        const int m = min(DOT_TILE, n - base);
        for (int i = threadIdx.x; i < m; i += blockDim.x) prod[i] = __dmul_rn(x[base + i], y[base + i]);
        __syncthreads();
        if (threadIdx.x == 0) {
            int i = 0;
            for (; i + 8 <= m; i += 8) {
                double p0 = prod[i], p1 = prod[i + 1], p2 = prod[i + 2], p3 = prod[i + 3];
                double p4 = prod[i + 4], p5 = prod[i + 5], p6 = prod[i + 6], p7 = prod[i + 7];
                acc = __dadd_rn(acc, p0); acc = __dadd_rn(acc, p1); acc = __dadd_rn(acc, p2); acc = __dadd_rn(acc, p3);
                acc = __dadd_rn(acc, p4); acc = __dadd_rn(acc, p5); acc = __dadd_rn(acc, p6); acc = __dadd_rn(acc, p7);
            }
            for (; i < m; ++i) acc = __dadd_rn(acc, prod[i]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = acc;
}

// CG update (amg/Solve/SSS_cycle.cu:196-250): u += alpha p ; r += (-alpha) t ; tree partials of
// u.u, p.p and max|u| (they only steer the safeguards) -> partial[q*stride + block], q = 0..2
// alpha = temp1 / (t, p) is formed ON THE DEVICE from the dot product the previous kernel left in *tp (the same IEEE division the
// host would do), so that one iteration needs a single read-back; |(t, p)| <= 1e-40 (SSS_cycle.cu:190: leave the loop) -> no update
__global__ void __launch_bounds__(BLOCK) cg_update_kernel(int n, double temp1, const double *__restrict__ tp, const double *__restrict__ p,
                                                          const double *__restrict__ t, double *u, double *r, double *partial, int stride) {
    __shared__ double red[32];
    double uu = 0.0, pp = 0.0, um = 0.0;
    const double temp2 = *tp;
    if (!(fabs(temp2) > 1e-40)) return;
    const double alpha = __ddiv_rn(temp1, temp2);
    const double nalpha = -alpha;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double pi = p[i];
        const double ui = __dadd_rn(u[i], __dmul_rn(alpha, pi));
        r[i] = __dadd_rn(r[i], __dmul_rn(nalpha, t[i]));
        u[i] = ui;
        uu = __dadd_rn(uu, __dmul_rn(ui, ui));
        pp = __dadd_rn(pp, __dmul_rn(pi, pi));
        um = fmax(um, fabs(ui));
    }
    uu = block_sum(uu, red); if (threadIdx.x == 0) partial[0 * stride + blockIdx.x] = uu;
    pp = block_sum(pp, red); if (threadIdx.x == 0) partial[1 * stride + blockIdx.x] = pp;
    um = block_max(um, red); if (threadIdx.x == 0) partial[2 * stride + blockIdx.x] = um;
}

// ------------------------------------------------------------------------------------------
// The whole coarsest-level CG in ONE cooperative launch (amg/Solve/SSS_cycle.cu:15-437, stop_type = STOP_REL_RES, no preconditioner):
// the host-driven version above costs 6 launches and a read-back per iteration (~70 us on a 2 120-row matrix, 230 iterations per
// 128^3 solve; on the 39-row coarsest matrix of the 2D problem the launches ARE the solve).  Here the grid keeps the vectors in L2,
// grid-wide barriers separate the phases, and the scalar recurrences are recomputed redundantly by every thread from four numbers in
// global memory, so the control flow stays uniform without a host.  The arithmetic is the host-driven path's, operation for
// operation: rows of A p summed in storage order (warp per row / thread per row), (t, p) and (r, r) folded left to right by one
// thread, alpha = (z, r) / (t, p) and beta by IEEE division, u += alpha p, r += (-alpha) t, p = 1.0 r + beta p without FMA.
// The safeguard branches that are (almost) never taken -- stagnation restarts, the second chance after a false convergence,
// a vanishing iterate -- are NOT in the kernel: it stops with status CG_FALLBACK and the host redoes the solve from the saved
// initial guess with the host-driven code, which is bit-identical by construction.
// ------------------------------------------------------------------------------------------
constexpr int CG_FALLBACK = -1000;
struct CgArgs {
    DMat A;
    const double *b;
    double *u, *p, *r, *t, *u_best;
    double tol;
    int maxit, beta_fix;
    double *scal;          // [0] (t, p) | (r, r)  [1] u.u  [2] p.p  [3] max|u|   (written by block 0 between two grid barriers)
    double *partial;       // 3 x gridDim.x block partials
    int *status;           // [0] iterations or a negative SSS error code or CG_FALLBACK
    unsigned *barrier;     // arrival counter of the grid barrier (never reset) and the number of barriers earlier launches went through
    unsigned barrier_gen;
};
// out = sum_i x_i y_i from 0.0, left to right (SSS_blas_array_dot): block-wide staging of the separately rounded products, one thread
// folds.  Called by ONE block; result valid in thread 0.
__device__ __forceinline__ double cg_dot_seq(int n, const double *x, const double *y, double *prod) {
    double acc = 0.0;
    for (int base = 0; base < n; base += DOT_TILE) {
        const int m = min(DOT_TILE, n - base), m8 = (m + 7) & ~7;
        for (int i = threadIdx.x; i < m8 + 16; i += blockDim.x) prod[i] = i < m ? __dmul_rn(__ldcg(x + base + i), __ldcg(y + base + i)) : 0.0;   // (+0.0: exact no-ops)
        __syncthreads();
        if (threadIdx.x < 32) acc = chain_fold<false>(acc, reinterpret_cast<const double2 *>(prod), m8);   // (all lanes of warp 0: broadcast loads, 8.4 cycles per term)
        __syncthreads();
    }
    return acc;
}
// grid-wide barrier of the cooperative launch: one arrival counter that only grows (target = generation x CTAs), ~1 us against the
// 4-5 us measured for cooperative_groups' grid.sync(); a single-CTA launch (coarsest matrices of a few dozen rows) needs no more than
// __syncthreads().  Vectors are read and written at L2 (__ldcg / __stcg), so the fence + acquire pair orders them.
struct GridBarrier {
    unsigned *counter;
    unsigned gen;
    __device__ __forceinline__ void sync() {
        __syncthreads();
        if (gridDim.x > 1) {
            if (threadIdx.x == 0) {
                ++gen;
                __threadfence();
                atomicAdd(counter, 1u);
                const unsigned target = gen * gridDim.x;
                while ((int)(ld_acquire_u32(counter) - target) < 0) { }
                __threadfence();
            }
            __syncthreads();
        }
    }
};
// y = A x (RESID = false) or y = b - A x, every row exactly like spmv_kernel; x and y live in L2 (other SMs wrote x in this launch)
template <int KIND, bool RESID>
__device__ __forceinline__ void cg_spmv(const DMat &A, const double *x, const double *b, double *y, double *sprod) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * WARPS_PER_BLOCK + warp, nw = gridDim.x * WARPS_PER_BLOCK;
    for (int it = gw; it < A.nitems; it += nw) {
        if constexpr (KIND == 0) {
            SellItem<8> w;
            w.prologue(A, it, lane);
            const double t = spmv_finish_sell<8, true, true>(w, x);
            if (w.k < w.r1) __stcg(y + w.k, RESID ? __dadd_rn(__ldcg(b + w.k), __dmul_rn(t, -1.0)) : t);
        } else {
            CsrItem w;
            w.prologue(A, it, lane, nullptr);
            double d;
            const double t = csr_row_exact<true, false>(A, w, x, 0.0, d, lane, sprod + warp * STAGE);
            if (lane == 0) __stcg(y + it, RESID ? __dadd_rn(__ldcg(b + it), __dmul_rn(t, -1.0)) : t);
        }
    }
}
template <int KIND>
__global__ void __launch_bounds__(BLOCK) coarse_cg_kernel(CgArgs a) {
    GridBarrier grid{a.barrier, a.barrier_gen};
    __shared__ __align__(16) double prod[DOT_TILE + 16];
    __shared__ double sprod[KIND == 1 ? WARPS_PER_BLOCK * STAGE : 1];
    __shared__ double red[32];
    const int m = a.A.nrows;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const double maxdiff = a.tol * 1e-4;
    int iter = 0, iter_best = 0;
    double absres = 1e+20, absres_best = 1e+20, relres, normr0, temp1, temp2, alpha, beta;

    // r = b - A u ; (r, r)
    cg_spmv<KIND, true>(a.A, a.u, a.b, a.r, sprod);
    grid.sync();
    if (blockIdx.x == 0) { const double rr = cg_dot_seq(m, a.r, a.r, prod); if (threadIdx.x == 0) __stcg(a.scal, rr); }
    grid.sync();
    {
        const double rr = __ldcg(a.scal);
        const double absres0 = sqrt(rr);
        normr0 = fmax(1e-20, absres0);
        relres = absres0 / normr0;
        temp1 = rr;
    }
    if (relres < a.tol) { if (tid == 0) a.status[0] = 0; return; }
    for (int i = tid; i < m; i += nthr) __stcg(a.p + i, __ldcg(a.r + i));
    grid.sync();

    bool stopped = false;                                    // left the loop through a break (no restore needed when iter == iter_best)
    while (iter++ < a.maxit) {
        cg_spmv<KIND, false>(a.A, a.p, nullptr, a.t, sprod);  // t = A p
        grid.sync();
        if (blockIdx.x == 0) { const double tp = cg_dot_seq(m, a.t, a.p, prod); if (threadIdx.x == 0) __stcg(a.scal, tp); }
        grid.sync();
        temp2 = __ldcg(a.scal);
        if (!(fabs(temp2) > 1e-40)) break;                    // SSS_cycle.cu:190 -> restore
        alpha = __ddiv_rn(temp1, temp2);
        {
            const double nalpha = -alpha;
            double uu = 0.0, pp = 0.0, um = 0.0;
            for (int i = tid; i < m; i += nthr) {
                const double pi = __ldcg(a.p + i);
                const double ui = __dadd_rn(__ldcg(a.u + i), __dmul_rn(alpha, pi));
                __stcg(a.r + i, __dadd_rn(__ldcg(a.r + i), __dmul_rn(nalpha, __ldcg(a.t + i))));
                __stcg(a.u + i, ui);
                uu = __dadd_rn(uu, __dmul_rn(ui, ui));
                pp = __dadd_rn(pp, __dmul_rn(pi, pi));
                um = fmax(um, fabs(ui));
            }
            uu = block_sum(uu, red); if (threadIdx.x == 0) __stcg(a.partial + 0 * gridDim.x + blockIdx.x, uu);
            pp = block_sum(pp, red); if (threadIdx.x == 0) __stcg(a.partial + 1 * gridDim.x + blockIdx.x, pp);
            um = block_max(um, red); if (threadIdx.x == 0) __stcg(a.partial + 2 * gridDim.x + blockIdx.x, um);
        }
        grid.sync();
        if (blockIdx.x == 0) {
            const double rr = cg_dot_seq(m, a.r, a.r, prod);
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
                v0 = __dadd_rn(v0, __ldcg(a.partial + i));
                v1 = __dadd_rn(v1, __ldcg(a.partial + gridDim.x + i));
                v2 = fmax(v2, __ldcg(a.partial + 2 * gridDim.x + i));
            }
            v0 = block_sum(v0, red); v1 = block_sum(v1, red); v2 = block_max(v2, red);
            if (threadIdx.x == 0) { __stcg(a.scal, rr); __stcg(a.scal + 1, v0); __stcg(a.scal + 2, v1); __stcg(a.scal + 3, v2); }
        }
        grid.sync();
        const double rr = __ldcg(a.scal), uu = __ldcg(a.scal + 1), pp = __ldcg(a.scal + 2), infnormu = __ldcg(a.scal + 3);
        absres = sqrt(rr);
        relres = absres / normr0;
        if (absres < absres_best - maxdiff) {
            absres_best = absres;
            iter_best = iter;
            for (int i = tid; i < m; i += nthr) __stcg(a.u_best + i, __ldcg(a.u + i));      // (the thread that wrote u_i copies it)
        }
        if (infnormu <= 1e-20) { if (tid == 0) a.status[0] = CG_FALLBACK; return; }
        const double reldiff = fabs(alpha) * sqrt(pp) / sqrt(uu);
        if (reldiff < maxdiff) { if (tid == 0) a.status[0] = CG_FALLBACK; return; }         // stagnation handling: host
        if (relres < a.tol) {
            grid.sync();                                                                      // u complete everywhere
            cg_spmv<KIND, true>(a.A, a.u, a.b, a.r, sprod);
            grid.sync();
            if (blockIdx.x == 0) { const double r2 = cg_dot_seq(m, a.r, a.r, prod); if (threadIdx.x == 0) __stcg(a.scal, r2); }
            grid.sync();
            absres = sqrt(__ldcg(a.scal));
            relres = absres / normr0;
            if (relres < a.tol) { stopped = true; break; }
            if (tid == 0) a.status[0] = CG_FALLBACK;                                         // false convergence: host
            return;
        }
        temp2 = rr;
        if (a.beta_fix) { beta = __ddiv_rn(temp2, temp1); temp1 = temp2; }
        else beta = __ddiv_rn(temp1, temp1);
        for (int i = tid; i < m; i += nthr) __stcg(a.p + i, __dadd_rn(__dmul_rn(1.0, __ldcg(a.r + i)), __dmul_rn(beta, __ldcg(a.p + i))));
        grid.sync();
    }
    (void)stopped;
    // restore the best iterate seen if the last one is worse (SSS_cycle.cu:400-420)
    if (iter != iter_best) {
        grid.sync();
        cg_spmv<KIND, true>(a.A, a.u_best, a.b, a.r, sprod);
        grid.sync();
        if (blockIdx.x == 0) { const double r2 = cg_dot_seq(m, a.r, a.r, prod); if (threadIdx.x == 0) __stcg(a.scal, r2); }
        grid.sync();
        absres_best = sqrt(__ldcg(a.scal));
        if (absres > absres_best + maxdiff)
            for (int i = tid; i < m; i += nthr) __stcg(a.u + i, __ldcg(a.u_best + i));
    }
    if (tid == 0) a.status[0] = iter > a.maxit ? -48 : iter;
}

// y = a*x + b*y   (SSS_blas_array_axpby, SSS_utils.c:248-253)
__global__ void __launch_bounds__(BLOCK) axpby_kernel(int n, double a, const double *__restrict__ x, double b, double *y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        y[i] = __dadd_rn(__dmul_rn(a, x[i]), __dmul_rn(b, y[i]));
}
// y += a*x   (SSS_blas_array_axpy, :217-222); x may alias y
__global__ void __launch_bounds__(BLOCK) axpy_kernel(int n, double a, const double *x, double *y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i]));
}
// x *= a   (SSS_blas_array_ax, :255-260)
__global__ void __launch_bounds__(BLOCK) scale_kernel(int n, double a, double *x) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = __dmul_rn(x[i], a);
}
// partial[block] = tree sum of x_i*y_i (norms that do not feed back into x)
__global__ void __launch_bounds__(BLOCK) dot_tree_kernel(int n, const double *__restrict__ x, const double *__restrict__ y, double *partial) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s = __dadd_rn(s, __dmul_rn(x[i], y[i]));
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// permutations between natural and schedule numbering
// test hook: gs_quotient_pre against __ddiv_rn, bit for bit, on generated operand pairs (mode 0: random significands,
// moderate exponents; 1: raw 64-bit patterns; 2: t within a few ulps of q*d for random q -- quotients next to
// representable numbers and rounding midpoints; 3: divisors next to powers of two and with saturated significands)
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long &s) {
    unsigned long long z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(BLOCK) quotient_check_kernel(long long n, unsigned long long seed, int mode, unsigned long long *mismatch) {
    unsigned long long bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long s = seed + 0x632be59bd9b4e019ULL * (unsigned long long)(i + 1);
        const unsigned long long a = splitmix64(s), b = splitmix64(s), c = splitmix64(s);
        double t, d;
        auto mk = [](unsigned long long bits, int e) { return __longlong_as_double((long long)((bits & 0x800fffffffffffffULL) | ((unsigned long long)(1023 + e) << 52))); };
        if (mode == 0) { t = mk(a, (int)(c % 121) - 60); d = mk(b, (int)((c >> 8) % 41) - 20); }
        else if (mode == 1) { t = __longlong_as_double((long long)a); d = __longlong_as_double((long long)b); }
        else if (mode == 2) {
            const double q = mk(a, (int)(c % 61) - 30);
            d = mk(b, (int)((c >> 8) % 41) - 20);
            t = __dmul_rn(q, d);
            t = __longlong_as_double(__double_as_longlong(t) + (long long)((c >> 20) % 5) - 2);
        } else {
            const unsigned long long lowbits = (c >> 12) % 3 == 0 ? 0ULL : ((c >> 12) % 3 == 1 ? 0x000fffffffffffffULL : 0x000ffffffffffff0ULL);
            d = mk((b & 0x8000000000000000ULL) | ((lowbits + (c & 7)) & 0x000fffffffffffffULL), (int)((c >> 8) % 41) - 20);
            t = mk(a, (int)(c % 121) - 60);
        }
        const double ref = __ddiv_rn(t, d);
        const double got = gs_quotient_pre(t, d, __ddiv_rn(1.0, d), gs_quotient_dsafe(d), 0);
        const bool same = __double_as_longlong(ref) == __double_as_longlong(got) || (ref != ref && got != got);
        if (!same) ++bad;
    }
    if (bad) atomicAdd(mismatch, bad);
}

// Device-side construction of a SELL-32 layout from the raw CSR arrays of the host hierarchy (one warp per slice, lane = row):
// rows taken in `order` (schedule position -> natural row, or identity), columns renumbered through `col_pos`, padding
// col = -1 / val = +0.0, row-internal storage order untouched.  Writes are fully coalesced; the host only computes the O(rows)
// slice table (analysis.cpp: build_sell_structure).
__global__ void __launch_bounds__(BLOCK) sell_fill_kernel(int nslices, const int *__restrict__ slice_row, const long long *__restrict__ slice_ptr,
                                                           const int *__restrict__ order, const int *__restrict__ col_pos,
                                                           const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va,
                                                           int *__restrict__ col, double *__restrict__ val) {
    const int lane = threadIdx.x & 31;
    for (int s = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); s < nslices; s += gridDim.x * WARPS_PER_BLOCK) {
        const long long base = slice_ptr[s];
        const int width = (int)((slice_ptr[s + 1] - base) >> 5);
        const int r0 = slice_row[s], nr = slice_row[s + 1] - r0;
        int p0 = 0, len = 0;
        if (lane < nr) {
            const int i = order ? order[r0 + lane] : r0 + lane;
            p0 = rp[i]; len = rp[i + 1] - p0;
        }
        for (int e = 0; e < width; ++e) {
            int j = -1;
            double a = 0.0;
            if (e < len) { j = ci[p0 + e]; a = va[p0 + e]; if (col_pos) j = col_pos[j]; }
            col[base + 32LL * e + lane] = j;
            val[base + 32LL * e + lane] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Peer-memory exchange (multi-GPU, one process per GPU): ghost entries are STORED straight into the owning vector of the peer
// (mapped through CUDA IPC, NVLink / NVSwitch P2P), then an epoch is written to the peer's flag word for this rank; the peer's
// stream waits on its flag words.  No host round trip, no collective.
// ------------------------------------------------------------------------------------------
struct PeerPush {              // copy src[idx[i]] -> dst[idx[i]] (idx != nullptr) or src[i] -> dst[i], i in [0, count)
    const double *src;
    double *dst;
    const int *idx;
    int count;
};
__global__ void __launch_bounds__(BLOCK) peer_push_kernel(const PeerPush *__restrict__ d) {
    const PeerPush p = d[blockIdx.y];
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < p.count; i += gridDim.x * BLOCK) {
        const int k = p.idx ? p.idx[i] : i;
        p.dst[k] = p.src[k];
    }
}
// (runs after peer_push_kernel in stream order: its stores have been performed; the fence orders them before the flag at system scope)
__global__ void peer_flag_kernel(unsigned *const *__restrict__ flag, int n, unsigned epoch) {
    if ((int)threadIdx.x < n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag[threadIdx.x]), "r"(epoch) : "memory");
    }
}
__global__ void peer_wait_kernel(const unsigned *__restrict__ flags, const int *__restrict__ src, int n, unsigned epoch) {
    if ((int)threadIdx.x < n) {
        unsigned f;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(flags + src[threadIdx.x]) : "memory");
        } while ((int)(f - epoch) < 0);
    }
}

__global__ void __launch_bounds__(BLOCK) gather_kernel(int n, const int *__restrict__ order, const double *__restrict__ in, double *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[order[k]];
}
__global__ void __launch_bounds__(BLOCK) scatter_kernel(int n, const int *__restrict__ order, const double *__restrict__ in, double *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[order[k]] = in[k];
}
// the same two moves between the forward and the backward numbering of a natural-order level
__global__ void __launch_bounds__(BLOCK) gather_idx_kernel(int n, const int *__restrict__ idx, const double *__restrict__ in, double *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[idx[k]];
}
__global__ void __launch_bounds__(BLOCK) scatter_idx_kernel(int n, const int *__restrict__ idx, const double *__restrict__ in, double *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[idx[k]] = in[k];
}

}  // namespace amgb200
