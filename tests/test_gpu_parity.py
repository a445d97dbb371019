"""GPU parity tests: every call goes through the C ABI of libamgb200.so and is compared with the
CPU oracle (oracle/amg_oracle.c, pinned bit-for-bit against the reference's own objects) on the
same inputs.

Bars (EXACT mode, the default):
  * every vector the path produces -- SpMV, residual, restriction, prolongation, Gauss-Seidel
    sweeps, coarsest-level CG/GMRES, whole V-cycles, the final solution -- is BIT-IDENTICAL to
    the oracle: all kernels accumulate in the reference's order without FMA;
  * ||r|| values are tree-reduced (they never feed back into x): within 1e-13 relative;
  * whole solves: same V-cycle count, residual history within 1e-10 relative per iteration
    (the tolerance BASELINE.json's north_star states; measured deviation is ~1e-15).
FAST mode (opt-in) is checked against looser, documented bounds.
"""
import numpy as np
import pytest

import oracle_ffi
from amg_b200 import DeviceHierarchy, HostHierarchy, capi, generate, solve_dropin

pytestmark = pytest.mark.gpu

RTOL_VECTOR = 1e-13
RTOL_HISTORY = 1e-10

CASES = {
    "p2d64": ("p2d", 64, 0.0),
    "p2d256": ("p2d", 256, 0.0),
    "p3d16": ("p3d", 16, 0.0),
    "p3d32": ("p3d", 32, 0.0),
    "aniso32": ("aniso3d", 32, 1e-3),
    "v27_12": ("v27", 12, 0.0),
    "v27_16": ("v27", 16, 0.0),
}

_cache = {}


def case(name, tol=1e-8):
    key = (name, tol)
    if key not in _cache:
        kind, N, eps = CASES[name]
        A = generate(kind, N, eps)
        hier = HostHierarchy(A, tol=tol)
        dev = DeviceHierarchy(hier)
        _cache[key] = (A, hier, dev)
    return _cache[key]


def rel_err(a, b):
    d = np.abs(a - b).max()
    s = np.abs(b).max()
    return d / s if s > 0 else d


def check_vec(dev, l, got, want, what):
    kind = "SELL thread/row" if dev.info(l)["kind"] == 0 else "CSR warp/row"
    assert got.tobytes() == want.tobytes(), f"{what}: level {l} ({kind}) not bit-identical, rel err {rel_err(got, want):.3e}"


def rng_vec(n, seed):
    return np.random.default_rng(seed).standard_normal(n)


@pytest.mark.parametrize("name", ["p2d64", "p3d16", "aniso32", "v27_12"])
def test_spmv_all_levels(name, oracle):
    """y = A x, y += alpha A x on every level; restriction y = R x; prolongation y += P x
    (amg/SSS_utils.c:161-201 as called from SSS_cycle.cu:917,921,942)"""
    A, hier, dev = case(name)
    for l in range(hier.num_levels):
        c = hier.level(l)
        n = c.A.num_rows
        x = rng_vec(n, 10 + l)
        y0 = rng_vec(n, 20 + l)
        check_vec(dev, l, dev.spmv(l, "A", x), oracle.mxy(c.A, x), "mxy A")
        check_vec(dev, l, dev.spmv(l, "A", x, y0, alpha=-1.0), oracle.amxpy(-1.0, c.A, x, y0), "amxpy A")
        if l < hier.num_levels - 1:
            nc = c.P.num_cols
            xc = rng_vec(nc, 30 + l)
            assert dev.spmv(l, "R", x).tobytes() == oracle.mxy(c.R, x).tobytes(), f"restriction level {l}"
            assert dev.spmv(l, "P", xc, y0, alpha=1.0).tobytes() == oracle.amxpy(1.0, c.P, xc, y0).tobytes(), f"prolongation level {l}"


@pytest.mark.parametrize("name", ["p2d64", "p3d16", "p3d32", "aniso32", "v27_12"])
@pytest.mark.parametrize("sweeps", [1, 2])
def test_gauss_seidel_cf_all_levels(name, sweeps, oracle):
    """C/F-ordered Gauss-Seidel (amg/Solve/SSS_smooth.c:4-87) on every smoothed level, from a
    random state: fully parallel two-colour levels and ordered (wavefront) levels alike"""
    A, hier, dev = case(name)
    for l in range(hier.num_levels - 1):
        c = hier.level(l)
        n = c.A.num_rows
        x0 = rng_vec(n, 70 + l)
        b = rng_vec(n, 80 + l)
        got = dev.smooth(l, sweeps, x0, b)
        want = oracle.gs_cf(c.A, hier.cfmark(l), x0, b, sweeps, 1)
        check_vec(dev, l, got, want, f"GS x{sweeps}")


@pytest.mark.parametrize("strategy,kernel", [(None, "gs_stream_cta_kernel"), ("2", "gs_ordered_cta_kernel"), ("3", "gs_ordered_cluster_kernel")])
@pytest.mark.parametrize("name", ["p3d32", "v27_16"])
def test_every_ordered_smoother_kernel_is_bit_identical(name, strategy, kernel, oracle, monkeypatch):
    """the barrier-per-wavefront launch strategies for ordered levels -- streaming single CTA (default where x fits in
    shared memory), single CTA, 16-CTA cluster -- all reproduce the sequential sweep bit for bit,
    for 1, 2 and 3 sweeps per launch (3 sweeps wrap the shared-memory ring and the mbarrier phases several times)"""
    if strategy is None:
        monkeypatch.delenv("AMGB200_GS_STRATEGY", raising=False)
    else:
        monkeypatch.setenv("AMGB200_GS_STRATEGY", strategy)
    kind, N, eps = CASES[name]
    hier = HostHierarchy(generate(kind, N, eps), tol=1e-8)
    dev = DeviceHierarchy(hier)
    used = set()
    for l in range(hier.num_levels - 1):
        c = hier.level(l)
        n = c.A.num_rows
        used.add(dev.gs_kernel(l))
        for sweeps in (1, 2, 3):
            x0 = rng_vec(n, 170 + l)
            b = rng_vec(n, 180 + l)
            got = dev.smooth(l, sweeps, x0, b)
            want = oracle.gs_cf(c.A, hier.cfmark(l), x0, b, sweeps, 1)
            check_vec(dev, l, got, want, f"GS x{sweeps} ({dev.gs_kernel(l)})")
    assert kernel in used, f"{kernel} was not exercised (kernels used: {sorted(used)})"


@pytest.mark.parametrize("per_sm", [None, "1"])
@pytest.mark.parametrize("name", ["p2d64", "p3d32", "aniso32", "v27_16"])
def test_dataflow_smoother_is_bit_identical(name, per_sm, oracle, monkeypatch):
    """the data-flow (sync-free) smoother of the wide thread-per-row levels -- every x_k travels as a {value, version} record,
    rows poll the records of their columns, no wavefront barrier -- reproduces the sequential sweep (SSS_smooth.c:16-48) bit for
    bit for 1, 2 and 3 sweeps per launch, with the grid the occupancy allows and with one CTA per SM (different item-to-warp
    mapping, wraps around the sweeps differently); AMGB200_DF_ALL=1 also sends the levels the streaming kernels would take to it"""
    monkeypatch.setenv("AMGB200_DF_ALL", "1")
    monkeypatch.setenv("AMGB200_DFW_MIN_WIDTH", "4")           # (the warp-per-row form for every level with >= 4 rows per wavefront)
    if per_sm:
        monkeypatch.setenv("AMGB200_DF_PER_SM", per_sm)
    kind, N, eps = CASES[name]
    hier = HostHierarchy(generate(kind, N, eps), tol=1e-8)
    dev = DeviceHierarchy(hier)
    used = set()
    for l in range(hier.num_levels - 1):
        c = hier.level(l)
        n = c.A.num_rows
        used.add(dev.gs_kernel(l))
        for sweeps in (1, 2, 3):
            x0 = rng_vec(n, 270 + l)
            b = rng_vec(n, 280 + l)
            got = dev.smooth(l, sweeps, x0, b)
            want = oracle.gs_cf(c.A, hier.cfmark(l), x0, b, sweeps, 1)
            check_vec(dev, l, got, want, f"GS x{sweeps} ({dev.gs_kernel(l)})")
    assert "gs_dataflow_kernel" in used, f"data-flow kernel not exercised (kernels used: {sorted(used)})"
    if name in ("p3d32", "aniso32"):      # (the other two hierarchies have no warp-per-row level that does not fit the single-SM streaming kernel)
        assert "gs_dataflow_csr_kernel" in used, f"warp-per-row data-flow kernel not exercised (kernels used: {sorted(used)})"
    n0 = hier.level(0).A.num_rows
    rtn, x, hist = dev.solve(np.ones(n0), np.ones(n0))
    rtn_o, x_o, hist_o = oracle.solve(hier, np.ones(n0), np.ones(n0), 0)
    assert rtn.nits == rtn_o.nits and x.tobytes() == x_o.tobytes()
    dev.close()
    hier.close()


@pytest.mark.parametrize("name", ["p2d64", "p3d16", "v27_12"])
def test_residual_and_norm(name, oracle):
    A, hier, dev = case(name)
    for l in range(hier.num_levels):
        c = hier.level(l)
        n = c.A.num_rows
        x = rng_vec(n, 90 + l)
        b = rng_vec(n, 95 + l)
        r, nrm = dev.residual(l, x, b)
        want = oracle.amxpy(-1.0, c.A, x, b)
        check_vec(dev, l, r, want, "residual")
        assert abs(nrm - np.sqrt(np.sum(want * want))) <= 1e-13 * nrm


@pytest.mark.parametrize("name", ["p2d64", "p2d256", "p3d16", "p3d32", "aniso32", "v27_12"])
@pytest.mark.parametrize("env", [{}, {"AMGB200_RR_ALL": "1"}, {"AMGB200_RR_ALL": "1", "AMGB200_RR_CHUNKS": "7", "AMGB200_RR_LAG": "0"},
                                 {"AMGB200_RR_CHUNKS": "300", "AMGB200_RR_LAG": "2", "AMGB200_RR_PER_SM": "1"}, {"AMGB200_NO_FUSED": "1"}])
def test_fused_residual_restriction(name, env, oracle, monkeypatch):
    """r = b - A x and b_{l+1} = R r (amg/Solve/SSS_cycle.cu:916-921) in ONE launch (resid_restrict_kernel) on every level whose A and
    R are thread-per-row layouts, the separate kernels elsewhere: both vectors bit-identical to the oracle, repeated launches
    (the completion counters are never reset) included"""
    monkeypatch.setenv("AMGB200_RR_MIN_ROWS", "0")                # (by default only levels of >= 262 144 rows take it)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    A, hier, _ = case(name)
    dev = DeviceHierarchy(hier)
    fused_levels = 0
    for l in range(hier.num_levels - 1):
        c = hier.level(l)
        n = c.A.num_rows
        for rep in range(3):
            x = rng_vec(n, 130 + 7 * l + rep)
            b = rng_vec(n, 140 + 7 * l + rep)
            r, bc, fused = dev.resid_restrict(l, x, b)
            want_r = oracle.amxpy(-1.0, c.A, x, b)
            check_vec(dev, l, r, want_r, "residual (fused)" if fused else "residual")
            assert bc.tobytes() == oracle.mxy(c.R, want_r).tobytes(), f"restricted residual, level {l}, fused={fused}"
        fused_levels += fused
    if "AMGB200_NO_FUSED" in env:
        assert fused_levels == 0
    elif name != "v27_12" or "AMGB200_RR_ALL" in env:             # (27-point rows need the long-row instances: opt-in)
        assert fused_levels >= 1, "level 0 must take the fused launch"
    dev.close()


@pytest.mark.parametrize("name", ["p2d64", "p3d16", "aniso32", "v27_12"])
@pytest.mark.parametrize("mode", [0, 1])
def test_coarse_solve(name, mode, oracle):
    """coarsest-level CG (+ GMRES fallback, reached in AS_COMPILED mode): same Krylov status /
    iteration counts as the oracle and a bit-identical solution"""
    A, hier, _ = case(name)
    dev = DeviceHierarchy(hier, coarse_mode=mode)
    c = hier.level(hier.num_levels - 1)
    n = c.A.num_rows
    b = rng_vec(n, 7)
    x0 = np.zeros(n)
    st, x, its = dev.coarse_solve(x0, b, 1e-9)
    st_o, x_o, its_o = oracle.coarse_solve(c.A, x0, b, 1e-9, mode)
    assert its == its_o, f"Krylov iteration counts differ: {its} vs {its_o}"
    assert st == st_o
    assert x.tobytes() == x_o.tobytes(), f"coarse solution not bit-identical: rel err {rel_err(x, x_o):.3e}"
    dev.close()


@pytest.mark.parametrize("name", ["p2d64", "p3d16", "aniso32", "v27_12"])
def test_one_vcycle(name, oracle):
    A, hier, dev = case(name)
    n = A.nrows
    x0 = np.ones(n)
    b = np.ones(n)
    got = dev.cycle(x0, b)
    want = oracle.cycle(hier, x0, b, 0)
    assert got.tobytes() == want.tobytes(), f"V-cycle result not bit-identical: rel err {rel_err(got, want):.3e}"


@pytest.mark.parametrize("name", list(CASES))
def test_solve_history(name, oracle):
    """full solve, tol 1e-8, b = 1, x0 = 1 (amg/SSS_main.c:141-145): same V-cycle count, residual
    history within 1e-10 relative per iteration, bit-identical solution"""
    A, hier, dev = case(name)
    n = A.nrows
    x0 = np.ones(n)
    b = np.ones(n)
    rtn, x, hist = dev.solve(x0, b)
    rtn_o, x_o, hist_o = oracle.solve(hier, x0, b, 0)
    assert rtn.nits == rtn_o.nits, f"V-cycle counts differ: {rtn.nits} vs {rtn_o.nits}"
    rel = np.abs(hist - hist_o) / hist_o
    assert rel.max() <= RTOL_HISTORY, f"residual history deviates: {rel}"
    assert abs(rtn.ares - rtn_o.ares) <= RTOL_HISTORY * rtn_o.ares
    assert abs(rtn.rres - rtn_o.rres) <= RTOL_HISTORY * rtn_o.rres
    assert x.tobytes() == x_o.tobytes(), f"solution not bit-identical: rel err {rel_err(x, x_o):.3e}"
    # independent property: the returned x really has that residual
    r = b - A.matvec(x)
    assert abs(np.linalg.norm(r) - rtn.ares) <= 1e-9 * np.linalg.norm(b)


@pytest.mark.parametrize("name", ["p2d64", "p3d16"])
def test_solve_history_as_compiled_mode(name, oracle):
    A, hier, _ = case(name)
    dev = DeviceHierarchy(hier, coarse_mode=1)
    n = A.nrows
    rtn, x, hist = dev.solve(np.ones(n), np.ones(n))
    rtn_o, x_o, hist_o = oracle.solve(hier, np.ones(n), np.ones(n), 1)
    assert rtn.nits == rtn_o.nits
    assert (np.abs(hist - hist_o) / hist_o).max() <= RTOL_HISTORY
    assert x.tobytes() == x_o.tobytes()
    dev.close()


@pytest.mark.parametrize("name", ["p2d64", "p3d32", "v27_16"])
def test_fast_mode_bounds(name, oracle):
    """FAST mode (tree reductions on long rows and Krylov dots): same V-cycle count, solution
    equal to 1e-10 relative, residual history within 1e-5 relative (the |x|/|r| amplification of
    1e-16 rounding differences; documented in DESIGN.md)"""
    A, hier, _ = case(name)
    dev = DeviceHierarchy(hier, fast=1)
    n = A.nrows
    rtn, x, hist = dev.solve(np.ones(n), np.ones(n))
    rtn_o, x_o, hist_o = oracle.solve(hier, np.ones(n), np.ones(n), 0)
    assert rtn.nits == rtn_o.nits
    assert (np.abs(hist - hist_o) / hist_o).max() <= 1e-5
    assert rel_err(x, x_o) <= 1e-10
    dev.close()


def test_dropin_sss_amg_solve(oracle, capfd):
    """the reference-facing entry point: host hierarchy in, x overwritten, table printed"""
    A = generate("p2d", 64)
    hier = HostHierarchy(A, tol=1e-8)
    n = A.nrows
    rtn, x = solve_dropin(hier, np.ones(n), np.ones(n))
    out = capfd.readouterr().out
    hier2 = HostHierarchy(A, tol=1e-8)
    rtn_o, x_o, hist_o = oracle.solve(hier2, np.ones(n), np.ones(n), 0)
    assert rtn.nits == rtn_o.nits
    assert abs(rtn.ares - rtn_o.ares) <= RTOL_HISTORY * rtn_o.ares
    assert x.tobytes() == x_o.tobytes()
    assert hier.mg.rtn.nits == rtn.nits
    assert "It Num |   ||r||/||b||   |     ||r||      |  Conv. Factor" in out
    assert "AMG solve time:" in out
    assert out.count(" | ") >= 2 * (rtn.nits + 1)


def test_zero_rhs_returns_zero_solution():
    """||b|| = 0 => x = 0, zero iterations (amg/Solve/SSS_SOLVE.c:41-46)"""
    A, hier, dev = case("p2d64")
    rtn, x, hist = dev.solve(np.ones(A.nrows), np.zeros(A.nrows))
    assert rtn.nits == 0 and rtn.ares == 0 and not x.any()


def test_w_cycle(oracle):
    """cycle_type = 2 exercises the num_lvl bookkeeping of SSS_cycle.cu:960-966"""
    A = generate("p2d", 64)
    hier = HostHierarchy(A, tol=1e-8, cycle_type=2)
    dev = DeviceHierarchy(hier)
    n = A.nrows
    rtn, x, hist = dev.solve(np.ones(n), np.ones(n))
    rtn_o, x_o, hist_o = oracle.solve(hier, np.ones(n), np.ones(n), 0)
    assert rtn.nits == rtn_o.nits
    assert (np.abs(hist - hist_o) / hist_o).max() <= RTOL_HISTORY


@pytest.mark.parametrize("name", ["p2d64", "p3d16", "v27_12"])
def test_natural_order_gauss_seidel(name, oracle):
    """cf_order = 0: forward sweeps before, backward sweeps after the correction, x = t * (1/d)
    (amg/Solve/SSS_smooth.c:90-137, reached through :176 and :261) -- every level, then whole solves"""
    kind, N, eps = CASES[name]
    A = generate(kind, N, eps)
    hier = HostHierarchy(A, tol=1e-8, cf_order=0)
    dev = DeviceHierarchy(hier)
    for l in range(hier.num_levels - 1):
        c = hier.level(l)
        n = c.A.num_rows
        x0, b = rng_vec(n, 200 + l), rng_vec(n, 300 + l)
        for sweeps, backward in ((2, False), (2, True), (1, True)):
            got = dev.smooth(l, -sweeps if backward else sweeps, x0, b)
            want = oracle.gs_natural(c.A, x0, b, sweeps, backward)
            assert got.tobytes() == want.tobytes(), f"natural GS level {l} backward={backward}: rel err {rel_err(got, want):.3e}"
    n = A.nrows
    rtn, x, hist = dev.solve(np.ones(n), np.ones(n))
    rtn_o, x_o, hist_o = oracle.solve(hier, np.ones(n), np.ones(n), 0)
    assert rtn.nits == rtn_o.nits
    assert (np.abs(hist - hist_o) / hist_o).max() <= RTOL_HISTORY
    assert x.tobytes() == x_o.tobytes()
    dev.close()


def test_natural_order_oracle_matches_reference(oracle, reference):
    """pins the oracle's natural-order path against the reference's own objects"""
    import ctypes as C
    A = generate("p3d", 10)
    pars = capi.default_pars(1e-8)
    pars.cf_order = 0
    mg = reference.setup(A, pars)
    n = A.nrows
    x_ref, hist_ref = reference.solve_history(mg, np.ones(n), np.ones(n), 1e-8)
    reference.destroy(mg)
    hier = HostHierarchy(A, tol=1e-8, cf_order=0)
    rtn, x, hist = oracle.solve(hier, np.ones(n), np.ones(n), 0)
    assert list(hist) == list(hist_ref) and x.tobytes() == x_ref.tobytes()


def test_random_rhs_and_initial_guess(oracle):
    A, hier, dev = case("p3d16")
    n = A.nrows
    b = rng_vec(n, 123)
    x0 = rng_vec(n, 321)
    rtn, x, hist = dev.solve(x0, b)
    rtn_o, x_o, hist_o = oracle.solve(hier, x0, b, 0)
    assert rtn.nits == rtn_o.nits
    assert (np.abs(hist - hist_o) / hist_o).max() <= RTOL_HISTORY


def test_hierarchy_info_matches_host():
    A, hier, dev = case("p3d32")
    assert dev.num_levels == hier.num_levels
    for l, (rows, nnz) in enumerate(hier.table()):
        info = dev.info(l)
        assert info["rows"] == rows and info["nnz"] == nnz
    # level 0 of 7-point Poisson is an exact red/black split: both passes are one wavefront
    assert dev.info(0)["wf_F"] == 1 and dev.info(0)["wf_C"] == 1


def test_linearity_of_spmv_property():
    """size-independent property on a larger operator: A(ax + y) = a Ax + Ay to rounding"""
    A = generate("p3d", 48)
    hier = HostHierarchy(A, tol=1e-8)
    dev = DeviceHierarchy(hier)
    n = A.nrows
    x, y = rng_vec(n, 1), rng_vec(n, 2)
    lhs = dev.spmv(0, "A", 2.5 * x + y)
    rhs = 2.5 * dev.spmv(0, "A", x) + dev.spmv(0, "A", y)
    assert rel_err(lhs, rhs) <= 1e-13
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_quotient_fast_path(mode):
    """the ordered smoothers' quotient with the reciprocal taken off the dependency path (kernels.cuh: gs_quotient_pre)
    is bit-identical to IEEE division (SSS_smooth.c:32) on 2^26 generated operand pairs per mode: random operands, raw bit
    patterns (incl. zeros, subnormals, infinities, NaNs), quotients within 2 ulps of representable numbers / midpoints,
    divisors next to powers of two and with saturated significands"""
    import ctypes as C
    L = capi.lib()
    L.amgb200_debug_quotient_check.restype = C.c_longlong
    L.amgb200_debug_quotient_check.argtypes = [C.c_longlong, C.c_ulonglong, C.c_int]
    for seed in (1, 2026):
        assert L.amgb200_debug_quotient_check(1 << 26, seed, mode) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", [("p2d", 96, 0.0), ("p3d", 40, 0.0), ("aniso3d", 32, 1e-3), ("v27", 20, 0.0), ("p3d", 128, 0.0)])
def test_device_interpolation_builds_the_same_hierarchy(case):
    """direct-interpolation weights, coarse renumbering and truncation of P on the device (amgb200_interp_device, setup_dev.cu) against
    the host loop that is byte-pinned to the reference's interp_DIR + SSS_amg_interp_trunc (Setup/SSS_inter.cu:400-547, :16-102;
    tests/test_setup_parity.py): every array of every level of the resulting hierarchy (A, P, R, cfmark) is byte-identical -- P feeds
    R = P^T and the Galerkin product, so one differing bit would show up on every coarser level.  For the sizes with a fixture the
    sha256 of P is also compared with the one of the reference's own setup (tests/golden/golden.json)."""
    import hashlib
    import json
    import os
    kind, N, eps = case
    A = generate(kind, N, eps)
    host = HostHierarchy(A, tol=1e-8)
    devh = HostHierarchy(A, tol=1e-8, device_interp=True)
    assert host.table() == devh.table()
    for l in range(host.num_levels):
        for which in ("A", "P", "R"):
            if which != "A" and l == host.num_levels - 1:
                continue
            a, b = host.level_matrix(l, which), devh.level_matrix(l, which)
            assert a.row_ptr.tobytes() == b.row_ptr.tobytes() and a.col_idx.tobytes() == b.col_idx.tobytes() and a.val.tobytes() == b.val.tobytes(), (l, which)
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))
    name = f"{kind}{N}"
    if name in gold:
        for l in range(devh.num_levels - 1):
            P = devh.level_matrix(l, "P")
            assert hashlib.sha256(P.val.tobytes()).hexdigest() == gold[name]["levels"][l]["P"]["val"]
            assert hashlib.sha256(P.col_idx.tobytes()).hexdigest() == gold[name]["levels"][l]["P"]["col_idx"]
    host.close(); devh.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [("p2d", 96, 0.0), ("p2d", 256, 0.0), ("p3d", 40, 0.0), ("aniso3d", 32, 1e-3), ("v27", 20, 0.0), ("p3d", 128, 0.0)])
def test_device_transpose_and_galerkin_build_the_same_hierarchy(case):
    """R = P^T and A_{l+1} = R A P on the device (amgb200_rap_device, setup_rap.cu) against the host loops that are byte-pinned to the
    reference's SSS_mat_trans / SSS_blas_mat_rap (SSS_matvec.c:330-387, :398-534; tests/test_setup_parity.py): every array of every level
    is byte-identical -- row order of R, diagonal-first discovery order and accumulation order of the product included, since they are
    the summation order of the solve phase.  For the sizes with a fixture the sha256 of every A and R array is also compared with the
    one of the reference's own setup (tests/golden/golden.json)."""
    import hashlib
    import json
    import os
    kind, N, eps = case
    A = generate(kind, N, eps)
    host = HostHierarchy(A, tol=1e-8)
    devh = HostHierarchy(A, tol=1e-8, device_rap=True, device_interp=True)
    assert host.table() == devh.table()
    for l in range(host.num_levels):
        for which in ("A", "P", "R"):
            if which != "A" and l == host.num_levels - 1:
                continue
            a, b = host.level_matrix(l, which), devh.level_matrix(l, which)
            assert a.row_ptr.tobytes() == b.row_ptr.tobytes(), (l, which, "row_ptr")
            assert a.col_idx.tobytes() == b.col_idx.tobytes(), (l, which, "col_idx")
            assert a.val.tobytes() == b.val.tobytes(), (l, which, "val")
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))
    name = f"{kind}{N}"
    if name in gold:
        for l in range(devh.num_levels):
            for which in ("A", "R"):
                if which not in gold[name]["levels"][l]:
                    continue
                M = devh.level_matrix(l, which)
                for arr in ("row_ptr", "col_idx", "val"):
                    assert hashlib.sha256(getattr(M, arr).tobytes()).hexdigest() == gold[name]["levels"][l][which][arr], (l, which, arr)
    host.close(); devh.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["p3d16", "v27_12"])
def test_host_and_device_sell_fill_agree(name, oracle, monkeypatch):
    """SELL-32 layouts permuted/padded on the device (sell_fill_kernel, the default) and on the host
    (AMGB200_HOST_LAYOUT=1, analysis.cpp build_layout) give bit-identical SpMV on every level and the same solve"""
    A, hier, dev = case(name)
    monkeypatch.setenv("AMGB200_HOST_LAYOUT", "1")
    dev_h = DeviceHierarchy(hier)
    monkeypatch.delenv("AMGB200_HOST_LAYOUT")
    for l in range(dev.num_levels):
        n = dev.info(l)["rows"]
        x = rng_vec(n, 7 + l)
        assert dev.spmv(l, "A", x).tobytes() == dev_h.spmv(l, "A", x).tobytes()
        if l + 1 < dev.num_levels:
            nc = dev.info(l + 1)["rows"]
            xc = rng_vec(nc, 70 + l)
            assert dev.spmv(l, "P", xc).tobytes() == dev_h.spmv(l, "P", xc).tobytes()
            assert dev.spmv(l, "R", x).tobytes() == dev_h.spmv(l, "R", x).tobytes()
    n0 = A.nrows
    r1, x1, h1 = dev.solve(np.ones(n0), np.ones(n0))
    r2, x2, h2 = dev_h.solve(np.ones(n0), np.ones(n0))
    assert r1.nits == r2.nits and x1.tobytes() == x2.tobytes()
    dev_h.close()


GOLDEN_CASES = {"p3d64": ("p3d", 64, 0.0), "aniso3d64": ("aniso3d", 64, 1e-3), "v2732": ("v27", 32, 0.0), "p2d256": ("p2d", 256, 0.0),
                "p3d128": ("p3d", 128, 0.0)}     # p3d128 = BASELINE.json configs[1], the bench workload, at full size


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_solve_matches_reference_fixture(name):
    """full solves on the device against the committed fixtures generated from the REFERENCE ITSELF (tests/golden/golden.json,
    make_golden.py): same V-cycle count, ||r|| history within 1e-10 relative per iteration (the north-star's bar; observed ~1e-15,
    the norm is tree-reduced), solution bit-identical (sha256); plus size-independent properties at these sizes: solving twice
    gives the same bytes, and the returned residual is the true residual of the returned solution"""
    import hashlib
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))[name]
    kind, N, eps = GOLDEN_CASES[name]
    A = generate(kind, N, eps)
    assert (A.nrows, A.nnz) == (gold["n"], gold["nnz"])
    hier = HostHierarchy(A, tol=gold["tol"])
    dev = DeviceHierarchy(hier)
    n = A.nrows
    rtn, x, hist = dev.solve(np.ones(n), np.ones(n))
    want = np.array([float.fromhex(h) for h in gold["history_fix"]])
    assert rtn.nits == len(want)
    assert np.max(np.abs(hist - want) / want) <= RTOL_HISTORY
    assert hashlib.sha256(x.tobytes()).hexdigest() == gold["x_sha_fix"], "solution not bit-identical to the reference's"
    rtn2, x2, hist2 = dev.solve(np.ones(n), np.ones(n))
    assert x2.tobytes() == x.tobytes() and hist2.tobytes() == hist.tobytes()
    r, nrm = dev.residual(0, x, np.ones(n))
    assert abs(nrm - want[-1]) <= 1e-10 * want[-1]
    assert abs(np.linalg.norm(np.ones(n) - A.matvec(x)) - want[-1]) <= 1e-6 * want[-1]
    dev.close()
    hier.close()


# BASELINE.json configs[2], [3], [4] at their stated sizes: 3D 7-point 256^3, 27-point variable-coefficient
# 192^3, anisotropic (1,1,1e-3) 256^3.  Fixtures: tests/golden/make_golden.py add p3d 256 | add v27 192 | add aniso3d 256 1e-3, i.e. the
# reference's own SSS_amg_setup + SSS_amg_solve (/root/reference/amg/SSS_main.c:121-160 with tol 1e-8).
FULL_SIZE_CASES = {"p3d256": ("p3d", 256, 0.0, 20), "v27192": ("v27", 192, 0.0, 23), "aniso3d256": ("aniso3d", 256, 1e-3, 8)}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FULL_SIZE_CASES))
def test_baseline_config_full_size(name):
    """the three large BASELINE configs on the device against fixtures generated by the reference itself: same level table (rows/nnz
    per level of the hierarchy the host setup produces), same V-cycle count (20 / 23 / 8), ||r|| history within
    1e-10 relative per iteration, solution bit-identical (sha256), and the returned residual is the true residual of the solution"""
    import hashlib
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))[name]
    kind, N, eps, cycles = FULL_SIZE_CASES[name]
    A = generate(kind, N, eps)
    assert (A.nrows, A.nnz) == (gold["n"], gold["nnz"])
    hier = HostHierarchy(A, tol=gold["tol"])
    assert hier.table() == [(lv["A"]["rows"], lv["A"]["nnz"]) for lv in gold["levels"]]
    dev = DeviceHierarchy(hier)
    n = A.nrows
    rtn, x, hist = dev.solve(np.ones(n), np.ones(n))
    want = np.array([float.fromhex(h) for h in gold["history_fix"]])
    assert rtn.nits == len(want) == cycles
    assert np.max(np.abs(hist - want) / want) <= RTOL_HISTORY
    assert hashlib.sha256(x.tobytes()).hexdigest() == gold["x_sha_fix"], "solution not bit-identical to the reference's"
    r, nrm = dev.residual(0, x, np.ones(n))
    assert abs(nrm - want[-1]) <= 1e-10 * want[-1]
    dev.close()
    hier.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_1138_bus_matches_reference_fixture(mode):
    """the reference's only fixture matrix (Matrix/1138_bus.mtx: symmetric MatrixMarket file, irregular rows, reference default
    tol 1e-6) in both CG-beta modes against the committed results of the reference itself"""
    import hashlib
    import json
    import os
    from amg_b200 import read_mtx
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", "1138_bus.mtx")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/1138_bus.mtx not staged")
    gold = json.load(open(os.path.join(root, "tests", "golden", "golden.json")))["1138_bus_tol1e-6"]
    key = "fix" if mode == 0 else "asc"
    A = read_mtx(path)
    hier = HostHierarchy(A, tol=gold["tol"])
    dev = DeviceHierarchy(hier, coarse_mode=mode)
    n = A.nrows
    rtn, x, hist = dev.solve(np.ones(n), np.ones(n))
    want = np.array([float.fromhex(h) for h in gold[f"history_{key}"]])
    assert rtn.nits == len(want)
    assert np.max(np.abs(hist - want) / want) <= RTOL_HISTORY
    assert hashlib.sha256(x.tobytes()).hexdigest() == gold[f"x_sha_{key}"]
    dev.close()
    hier.close()
