"""Host hierarchy (amg_b200's from-scratch setup) vs the reference's SSS_amg_setup: every array of
every level must be byte-identical -- level count, per-level rows/nnz, C/F marks, CSR entry order
and values of A, P, R.  Checked against committed sha256 fixtures generated from the reference
itself (tests/golden/make_golden.py) and, where oracle/_ref exists, against it directly."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_ffi
from amg_b200 import HostHierarchy, capi, generate, read_mtx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
GEN = {"p2d64": ("p2d", 64, 0.0), "p2d256": ("p2d", 256, 0.0), "p3d16": ("p3d", 16, 0.0), "p3d32": ("p3d", 32, 0.0),
       "p3d64": ("p3d", 64, 0.0), "aniso3d32": ("aniso3d", 32, 1e-3), "aniso3d64": ("aniso3d", 64, 1e-3),
       "v2712": ("v27", 12, 0.0), "v2716": ("v27", 16, 0.0), "v2732": ("v27", 32, 0.0),
       "p3d128": ("p3d", 128, 0.0)}            # the bench size (BASELINE.json configs[1])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mat_hash(m):
    rp, ci, va = capi.mat_arrays(m)
    return {"rows": m.num_rows, "cols": m.num_cols, "nnz": m.num_nnzs, "row_ptr": sha(rp), "col_idx": sha(ci), "val": sha(va)}


@pytest.mark.parametrize("name", sorted(GEN))
def test_hierarchy_matches_reference_fixture(name):
    kind, N, eps = GEN[name]
    hier = HostHierarchy(generate(kind, N, eps), tol=1e-8)
    gold = GOLD[name]["levels"]
    assert hier.num_levels == len(gold)
    for l, g in enumerate(gold):
        c = hier.level(l)
        assert mat_hash(c.A) == g["A"], f"A level {l}"
        if l < hier.num_levels - 1:
            assert mat_hash(c.P) == g["P"], f"P level {l}"
            assert mat_hash(c.R) == g["R"], f"R level {l}"
            assert sha(hier.cfmark(l)) == g["cfmark"], f"cfmark level {l}"


# level tables measured by the survey with the reference's CPU path
APPENDIX_C = {
    ("p2d", 256, 0.0): ([65536, 32768, 8318, 2114, 542, 149, 39], [326656, 292866, 74586, 18544, 4628, 1343, 333]),
    ("p3d", 64, 0.0): ([262144, 131072, 23792, 5104, 1981, 944], [1810432, 2417024, 806862, 330918, 329407, 225862]),
    ("aniso3d", 64, 1e-3): ([262144, 131072, 34688, 9344, 2432, 580, 256, 35], [1810432, 3405180, 910860, 231040, 53960, 10170, 6170, 179]),
    ("v27", 32, 0.0): ([32768, 7974, 3436, 1580, 706], [830584, 542716, 386104, 263972, 139176]),
}


@pytest.mark.parametrize("key", sorted(APPENDIX_C))
def test_level_table_matches_survey(key):
    rows, nnz = APPENDIX_C[key]
    hier = HostHierarchy(generate(*key), tol=1e-8)
    assert [r for r, _ in hier.table()] == rows
    assert [z for _, z in hier.table()] == nnz


@pytest.mark.parametrize("case", [("p2d", 48, 0.0), ("p3d", 20, 0.0), ("aniso3d", 24, 1e-2), ("v27", 10, 0.0)])
def test_hierarchy_equals_reference_setup_directly(case, reference):
    A = generate(*case)
    hier = HostHierarchy(A, tol=1e-8)
    mg = reference.setup(A, capi.default_pars(1e-8))
    assert hier.num_levels == mg.num_levels
    for l in range(mg.num_levels):
        assert mat_hash(hier.level(l).A) == mat_hash(mg.cg[l].A)
        if l < mg.num_levels - 1:
            assert mat_hash(hier.level(l).P) == mat_hash(mg.cg[l].P)
            assert mat_hash(hier.level(l).R) == mat_hash(mg.cg[l].R)
            n = mg.cg[l].A.num_rows
            assert hier.cfmark(l).tobytes() == np.ctypeslib.as_array(mg.cg[l].cfmark.d, shape=(n,)).tobytes()
    reference.destroy(mg)


def test_reference_fixture_matrix_1138_bus():
    """the reference's only fixture: loader semantics (symmetric expansion in file order) + hierarchy"""
    path = os.path.join(ROOT, "oracle", "_ref", "1138_bus.mtx")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/1138_bus.mtx not staged")
    A = read_mtx(path)
    assert (A.nrows, A.nnz) == (1138, 4054)
    hier = HostHierarchy(A, tol=1e-6)
    assert hier.table() == [(1138, 4054), (511, 2465), (230, 1416), (111, 939), (59, 733)]
    gold = GOLD["1138_bus_tol1e-6"]["levels"]
    for l, g in enumerate(gold):
        assert mat_hash(hier.level(l).A) == g["A"]


def test_generators_shapes_and_symmetry():
    for kind, N, nnz in [("p2d", 16, 5 * 256 - 64), ("p3d", 8, 7 * 512 - 6 * 64), ("v27", 6, (3 * 6 - 2) ** 3)]:
        A = generate(kind, N)
        assert A.nnz == nnz
        # ascending columns inside each row, symmetric values
        d = np.diff(A.col_idx)
        row_starts = A.row_ptr[1:-1]
        mask = np.ones(len(d), bool); mask[row_starts - 1] = False
        assert (d[mask] > 0).all()
        import scipy.sparse as sp
        M = sp.csr_matrix((A.val, A.col_idx, A.row_ptr), shape=(A.nrows, A.nrows))
        assert abs(M - M.T).max() < 1e-14
