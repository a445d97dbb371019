"""amgb200_read_mtx (loader fast path for the kept C host) against the Python restatement of the reference's
loader semantics (amg_b200.host.read_mtx, which the 1138_bus fixtures pin to the reference itself): file order inside each row,
symmetric expansion entry by entry, pattern / integer / complex fields, comments, duplicates kept, ragged whitespace."""
import os

import numpy as np
import pytest

from amg_b200 import generate, read_mtx, read_mtx_fast

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def same(a, b):
    assert (a.nrows, a.ncols) == (b.nrows, b.ncols)
    assert a.row_ptr.tobytes() == b.row_ptr.tobytes() and a.col_idx.tobytes() == b.col_idx.tobytes() and a.val.tobytes() == b.val.tobytes()


def write(path, banner, shape, lines, comments=("% a comment", "%another")):
    with open(path, "w") as f:
        f.write(banner + "\n")
        for c in comments:
            f.write(c + "\n")
        f.write("%d %d %d\n" % shape)
        for ln in lines:
            f.write(ln + "\n")


def test_general_real_in_file_order_with_duplicates(tmp_path):
    rng = np.random.default_rng(3)
    m, n, nz = 37, 41, 400
    ri, ci = rng.integers(1, m + 1, nz), rng.integers(1, n + 1, nz)          # unsorted, with duplicates
    va = rng.standard_normal(nz) * 10.0 ** rng.integers(-8, 8, nz)
    p = str(tmp_path / "g.mtx")
    write(p, "%%MatrixMarket matrix coordinate real general", (m, n, nz), [f"{i} {j} {v:.17g}" for i, j, v in zip(ri, ci, va)])
    same(read_mtx_fast(p), read_mtx(p))


def test_symmetric_pattern_integer_complex(tmp_path):
    rng = np.random.default_rng(4)
    n, nz = 29, 120
    ri = rng.integers(1, n + 1, nz); ci = np.minimum(ri, rng.integers(1, n + 1, nz))      # lower triangle incl. diagonal
    p = str(tmp_path / "s.mtx")
    write(p, "%%MatrixMarket matrix coordinate real symmetric", (n, n, nz), [f"{i} {j} {float(v)!r}" for i, j, v in zip(ri, ci, rng.standard_normal(nz))])
    same(read_mtx_fast(p), read_mtx(p))
    write(p, "%%MatrixMarket matrix coordinate pattern symmetric", (n, n, nz), [f"{i}   {j}" for i, j in zip(ri, ci)])
    same(read_mtx_fast(p), read_mtx(p))
    write(p, "%%MatrixMarket MATRIX Coordinate Integer General", (n, n, nz), [f"{i}\t{j} {k}" for i, j, k in zip(ri, ci, rng.integers(-9, 9, nz))])
    same(read_mtx_fast(p), read_mtx(p))
    write(p, "%%MatrixMarket matrix coordinate complex general", (n, n, nz), [f"{i} {j} {v:.17g} {w:.17g}" for i, j, v, w in zip(ri, ci, rng.standard_normal(nz), rng.standard_normal(nz))])
    A = read_mtx_fast(p)
    assert A.nnz == nz                                                     # (imaginary parts dropped: mmio_highlevel.h:203-206)


def test_entries_not_one_per_line_and_errors(tmp_path):
    p = str(tmp_path / "w.mtx")
    write(p, "%%MatrixMarket matrix coordinate real general", (3, 3, 4), ["1 1 2.0 2 2", "3.0", "3 1 -1e-3", "", "   1 3 4"])
    A = read_mtx_fast(p)
    assert A.row_ptr.tolist() == [0, 2, 3, 4] and A.col_idx.tolist() == [0, 2, 1, 0] and A.val.tolist() == [2.0, 4.0, 3.0, -1e-3]
    write(p, "%%MatrixMarket matrix array real general", (3, 3, 4), [])
    with pytest.raises(ValueError):
        read_mtx_fast(p)
    write(p, "%%MatrixMarket matrix coordinate real general", (3, 3, 4), ["1 1 2.0"])
    with pytest.raises(ValueError):
        read_mtx_fast(p)
    with pytest.raises(ValueError):
        read_mtx_fast(str(tmp_path / "missing.mtx"))


def test_generated_operator_round_trip_and_cache(tmp_path, monkeypatch):
    A = generate("p3d", 12)
    p = str(tmp_path / "p3d12.mtx")
    rows = np.repeat(np.arange(A.nrows), np.diff(A.row_ptr))
    write(p, "%%MatrixMarket matrix coordinate real general", (A.nrows, A.ncols, A.nnz), [f"{r + 1} {c + 1} {v:.17g}" for r, c, v in zip(rows, A.col_idx, A.val)], comments=())
    same(read_mtx_fast(p), A)
    monkeypatch.setenv("AMGB200_MTX_CACHE", "1")
    same(read_mtx_fast(p), A)                                              # writes the cache
    assert os.path.exists(p + ".amgb200cache")
    same(read_mtx_fast(p), A)                                              # reads it


def test_reference_fixture_1138_bus():
    path = os.path.join(ROOT, "oracle", "_ref", "1138_bus.mtx")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/1138_bus.mtx not staged")
    A = read_mtx_fast(path)
    assert (A.nrows, A.nnz) == (1138, 4054)
    same(A, read_mtx(path))
