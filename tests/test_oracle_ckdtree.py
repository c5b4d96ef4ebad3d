"""CPU tests: the cKDTree restatement (oracle/ckdtree_oracle.c) reproduces the scipy binary -- tree
permutation, and for k-NN queries the neighbour INDICES IN ORDER and the distances -- on the
tie-heavy inputs dense_lucaskanade produces (integer corners, half-integer medians, coincident
vectors, pixel-grid queries)."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle.ckdtree import KDTree


def _points(rng, n, W, kind):
    pts = np.floor(rng.uniform(0, W, (n, 2)))
    if kind == "half":
        pts = np.floor(rng.uniform(0, W, (n, 2)) * 2) / 2     # medians of integer coordinates
    elif kind == "dup":
        pts[: n // 4] = pts[n // 4: 2 * (n // 4)]              # coincident vectors
    elif kind == "real":
        pts = rng.uniform(0, W, (n, 2))
    elif kind == "few_values":
        pts = np.floor(rng.uniform(0, 4, (n, 2)))              # almost everything tied
    return pts


@pytest.mark.parametrize("kind", ["int", "half", "dup", "real", "few_values"])
@pytest.mark.parametrize("n", [1, 5, 17, 60, 400, 2000])
def test_tree_and_queries_match_scipy(kind, n):
    rng = np.random.default_rng(n * 7 + len(kind))
    W = int(rng.choice([16, 64, 300]))
    pts = _points(rng, n, W, kind)
    ref, mine = cKDTree(pts), KDTree(pts)
    assert np.array_equal(ref.indices, mine.indices)
    step = max(1, W // 25)
    gy, gx = np.meshgrid(np.arange(-2, W + 2, step), np.arange(-2, W + 2, step), indexing="ij")
    queries = np.concatenate([np.column_stack([gx.ravel(), gy.ravel()]).astype(float), pts[:300],
                              rng.uniform(-5, W + 5, (100, 2))])
    for k in (1, 2, 21, 31):
        dr, ir = ref.query(queries, k=k)
        dm, im = mine.query(queries, k=k)
        assert np.array_equal(ir, im), (kind, n, k)
        assert np.array_equal(dr, dm), (kind, n, k)


def test_single_query_and_missing_neighbours():
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    ref, mine = cKDTree(pts), KDTree(pts)
    for k in (1, 3, 5):
        dr, ir = ref.query([0.5, 0.5], k=k)
        dm, im = mine.query([0.5, 0.5], k=k)
        assert np.array_equal(np.atleast_1d(ir), np.atleast_1d(im))
        assert np.array_equal(np.atleast_1d(dr), np.atleast_1d(dm))


def test_nth_element_matches_libstdcxx(tmp_path):
    """std::nth_element itself (the tree build depends on the arrangement it leaves, not just on
    the selected element): compared with a C++ program compiled here, on inputs that also reach
    introselect's heap_select fallback."""
    import ctypes
    import shutil
    import subprocess
    from oracle import lib
    cxx = shutil.which("g++") or "/usr/bin/g++"
    src = tmp_path / "nth.cpp"
    src.write_text(r'''
#include <algorithm>
#include <cstdint>
extern "C" void ref_nth(const double *v, int64_t n, int64_t nth, int64_t *idx) {
    for (int64_t i = 0; i < n; i++) idx[i] = i;
    std::nth_element(idx, idx + nth, idx + n, [v](int64_t a, int64_t b) { return v[a] < v[b]; });
}
''')
    so = tmp_path / "libnth.so"
    subprocess.check_call([cxx, "-O2", "-shared", "-fPIC", "-o", str(so), str(src)])
    ref = ctypes.CDLL(str(so))
    L = lib()
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)
    rng = np.random.default_rng(3)
    cases = []
    for n in (1, 2, 3, 4, 5, 17, 100, 1000, 5000):
        cases += [rng.standard_normal(n), np.floor(rng.uniform(0, 4, n)), np.arange(n, dtype=float),
                  np.arange(n, dtype=float)[::-1].copy(), np.zeros(n),
                  np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]).astype(float)]
    # a median-of-3 killer sequence drives introselect into its heap_select fallback
    n = 4096
    killer = np.zeros(n)
    k = n // 2
    for i in range(1, k + 1):
        if i & 1:
            killer[i - 1] = i
            killer[i] = k + i
        killer[k + i - 1] = 2 * i
    cases.append(killer)
    for v in cases:
        v = np.ascontiguousarray(v, dtype=np.float64)
        for nth in sorted({0, len(v) // 2, len(v) - 1, len(v) // 3}):
            a = np.empty(len(v), dtype=np.int64)
            b = np.arange(len(v), dtype=np.int64)
            ref.ref_nth(v.ctypes.data_as(dp), len(v), nth, a.ctypes.data_as(ip))
            L.ora_kd_nth_element(v.ctypes.data_as(dp), ctypes.c_int64(len(v)), ctypes.c_int64(nth), -1,
                                 b.ctypes.data_as(ip))
            assert np.array_equal(a, b), (len(v), nth)
