"""Host side of the tile-sparse numeric factor (csrc/cabi.cu: tile_symbolic; device side
csrc/ipm_tiles.cuh): the symbolic analysis is exercised on the CPU by replaying the numeric
factorisation in numpy from nothing but the tile structure and the update-pair lists it
produced -- if a tile of the fill or a pair were missing, L D L' would not reproduce M.
Reference algorithm: ldl.cl:381-502 (sparse_factor_primal_normal), sparse_ldl.py:72-152."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix, identity, hstack, rand

from pycllp_b200._cabi import tile_analysis
from pycllp_b200.problems import staircase_equality_arrays, sparse_equality_arrays


def replay_factor(A, d, ts, delta=1e-6):
    """The arithmetic of tiles_factor (ipm_tiles.cuh) in numpy, driven by the lists of `ts`."""
    m = A.shape[0]
    Ad = A.toarray()
    M = (Ad * d) @ Ad.T
    nbk, nt = ts["nbk"], ts["ntiles"]
    mp = 8 * nbk
    Mp = np.eye(mp)
    Mp[:m, :m] = M
    tid = {}
    for J in range(nbk):
        for t in range(ts["colptr"][J], ts["colptr"][J + 1]):
            assert ts["col"][t] == J
            tid[(int(ts["row"][t]), J)] = t
        assert ts["row"][ts["colptr"][J]] == J            # diagonal tile first
    T = np.zeros((nt, 8, 8))
    for i, j in zip(*np.nonzero(np.tril(Mp))):
        assert (i // 8, j // 8) in tid, "entry of A A' outside the tile pattern"
        T[tid[(i // 8, j // 8)], i % 8, j % 8] = Mp[i, j]
    beta = np.sqrt(np.abs(np.diag(M)).max())
    D = np.ones(mp)
    for J in range(nbk):
        c0, c1 = ts["colptr"][J], ts["colptr"][J + 1]
        for t in range(c0, c1):
            for p in range(ts["updptr"][t], ts["updptr"][t + 1]):
                a, b = ts["upda"][p], ts["updb"][p]
                K = ts["col"][a]
                assert ts["col"][b] == K and K < J
                assert ts["row"][a] == ts["row"][t] and ts["row"][b] == J
                T[t] -= T[a] @ np.diag(D[8 * K:8 * K + 8]) @ T[b].T
        col = T[c0:c1].reshape(-1, 8)                      # rows of the block column (view)
        for jj in range(8):
            w = col[:8, jj].copy()
            theta = np.abs(col[jj + 1:, jj]).max() if col.shape[0] > jj + 1 else 0.0
            Dj = max(abs(w[jj]), (theta / beta) ** 2, delta)
            D[8 * J + jj] = Dj
            col[jj, jj] = 1.0
            for R in range(jj + 1, col.shape[0]):
                l = col[R, jj] / Dj
                col[R, jj] = l
                jmax = R if R < 8 else 7
                for j2 in range(jj + 1, jmax + 1):
                    col[R, j2] -= l * w[j2]
    L = np.zeros((mp, mp))
    for (I, J), t in tid.items():
        blk = T[t] if I != J else np.tril(T[t])
        L[8 * I:8 * I + 8, 8 * J:8 * J + 8] = blk
    return Mp, L, D


CASES = [
    ("staircase", lambda: staircase_equality_arrays(96, 150, 24, 3, 1, seed=1)[0]),
    ("staircase_ragged_m", lambda: staircase_equality_arrays(77, 120, 16, 4, 1, seed=2)[0]),
    ("random_sparse", lambda: sparse_equality_arrays(60, 90, 0.03, 1, seed=3)[0]),
    ("arrow", lambda: hstack([csr_matrix(np.vstack([np.eye(40) * 2.0 + np.eye(40, k=1), np.ones((1, 40))])),
                              identity(41, format="csr")], format="csr")),
    ("dense_small", lambda: hstack([rand(20, 30, density=1.0, format="csr", random_state=5),
                                    identity(20, format="csr")], format="csr")),
]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_tile_structure_reproduces_the_factor(name, make):
    A = make().tocsr()
    ts = tile_analysis(A)
    rng = np.random.RandomState(7)
    d = 10.0 ** rng.uniform(-3, 3, A.shape[1])
    Mp, L, D = replay_factor(A, d, ts)
    R = L @ np.diag(D) @ L.T
    assert np.allclose(R, Mp, rtol=1e-9, atol=1e-9 * np.abs(Mp).max())
    assert ts["updptr"][-1] == ts["pairs"] == len(ts["upda"])


def test_tile_fill_of_a_staircase_is_small():
    A = staircase_equality_arrays(2000, 3000, 40, 4, 1)[0]
    ts = tile_analysis(A)
    full = ts["nbk"] * (ts["nbk"] + 1) // 2
    assert ts["ntiles"] < 0.05 * full                    # L on its pattern: < 5 % of the dense triangle
    # block rows ascend inside every block column
    for J in range(ts["nbk"]):
        r = ts["row"][ts["colptr"][J]:ts["colptr"][J + 1]]
        assert r[0] == J and np.all(np.diff(r) > 0)


def test_rcm_ordering_recovers_a_hidden_band():
    """A staircase LP whose constraints were shuffled: in the order given the factor fills in almost
    completely, the RCM ordering setup_sparse falls back on brings the tile count back to that of
    the unshuffled problem (within 2x), and the replayed factorisation of the reordered matrix is exact."""
    from pycllp_b200._cabi import rcm_ordering
    A0 = staircase_equality_arrays(480, 720, 24, 3, 1, seed=4)[0].tocsr()
    rng = np.random.RandomState(0)
    shuffle = rng.permutation(A0.shape[0])
    A = A0[shuffle]
    perm = rcm_ordering(A)
    assert sorted(perm.tolist()) == list(range(A.shape[0]))
    t_orig, t_shuf, t_rcm = (tile_analysis(X)["ntiles"] for X in (A0, A, A[perm]))
    assert t_shuf > 4 * t_orig
    assert t_rcm < 2 * t_orig
    Ar = A[perm][:96]                                    # replay a leading block (cheap)
    ts = tile_analysis(Ar)
    d = 10.0 ** rng.uniform(-2, 2, Ar.shape[1])
    Mp, L, D = replay_factor(Ar, d, ts)
    assert np.allclose(L @ np.diag(D) @ L.T, Mp, rtol=1e-9, atol=1e-9 * np.abs(Mp).max())
