"""QPS reader (jrl-qp_b200/qps.py mirrors tests/QPSReader.cpp of the reference) and the reference's
Maros-Meszaros "Test Suite" loop (tests/GoldfarbIdnaniSolverTest.cpp:246-307) on the QPS files this repository carries:
status, testKKT and the published optimal objective value to 1e-6 — on the CPU oracle, and on the GPU through the
GoldfarbIdnaniSolver / experimental::GoldfarbIdnaniSolver mirrors."""
import math
import os

import numpy as np
import pytest

import pyoracle as po
import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import problems as P, qps

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qps")
INF = math.inf


def _read(name, full=True):
    return qps.QPSReader(full).read(os.path.join(DIR, name + ".QPS"))


def _batch(pb):
    """QPProblem -> the [1, ...] arrays of the C-ABI layout (C [mc, n]: row = constraint)."""
    return P.ProblemBatch(pb.G[None].copy(), pb.a[None].copy(), pb.C[None].copy(), pb.l[None].copy(), pb.u[None].copy(),
                          pb.xl[None].copy(), pb.xu[None].copy())


# ---------------------------------------------------------------------------------------------- reader
def test_reader_known_problem():
    pb, pr = _read("qptest", full=False)
    assert (pr.nbVar, pr.nbCstr, pr.nbEq, pr.useBounds, pr.hasFixedVariables) == (2, 2, 0, True, False)
    assert pb.name == "QPTEST" and pb.objCst == 0.0
    assert np.array_equal(pb.G, [[8, 0], [2, 10]])  # lower triangle only without fullObjMat
    assert np.array_equal(qps.QPSReader(True).read(os.path.join(DIR, "qptest.QPS"))[0].G, [[8, 2], [2, 10]])
    assert _read("hs21")[0].objCst == -100.0 and _read("hs35")[0].objCst == 9.0
    assert np.array_equal(pb.a, [1.5, -2]) and np.array_equal(pb.C, [[2, 1], [-1, 2]])
    assert np.array_equal(pb.l, [2, -INF]) and np.array_equal(pb.u, [INF, 6])
    assert np.array_equal(pb.xl, [0, 0]) and np.array_equal(pb.xu, [20, INF])


def test_reader_semantics_ranges_bounds_and_layout():
    pb, pr = qps.QPSReader(True).read(os.path.join(DIR, "reader_semantics.qps"))
    assert (pr.nbVar, pr.nbCstr, pr.nbEq, pr.useBounds, pr.hasFixedVariables) == (6, 5, 2, True, True)
    assert pb.name == "SEMANTICS" and pb.objCst == -2.5
    # rows in order: e_pos, e_neg, l_rng, g_rng, g_plain
    assert np.array_equal(pb.l, [1.0, 1.5, 1.0, 4.0, 0.0])  # E +0.5 -> u ; E -0.5 -> l ; L: u - |R| ; G ; G without rhs
    assert np.array_equal(pb.u, [1.5, 2.0, 3.0, 6.0, INF])
    assert np.array_equal(pb.a, [1, 0, -1, 0, 0, 0])
    C = np.zeros((5, 6))
    C[0, 0], C[1, 0], C[2, 0], C[3, 1], C[4, 1], C[0, 3], C[2, 4], C[3, 5] = 1, 2, 3, 4, 5, 1, 1, 1
    assert np.array_equal(pb.C, C)
    assert np.array_equal(pb.xl, [-1, 0, 3, -INF, -INF, 0]) and np.array_equal(pb.xu, [INF, 7, 3, INF, INF, INF])
    G = np.zeros((6, 6))
    G[0, 0], G[1, 0], G[0, 1], G[1, 1] = 2, 0.5, 0.5, 3
    assert np.array_equal(pb.G, G)


@pytest.mark.parametrize("text,msg", [
    ("NAME\nROWS\n N obj\n", "Failed to read name"),
    ("NAME x\nROWS\n N obj\n N obj2\n", "no restriction"),
    ("NAME x\nROWS\n N obj\n E r\n E r\n", "Duplicate row name"),
    ("NAME x\nROWS\n Q r\n", "Unknown row type"),
    ("NAME x\nROWS\n N obj\n E r\nCOLUMNS\n a r\n", "Failed to read first value"),
    ("NAME x\nROWS\n N obj\n E r\nCOLUMNS\n a r 1.0 obj\n", "Failed to read second value"),
    ("NAME x\nROWS\n N obj\n E r\nCOLUMNS\n a r 1.0\nRHS\n r1 r 1.0\n r2 r 2.0\n", "different RHS name"),
    ("NAME x\nROWS\n N obj\n E r\nCOLUMNS\n a r 1.0\nRANGES\n g obj 1.0\n", "range on a N row"),
    ("NAME x\nROWS\n N obj\n E r\nCOLUMNS\n a r 1.0\nBOUNDS\n BV b a 1.0\n", "Unknown bound type"),
    ("NAME x\n a b 1.0\n", "NAME section"),
    ("NAME x\nENDATA\n a b 1.0\n", "ENDATA section"),
])
def test_reader_errors_carry_the_context(tmp_path, text, msg):
    f = tmp_path / "bad.qps"
    f.write_text(text)
    with pytest.raises(qps.QPSError) as e:
        qps.QPSReader().read(str(f))
    assert msg in str(e.value) and "(line " in str(e.value) and "section " in str(e.value)
    with pytest.raises(qps.QPSError):
        qps.QPSReader().read(str(tmp_path / "missing.qps"))


def test_write_read_round_trip(tmp_path):
    # a random problem of the reference's generator, two-sided rows, an equality, every kind of bound
    ch = P.ProblemCharacteristics(6, nEq=2, nIneq=5, nStrongActIneq=2, bounds=True, nStrongActBounds=1, doubleSidedIneq=True)
    b = P.random_problems(ch, 1, seed=77)
    xl, xu = b.xl[0].copy(), b.xu[0].copy()
    xl[0], xu[0] = -INF, INF
    xl[1] = -INF
    xu[2] = INF
    xl[3] = xu[3] = 0.25
    lo, up = b.bl[0].copy(), b.bu[0].copy()
    lo[2] = -INF
    up[3] = INF
    src = qps.QPProblem(b.G[0], b.a[0], b.C[0], lo, up, xl, xu, objCst=-1.75)
    path = str(tmp_path / "rt.qps")
    qps.write_qps(path, src, "RT")
    got, pr = qps.QPSReader(True).read(path)
    assert pr.nbVar == 6 and pr.nbCstr == 7 and pr.nbEq == 2 and pr.useBounds and pr.hasFixedVariables
    for k in ("G", "a", "C", "xl", "xu"):
        assert np.array_equal(getattr(got, k), getattr(src, k)), k
    assert got.objCst == src.objCst
    assert np.array_equal(got.u, src.u)
    # two-sided rows go through RANGES: l = u - |u - l| (one rounding)
    fin = np.isfinite(src.l)
    assert np.array_equal(np.isinf(got.l), ~fin) and np.abs(got.l[fin] - src.l[fin]).max() <= 1e-15 * np.abs(src.u[fin]).max()


# ---------------------------------------------------------------------------------------------- the suite
def test_table_matches_the_files():
    for row in qps.marosMeszarosPbList:
        pb, pr = _read(row.name)
        assert (pr.nbVar, pr.nbCstr) == (row.nbVar, row.nbCstr), row.name
        ev = np.linalg.eigvalsh(pb.G)
        if row.cond == INF:
            assert ev.min() <= 1e-14 * ev.max()
        elif row.cond < 1e8:  # the table's estimate, to its printed digits
            assert abs(ev.max() / ev.min() - row.cond) <= 1e-4 * row.cond, row.name
    assert {r.name: qps.suite_action(r) for r in qps.marosMeszarosPbList} == {
        "hs21": "solve", "hs35": "solve", "hs35mod": "solve", "hs76": "solve", "qptest": "solve", "tame": "skip",
        "zecevic2": "non_pos_hessian"}


@pytest.mark.parametrize("row", qps.marosMeszarosPbList, ids=lambda r: r.name)
def test_suite_on_the_oracle(row):
    action = qps.suite_action(row)
    if action == "skip":
        pytest.skip("skipped by the reference's selection rules (cond)")
    pb, pr = _read(row.name)
    b = _batch(pb)
    r = po.solve_batch(b.G.copy(), b.a, b.C, b.bl, b.bu, b.xl, b.xu, max_iter=qps.suite_max_iter(row))
    if action == "non_pos_hessian":
        assert r["status"][0] == 2  # TerminationStatus::NON_POS_HESSIAN
        return
    assert r["status"][0] == 0
    assert P.test_kkt(r["x"], r["u"], b).all()
    assert r["f"][0] + pb.objCst == pytest.approx(row.fstar, rel=1e-6, abs=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("experimental", [False, True])
@pytest.mark.parametrize("row", qps.marosMeszarosPbList, ids=lambda r: r.name)
def test_suite_on_the_gpu_reads_like_the_reference(row, experimental):
    from jrl_qp_b200 import solver as S
    action = qps.suite_action(row)
    if action == "skip":
        pytest.skip("skipped by the reference's selection rules (cond)")
    pb, pr = _read(row.name)
    G = pb.G.copy()  # copy for the later check
    T = S.experimental.GoldfarbIdnaniSolver if experimental else S.GoldfarbIdnaniSolver
    qp = T(3, 5, False)  # sizes are not the correct ones: the resize must work
    qp.options(S.SolverOptions().maxIter(qps.suite_max_iter(row)))
    before = S.launch_count()
    ret = qp.solve(pb.G, pb.a, pb.C.T, pb.l, pb.u, pb.xl, pb.xu)
    assert S.launch_count() > before
    if action == "non_pos_hessian":
        assert ret == S.TerminationStatus.NON_POS_HESSIAN
        return
    assert ret == S.TerminationStatus.SUCCESS
    b = _batch(pb)
    b.G = G[None]
    assert P.test_kkt(qp.solution()[None], qp.multipliers()[None], b).all()
    assert qp.objectiveValue() + pb.objCst == pytest.approx(row.fstar, rel=1e-6, abs=1e-12)
    # and bit for bit what the oracle computes
    r = po.solve_batch(G[None].copy(), b.a, b.C, b.bl, b.bu, b.xl, b.xu, max_iter=qps.suite_max_iter(row))
    if not experimental:
        assert np.array_equal(qp.solution(), r["x"][0]) and qp.iterations() == r["iterations"][0]
