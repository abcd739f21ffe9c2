"""CPU tests: the oracle (oracle/gi_oracle.cpp) against every known-answer test the reference holds
for the dense GoldfarbIdnaniSolver path. These pin the oracle; the GPU parity tests
(tests/test_gpu_parity.py) then compare the CUDA path with the oracle."""
import os

import numpy as np
import pytest

import pyoracle as po
import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import problems as P

I, L, U, E, LB, UB, FX = po.INACTIVE, po.LOWER, po.UPPER, po.EQUALITY, po.LOWER_BOUND, po.UPPER_BOUND, po.FIXED
INF = np.inf


# ---------------------------------------------------------------------------------------------
# tests/ActiveSetTest.cpp:60-133 — golden sequences, transcribed
# ---------------------------------------------------------------------------------------------
def _check(as_, status, active, nb):
    st, al, cnt = as_.query()
    assert st == status
    assert al == active
    assert cnt == list(nb)


def test_active_set_ctor():
    as_ = po.ActiveSet(5, 3)
    _check(as_, [I] * 8, [], (0, 0, 0, 0, 0, 0, 0, 0))


def test_active_set_activation_sequence():
    as_ = po.ActiveSet(5, 3)
    as_.activate(3, E)
    _check(as_, [I, I, I, E, I, I, I, I], [3], (1, 1, 0, 0, 0, 0, 0, 0))
    as_.activate(6, UB)
    _check(as_, [I, I, I, E, I, I, UB, I], [3, 6], (2, 1, 0, 0, 0, 1, 0, 1))
    as_.activate(2, L)
    _check(as_, [I, I, L, E, I, I, UB, I], [3, 6, 2], (3, 1, 1, 1, 0, 1, 0, 1))
    as_.activate(4, U)
    _check(as_, [I, I, L, E, U, I, UB, I], [3, 6, 2, 4], (4, 1, 2, 1, 1, 1, 0, 1))
    as_.deactivate(1)
    _check(as_, [I, I, L, E, U, I, I, I], [3, 2, 4], (3, 1, 2, 1, 1, 0, 0, 0))
    as_.activate(7, LB)
    _check(as_, [I, I, L, E, U, I, I, LB], [3, 2, 4, 7], (4, 1, 2, 1, 1, 1, 1, 0))
    as_.deactivate(2)
    _check(as_, [I, I, L, E, I, I, I, LB], [3, 2, 7], (3, 1, 1, 1, 0, 1, 1, 0))
    as_.deactivate(2)
    _check(as_, [I, I, L, E, I, I, I, I], [3, 2], (2, 1, 1, 1, 0, 0, 0, 0))
    as_.deactivate(0)
    _check(as_, [I, I, L, I, I, I, I, I], [2], (1, 0, 1, 1, 0, 0, 0, 0))
    as_.deactivate(0)
    _check(as_, [I] * 8, [], (0, 0, 0, 0, 0, 0, 0, 0))


# ---------------------------------------------------------------------------------------------
# canonical primitives
# ---------------------------------------------------------------------------------------------
def _fma(a, b, c):
    """Correctly rounded fused multiply-add via exact rational arithmetic."""
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def test_dot4_definition():
    rng = np.random.default_rng(1)
    for n in (0, 1, 3, 4, 5, 20, 50, 127):
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        acc = [0.0] * 4
        for k in range(n):
            acc[k & 3] = _fma(a[k], b[k], acc[k & 3])
        assert po.dot4(a, b) == (acc[0] + acc[1]) + (acc[2] + acc[3])


def test_dot32_definition():
    rng = np.random.default_rng(11)
    for n in (0, 1, 31, 32, 33, 50, 128, 387):
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        acc = [0.0] * 32
        for k in range(n):
            acc[k & 31] = _fma(a[k], b[k], acc[k & 31])
        off = 16
        while off >= 1:
            acc = [acc[l] + acc[l ^ off] for l in range(32)]
            off >>= 1
        assert po.dot32(a, b) == acc[0]


def test_dot_primitives_close_to_numpy():
    rng = np.random.default_rng(2)
    for n in (1, 7, 32, 50, 128, 387):
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        assert abs(po.dot4(a, b) - a @ b) <= 1e-12 * (1 + abs(a @ b))
        assert abs(po.dot32(a, b) - a @ b) <= 1e-12 * (1 + abs(a @ b))


def test_givens_conventions():
    # Eigen makeGivens: G = [c s; -s c], G^T [p; q] = [r; 0]  
    for p, q in ((3.0, 4.0), (-3.0, 4.0), (3.0, -4.0), (5.0, 1.0), (-5.0, -1.0), (0.0, 2.0), (0.0, -2.0), (2.0, 0.0),
                 (-2.0, 0.0), (0.0, 0.0)):
        c, s, r = po.givens(p, q)
        assert abs(c * p - s * q - r) <= 1e-15 * (1 + abs(r))
        assert abs(s * p + c * q) <= 1e-15 * (1 + abs(r))
        assert abs(c * c + s * s - 1) <= 1e-15
        assert r >= 0
    assert po.givens(-2.0, 0.0) == (-1.0, 0.0, 2.0)
    assert po.givens(0.0, -2.0) == (0.0, 1.0, 2.0)
    assert po.givens(0.0, 2.0) == (0.0, -1.0, 2.0)


# ---------------------------------------------------------------------------------------------
# tests/GoldfarbIdnaniSolverTest.cpp
# ---------------------------------------------------------------------------------------------
def _solve1(G, a, Cm, bl, bu, xl=None, xu=None, **kw):
    Cm = np.asarray(Cm, dtype=float)
    r = po.solve_batch(np.asarray(G, float)[None], np.asarray(a, float)[None], Cm[None], np.asarray(bl, float)[None],
                       np.asarray(bu, float)[None], None if xl is None else np.asarray(xl, float)[None],
                       None if xu is None else np.asarray(xu, float)[None], **kw)
    return r


def _pb1(G, a, Cm, bl, bu, xl=None, xu=None):
    f = lambda v: None if v is None else np.asarray(v, float)[None]
    return P.ProblemBatch(f(G), f(a), f(Cm), f(bl), f(bu), f(xl), f(xu))


def test_simple_problem():  # :23-49
    rng = np.random.default_rng(3)
    G = np.eye(3)
    a = np.zeros(3)
    Cm = rng.uniform(-1, 1, (5, 3))
    bl, bu = -np.ones(5), np.ones(5)
    r = _solve1(G, a, Cm, bl, bu)
    assert r["status"][0] == 0
    assert P.test_kkt(r["x"], r["u"], _pb1(G, a, Cm, bl, bu)).all()
    bl[1], bu[1] = -2, -1
    r = _solve1(G, a, Cm, bl, bu)
    assert r["status"][0] == 0 and r["iterations"][0] >= 1
    assert P.test_kkt(r["x"], r["u"], _pb1(G, a, Cm, bl, bu)).all()


def test_simple_problem_paper():  # :51-73, expected values: SURVEY.md §4.2 / §8c
    G = [[4.0, -2.0], [-2.0, 4.0]]
    r = _solve1(G, [6.0, 0.0], [[1.0, 1.0]], [2.0], [10.0], [0.0, 0.0], [10.0, 10.0])
    assert r["status"][0] == 0
    assert r["iterations"][0] == 1
    np.testing.assert_allclose(r["x"][0], [0.5, 1.5], rtol=0, atol=1e-15)
    np.testing.assert_allclose(r["u"][0], [-5.0, 0.0, 0.0], rtol=0, atol=1e-14)
    assert abs(r["f"][0] - 6.5) <= 1e-14
    assert r["active_set"][0].tolist() == [L, I, I]
    assert P.test_kkt(r["x"], r["u"], _pb1(G, [6.0, 0.0], [[1.0, 1.0]], [2.0], [10.0], [0.0, 0.0], [10.0, 10.0])).all()


REFERENCE_TEST_CHARACS = [  # tests/GoldfarbIdnaniSolverTest.cpp:77-82
    P.ProblemCharacteristics(5),
    P.ProblemCharacteristics(5, nEq=2),
    P.ProblemCharacteristics(5, nIneq=8, nStrongActIneq=4),
    P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=3),
    P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=1, bounds=True, nStrongActBounds=2),
]


@pytest.mark.parametrize("ch", REFERENCE_TEST_CHARACS)
def test_random_problems(ch):  # :75-99 (x200 seeds instead of one unseeded draw)
    pb = P.random_problems(ch, 200, seed=1234)
    r = po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    assert (r["status"] == 0).all()
    assert P.test_kkt(r["x"], r["u"], pb).all()
    assert P.is_approx(r["x"], pb.x, 1e-6).all()
    mc = pb.mc
    if ch.nEq:
        assert P.is_approx(r["u"][:, :ch.nEq], pb.lam[:, :ch.nEq], 1e-6).all()
    if ch.nIneq:
        assert P.is_approx(r["u"][:, ch.nEq:mc], pb.lam[:, ch.nEq:mc], 1e-6).all()
    if ch.bounds:
        assert P.is_approx(r["u"][:, mc:], pb.lam[:, mc:], 1e-6).all()


@pytest.mark.parametrize("cfg", ["config_A", "config_B", "config_D"])
def test_baseline_shapes_planted_solution(cfg):
    ch = getattr(P, cfg)()
    B = 16 if ch.nVar > 100 else 64
    pb = P.random_problems(ch, B)
    r = po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=4, instrument=True)
    assert (r["status"] == 0).all()
    assert P.test_kkt(r["x"], r["u"], pb).all()
    assert P.is_approx(r["x"], pb.x, 1e-6).all()
    assert (r["n_active"] == ch.nEq + ch.nStrongActIneq + ch.nStrongActBounds).all()
    # algorithmic flop counts of SURVEY.md §8(d) (0.81 / 0.053 / 12.9 MFLOP nominal)
    nominal = {"config_A": 0.81e6, "config_B": 0.053e6, "config_D": 12.9e6}[cfg]
    assert 0.8 * nominal < r["flops"].mean() < 1.2 * nominal


# ---------------------------------------------------------------------------------------------
# tests/BlockGISolverTest.in.cpp:172-188 / 273-284 — MultiIK fixtures with the dense solver
# ---------------------------------------------------------------------------------------------
def test_sequential_ik(golden_dir):
    d = np.load(os.path.join(golden_dir, "multiik_sequential.npz"))
    mc = d["u"].size
    r = po.solve_batch(d["G"][None], d["a"][None], d["C"][None], np.full((1, mc), -INF), d["u"][None])
    assert r["status"][0] == 0
    assert np.abs(r["x"][0] - d["sol"]).max() <= 1e-4  # the reference's own tolerance (:188)
    assert r["iterations"][0] == 4 and r["n_active"][0] == 4


def test_simultaneous_ik(golden_dir):
    d = np.load(os.path.join(golden_dir, "multiik_simultaneous.npz"))
    r = po.solve_batch(d["G"][None], d["a"][None], d["C"][None], np.full((1, 25), -INF), d["u"][None], d["xl"][None],
                       d["xu"][None])
    assert r["status"][0] == 0
    assert r["iterations"][0] == 5
    pb = P.ProblemBatch(d["G"][None], d["a"][None], d["C"][None], np.full((1, 25), -INF), d["u"][None], d["xl"][None],
                        d["xu"][None])
    assert P.test_kkt(r["x"], r["u"], pb).all()


# ---------------------------------------------------------------------------------------------
# independent cross-check and edge cases
# ---------------------------------------------------------------------------------------------
def test_solution_matches_kkt_system_on_final_active_set():
    """x from the oracle equals the equality-constrained minimiser on its reported active set."""
    pb = P.random_problems(P.config_B(), 32, seed=99)
    r = po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    n, mc = pb.n, pb.mc
    for b in range(pb.batch):
        rows, rhs = [], []
        for i, s in enumerate(r["active_set"][b]):
            if s == I:
                continue
            if i < mc:
                rows.append(pb.C[b, i])
                rhs.append(pb.bu[b, i] if s == U else pb.bl[b, i])
            else:
                e = np.zeros(n)
                e[i - mc] = 1
                rows.append(e)
                rhs.append(pb.xu[b, i - mc] if s == UB else pb.xl[b, i - mc])
        N = np.array(rows)
        q = len(rows)
        K = np.block([[pb.G[b], N.T], [N, np.zeros((q, q))]])
        sol = np.linalg.solve(K, np.concatenate([-pb.a[b], rhs]))
        np.testing.assert_allclose(r["x"][b], sol[:n], rtol=1e-8, atol=1e-9)


def test_non_positive_hessian():
    G = np.array([[1.0, 2.0], [2.0, 1.0]])
    r = _solve1(G, [1.0, 1.0], [[1.0, 0.0]], [-1.0], [1.0])
    assert r["status"][0] == 2


def test_infeasible():
    # x0 >= 1 and x0 <= -1 through two constraints
    r = _solve1(np.eye(2), [0.0, 0.0], [[1.0, 0.0], [1.0, 0.0]], [1.0, -INF], [INF, -1.0])
    assert r["status"][0] == 3


def test_max_iter():
    pb = P.random_problems(P.config_B(), 4)
    r = po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, max_iter=3)
    assert (r["status"] == 4).all() and (r["iterations"] == 3).all()


def test_unconstrained_and_empty_constraint_set():
    rng = np.random.default_rng(5)
    A = rng.standard_normal((6, 6))
    G = A.T @ A + np.eye(6)
    a = rng.standard_normal(6)
    r = po.solve_batch(G[None], a[None], np.zeros((1, 0, 6)), np.zeros((1, 0)), np.zeros((1, 0)))
    assert r["status"][0] == 0 and r["iterations"][0] == 0
    np.testing.assert_allclose(r["x"][0], -np.linalg.solve(G, a), rtol=1e-10)


def test_activation_status_quirk_is_reproduced():
    """With equality columns NOT first, the reference's activationStatus(k) quirk
    changes the trajectory but still terminates with a KKT point; the oracle must reproduce it."""
    ch = P.config_B()
    pb = P.random_problems(ch, 64, seed=7)
    base = po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    perm = np.arange(pb.mc)[::-1].copy()  # equalities last
    r = po.solve_batch(pb.G, pb.a, pb.C[:, perm], pb.bl[:, perm], pb.bu[:, perm], pb.xl, pb.xu)
    assert (r["status"] == 0).all()
    pbp = P.ProblemBatch(pb.G, pb.a, pb.C[:, perm], pb.bl[:, perm], pb.bu[:, perm], pb.xl, pb.xu)
    assert P.test_kkt(r["x"], r["u"], pbp).all()
    assert P.is_approx(r["x"], base["x"], 1e-6).all()
    assert (r["iterations"] != base["iterations"]).any()  # the quirk is visible in the iteration counts
