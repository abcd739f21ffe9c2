"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI
(include/jrlqp_b200.h), against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): final active set, status code and iteration count identical; x, u
and f within a relative error of 1e-9. The kernels reproduce the oracle's canonical operation order, so
the tests additionally assert bit-for-bit equality."""
import os

import numpy as np
import pytest
import torch

import pyoracle as po
import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import problems as P, solver as S

pytestmark = pytest.mark.gpu
RTOL = 1e-9  # north_star tolerance for x, multipliers and objective
INF = np.inf


def _oracle(pb, **kw):
    return po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count(), **kw)


def _gpu(pb, stage=-1, max_iter=None, big_bnd=None, want_L=False, scan_transposed=None):
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, pb.xl is not None, pb.batch)
    sv.set_stage_c(stage)
    if scan_transposed is not None:
        sv.set_scan_transposed(scan_transposed)
    if max_iter is not None or big_bnd is not None:
        o = S.SolverOptions()
        if max_iter is not None:
            o.maxIter(max_iter)
        if big_bnd is not None:
            o.bigBnd(big_bnd)
        sv.options(o)
    before = S.launch_count()
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, want_L=want_L)
    assert S.launch_count() > before, "no CUDA kernel was launched"
    return sv.last


def assert_parity(g, ref, bit_exact=True):
    for k in ("status", "iterations", "n_active", "active_set", "active_list"):
        assert np.array_equal(g[k], ref[k]), f"{k} differs from the oracle"
    ok = ref["status"] != 2  # x/u/f unspecified on NON_POS_HESSIAN
    for k in ("x", "u", "f"):
        a, b = g[k][ok], ref[k][ok]
        if a.size == 0:
            continue
        scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True) if b.ndim > 1 else np.abs(b), 1e-300)
        assert (np.abs(a - b) <= RTOL * np.maximum(scale, 1.0)).all(), f"{k} outside rtol {RTOL}"
        if bit_exact:
            assert np.array_equal(a, b), f"{k} not bit-identical to the oracle"


REFERENCE_TEST_CHARACS = [  # tests/GoldfarbIdnaniSolverTest.cpp:77-82
    P.ProblemCharacteristics(5),
    P.ProblemCharacteristics(5, nEq=2),
    P.ProblemCharacteristics(5, nIneq=8, nStrongActIneq=4),
    P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=3),
    P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=1, bounds=True, nStrongActBounds=2),
]


@pytest.mark.parametrize("ch", REFERENCE_TEST_CHARACS)
def test_reference_random_problem_sets(ch):
    pb = P.random_problems(ch, 256, seed=1234)
    g = _gpu(pb)
    assert_parity(g, _oracle(pb))
    assert (g["status"] == 0).all()
    assert P.test_kkt(g["x"], g["u"], pb).all()
    assert P.is_approx(g["x"], pb.x, 1e-6).all()
    assert P.is_approx(g["u"], pb.lam, 1e-6).all() or not (ch.nEq + ch.nIneq)


@pytest.mark.parametrize("cfg,B", [("config_B", 4096), ("config_A", 2048), ("config_D", 296)])
@pytest.mark.parametrize("stage", [0, 1])
def test_baseline_configs(cfg, B, stage):
    ch = getattr(P, cfg)()
    if stage == 1 and ch.nVar > 100:
        pytest.skip("C does not fit in shared memory next to J and R at n=128")
    pb = P.random_problems(ch, B)
    g = _gpu(pb, stage)
    assert_parity(g, _oracle(pb))
    assert (g["status"] == 0).all()
    assert P.test_kkt(g["x"], g["u"], pb).all()


@pytest.mark.parametrize("ch,B", [(P.config_A(), 3000), (P.config_D(), 300), (P.config_B(), 5000),
                                  (P.ProblemCharacteristics(70, 5, 150, nStrongActIneq=20, bounds=True, nStrongActBounds=5,
                                                            doubleSidedIneq=True), 500),
                                  (P.ProblemCharacteristics(33, 3, 7, nStrongActIneq=2), 700)])
def test_constraint_scan_transposed_copy_and_in_place_agree(ch, B):
    """C not staged: the scan over the CTA's transposed copy of C (default) and the in-place scan are the same
    arithmetic; both bit-identical to the oracle (mc > n, mc < n, every CTA width, more problems than slices)."""
    pb = P.random_problems(ch, B, seed=4242)
    ref = _oracle(pb)
    gt = _gpu(pb, stage=0, scan_transposed=True)
    gi = _gpu(pb, stage=0, scan_transposed=False)
    assert_parity(gt, ref)
    assert_parity(gi, ref)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 63, 64, 65, 96, 97, 127])
def test_sizes_across_slot_boundaries(n):
    ne = n // 5
    ni = n - ne
    ch = P.ProblemCharacteristics(n, ne, ni, min(ni, max(0, n // 4)), 0, min(n // 10, n - ne - min(ni, n // 4)), 0, True, True)
    pb = P.random_problems(ch, 96, seed=n)
    g = _gpu(pb)
    assert_parity(g, _oracle(pb))
    assert (g["status"] == 0).all()


def test_degenerate_weakly_active_constraints():
    ch = P.ProblemCharacteristics(12, 2, 14, 3, 3, 1, 2, True, True)
    pb = P.random_problems(ch, 512, seed=3)
    g = _gpu(pb)
    assert_parity(g, _oracle(pb))


def test_single_sided_no_bounds_and_bounds_only():
    for ch in (P.ProblemCharacteristics(20, 0, 30, 8, 0, 0, 0, False, False),
               P.ProblemCharacteristics(20, 0, 0, 0, 0, 6, 0, True, False),
               P.ProblemCharacteristics(20, 20, 0, 0, 0, 0, 0, False, False),   # as many equalities as variables
               P.ProblemCharacteristics(20, 0, 0, 0, 0, 0, 0, False, False)):  # unconstrained
        pb = P.random_problems(ch, 128, seed=11)
        g = _gpu(pb)
        assert_parity(g, _oracle(pb))
        assert (g["status"] == 0).all()


def test_equalities_not_first_reproduces_reference_quirk():
    pb = P.random_problems(P.config_B(), 512, seed=7)
    perm = np.arange(pb.mc)[::-1].copy()
    pbp = P.ProblemBatch(pb.G, pb.a, np.ascontiguousarray(pb.C[:, perm]), np.ascontiguousarray(pb.bl[:, perm]),
                         np.ascontiguousarray(pb.bu[:, perm]), pb.xl, pb.xu)
    g = _gpu(pbp)
    ref = _oracle(pbp)
    assert_parity(g, ref)
    assert (ref["iterations"] != _oracle(pb)["iterations"]).any()


def test_fixed_variables_and_equalities_mixed():
    pb = P.random_problems(P.config_B(), 256, seed=21)
    pb.xl[:, 3] = pb.xu[:, 3] = pb.x[:, 3]  # xl == xu -> FIXED (src/GoldfarbIdnaniSolver.cpp:279-286)
    pb.xl[:, 11] = pb.xu[:, 11] = pb.x[:, 11]
    g = _gpu(pb)
    ref = _oracle(pb)
    assert_parity(g, ref)
    assert (g["status"] == 0).all()
    # Because of the reference's activationStatus(k) indexing quirk a FIXED variable can be
    # dropped and re-activated as a plain bound; the kernel follows the reference, so only parity is asserted.
    assert np.isin(g["active_set"][:, pb.mc + 3], [0, 4, 5, 6]).all()
    assert (g["active_set"][:, pb.mc + 3] == S.ActivationStatus.FIXED).any()


def test_failure_statuses_in_a_mixed_batch():
    pb = P.random_problems(P.config_B(), 64, seed=5)
    pb.G[3] = -pb.G[3]  # not positive definite
    pb.G[10, 0, 0] = 0.0
    pb.G[10, 0, 1:] = 0.0
    pb.G[10, 1:, 0] = 0.0  # zero pivot
    pb.C[30, 6] = pb.C[30, 5]
    pb.bl[30, 5], pb.bu[30, 5] = 1.0, 2.0
    pb.bl[30, 6], pb.bu[30, 6] = -2.0, -1.0  # contradictory pair -> INFEASIBLE
    g = _gpu(pb)
    ref = _oracle(pb)
    assert_parity(g, ref)
    assert g["status"][3] == S.TerminationStatus.NON_POS_HESSIAN
    assert g["status"][10] == S.TerminationStatus.NON_POS_HESSIAN
    assert g["status"][30] == S.TerminationStatus.INFEASIBLE
    assert (np.delete(g["status"], [3, 10, 30]) == 0).all()


def test_inconsistent_bounds_take_the_exact_sequential_scan():
    """bl > bu: both slacks negative, the one case where a parallel first-min differs from the
    reference's else-if chain (src/GoldfarbIdnaniSolver.cpp:98-109); the kernel falls back to an exact
    sequential scan, so even these ill-posed inputs follow the reference trajectory."""
    pb = P.random_problems(P.config_B(), 128, seed=13)
    pb.bl[:, 7], pb.bu[:, 7] = pb.bu[:, 7] + 0.5, pb.bl[:, 7] - 0.5
    pb.xl[:, 2], pb.xu[:, 2] = pb.xu[:, 2] + 0.1, pb.xl[:, 2] - 0.1
    g = _gpu(pb, max_iter=60)
    assert_parity(g, _oracle(pb, max_iter=60))


def test_max_iter_and_big_bnd_options():
    pb = P.random_problems(P.config_B(), 64, seed=9)
    g = _gpu(pb, max_iter=3)
    assert_parity(g, _oracle(pb, max_iter=3))
    assert (g["status"] == S.TerminationStatus.MAX_ITER_REACHED).all() and (g["iterations"] == 3).all()
    g = _gpu(pb, big_bnd=1e30)
    assert_parity(g, _oracle(pb, big_bnd=1e30))


def test_batch_of_one_and_empty_batch():
    pb = P.random_problems(P.config_A(), 1, seed=2)
    assert_parity(_gpu(pb), _oracle(pb))
    sv = S.BatchedGoldfarbIdnaniSolver(5, 3, True, 4)
    pb0 = P.random_problems(P.ProblemCharacteristics(5, 1, 2, 1, 0, 0, 0, True, True), 1)
    assert sv.solve(pb0.G[:0], pb0.a[:0], pb0.C[:0], pb0.bl[:0], pb0.bu[:0], pb0.xl[:0], pb0.xu[:0]) == 0


@pytest.mark.parametrize("ch,B", [(P.config_A(), 1500), (P.config_B(), 3000), (P.config_D(), 200)])
def test_host_entry_moves_only_the_lower_triangle_of_G(ch, B):
    """jrlqp_solve_batch_host: with a pinned caller buffer and n <= 64 the kernels read G in place over the host link;
    otherwise G is uploaded as its left half plus the bottom-right block. The upper triangle is never read (as in the
    reference, src/GoldfarbIdnaniSolver.cpp:58): poisoning it changes nothing."""
    pb = P.random_problems(ch, B, seed=99)
    ref = _oracle(pb)
    n = pb.n
    iu = np.triu_indices(n, 1)
    Gp = torch.empty(pb.G.shape, dtype=torch.float64).pin_memory()
    Gp.numpy()[:] = pb.G
    Gp.numpy()[:, iu[1], iu[0]] = np.nan  # G [B][col][row]: entry (row < col) is the strict upper triangle
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, B)
    two_blocks = 8 * (n * (n // 2) + (n - n // 2) ** 2) if n >= 32 else 8 * n * n
    assert sv.host_g_bytes(True) == (8 * n * (n + 1) // 2 if n <= 64 else two_blocks)
    assert sv.host_g_bytes(False) == two_blocks
    sv.solve(Gp.numpy(), pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)  # pinned
    assert_parity(sv.last, ref)
    Gn = Gp.numpy().copy()  # pageable
    sv.solve(Gn, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    assert_parity(sv.last, ref)
    if n <= 64:  # the warm-start entry point through the same in-place read of G
        sv.options(S.SolverOptions().warmStart(True))
        sv.solve(Gp.numpy(), pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, experimental=True, as_in=ref["active_set"])
        assert (sv.last["iterations"] == 0).all() and np.array_equal(sv.last["active_set"], ref["active_set"])
        assert np.allclose(sv.last["x"], ref["x"], rtol=1e-6, atol=1e-8)


def test_shared_hessian_and_constraints_stride_zero():
    pb = P.random_problems(P.config_B(), 300, seed=31)
    rng = np.random.default_rng(0)
    a = pb.a[0] + 0.05 * rng.standard_normal(pb.a.shape)
    sh = P.ProblemBatch(pb.G[0], a, pb.C[0], pb.bl[0], pb.bu[0], pb.xl[0], pb.xu[0])
    ref = po.solve_batch(sh.G, sh.a, sh.C, sh.bl, sh.bu, sh.xl, sh.xu)
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, 300)
    sv.solve(sh.G, sh.a, sh.C, sh.bl, sh.bu, sh.xl, sh.xu)
    assert_parity(sv.last, ref)
    sv.set_stage_c(0)  # C read from global memory: the shared C is transposed once per CTA and re-used
    sv.solve(sh.G, sh.a, sh.C, sh.bl, sh.bu, sh.xl, sh.xu)
    assert_parity(sv.last, ref)


def test_cholesky_factor_is_returned_like_the_reference_leaves_it_in_G():
    pb = P.random_problems(P.config_A(), 64, seed=4)
    g = _gpu(pb, want_L=True)
    ref = _oracle(pb, want_L=True)
    il = np.tril_indices(pb.n)
    Lg = g["L"].transpose(0, 2, 1)[:, il[0], il[1]]  # column-major blocks -> (i, j)
    Lr = ref["L"].transpose(0, 2, 1)[:, il[0], il[1]]
    assert np.array_equal(Lg, Lr)
    L = np.tril(g["L"][0].T)
    np.testing.assert_allclose(L @ L.T, pb.G[0], rtol=1e-12, atol=1e-12)


def test_device_pointer_entry_point_with_leading_dimensions_and_stream():
    ch = P.config_B()
    pb = P.random_problems(ch, 777, seed=17)
    n, mc, B = pb.n, pb.mc, pb.batch
    ldg, ldc = n + 3, n + 5
    dev = torch.device("cuda:0")
    Gp = torch.full((B, n, ldg), float("nan"), dtype=torch.float64, device=dev)
    Gp[:, :, :n] = torch.from_numpy(pb.G).to(dev)
    Cp = torch.full((B, mc, ldc), float("nan"), dtype=torch.float64, device=dev)
    Cp[:, :, :n] = torch.from_numpy(pb.C).to(dev)
    t = lambda v: torch.from_numpy(v).to(dev)
    a, bl, bu, xl, xu = map(t, (pb.a, pb.bl, pb.bu, pb.xl, pb.xu))
    x = torch.empty((B, n), dtype=torch.float64, device=dev)
    u = torch.empty((B, mc + n), dtype=torch.float64, device=dev)
    f = torch.empty(B, dtype=torch.float64, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    act = torch.empty((B, mc + n), dtype=torch.int8, device=dev)
    al = torch.empty((B, n), dtype=torch.int32, device=dev)
    na = torch.empty(B, dtype=torch.int32, device=dev)
    sv = S.BatchedGoldfarbIdnaniSolver(n, mc, True, 1)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        sv.solve_device(B, Gp, a, Cp, bl, bu, xl, xu, x, u, f, it, st, act, al, na, stream=stream.cuda_stream,
                        ldg=ldg, ldc=ldc, strides={"G": n * ldg, "C": mc * ldc})
    stream.synchronize()
    g = dict(x=x.cpu().numpy(), u=u.cpu().numpy(), f=f.cpu().numpy(), iterations=it.cpu().numpy(),
             status=st.cpu().numpy(), active_set=act.cpu().numpy(), active_list=al.cpu().numpy(), n_active=na.cpu().numpy())
    assert_parity(g, _oracle(pb))


# ---- the reference's own solver tests, through the class mirror (tests/GoldfarbIdnaniSolverTest.cpp) ----
def test_simple_problem_paper_through_class_mirror():  # :51-73
    G = np.array([[4.0, -2.0], [-2.0, 4.0]])
    a = np.array([6.0, 0.0])
    Cmat = np.array([[1.0], [1.0]])  # n x nbCstr, one constraint per column
    qp = S.GoldfarbIdnaniSolver(2, 1, True)
    ret = qp.solve(G, a, Cmat, np.array([2.0]), np.array([10.0]), np.zeros(2), np.full(2, 10.0))
    assert ret == S.TerminationStatus.SUCCESS
    np.testing.assert_allclose(qp.solution(), [0.5, 1.5], atol=1e-15)
    np.testing.assert_allclose(qp.multipliers(), [-5.0, 0.0, 0.0], atol=1e-14)
    assert abs(qp.objectiveValue() - 6.5) < 1e-14 and qp.iterations() == 1
    assert qp.activeSet() == [S.ActivationStatus.LOWER, S.ActivationStatus.INACTIVE, S.ActivationStatus.INACTIVE]
    # G is an in/out argument: its lower triangle now holds the Cholesky factor
    np.testing.assert_allclose(np.tril(G), np.linalg.cholesky(np.array([[4.0, -2.0], [-2.0, 4.0]])), atol=1e-15)


def test_simple_problem_and_multiple_uses():  # :23-49 and :101-125 (one solver object re-used)
    rng = np.random.default_rng(3)
    qp = S.GoldfarbIdnaniSolver(3, 5, False)
    Cmat = rng.uniform(-1, 1, (5, 3)).T.copy()  # n x nbCstr
    bl, bu = -np.ones(5), np.ones(5)
    assert qp.solve(np.eye(3), np.zeros(3), Cmat, bl, bu, None, None) == 0
    assert np.abs(qp.solution()).max() == 0 and qp.iterations() == 0
    bl[1], bu[1] = -2, -1
    assert qp.solve(np.eye(3), np.zeros(3), Cmat, bl, bu, None, None) == 0
    pb = P.ProblemBatch(np.eye(3)[None], np.zeros((1, 3)), Cmat.T[None].copy(), bl[None], bu[None], None, None)
    assert P.test_kkt(qp.solution()[None], qp.multipliers()[None], pb).all()
    solver = S.GoldfarbIdnaniSolver(5, 8, True)
    for ch in REFERENCE_TEST_CHARACS:
        p1 = P.random_problems(ch, 1, seed=8)
        ret = solver.solve(p1.G[0].copy(), p1.a[0], p1.C[0].T, p1.bl[0], p1.bu[0],
                           None if p1.xl is None else p1.xl[0], None if p1.xu is None else p1.xu[0])
        assert ret == S.TerminationStatus.SUCCESS
        assert P.test_kkt(solver.solution()[None], solver.multipliers()[None], p1).all()
        assert P.is_approx(solver.solution(), p1.x[0], 1e-6)


# ---- BASELINE.json full size: size-independent properties + sampled oracle comparison ----
def test_headline_config_full_size_properties():
    ch = P.config_A()
    B = 65536
    pb = P.random_problems(ch, B, seed=P.DEFAULT_SEED)
    g = _gpu(pb)
    assert (g["status"] == 0).all()
    assert (g["n_active"] == ch.nEq + ch.nStrongActIneq + ch.nStrongActBounds).all()
    assert P.is_approx(g["x"], pb.x, 1e-6).all()          # planted solution (tests/GoldfarbIdnaniSolverTest.cpp:94)
    for lo in range(0, B, 8192):
        sl = pb.slice(lo, lo + 8192)
        assert P.test_kkt(g["x"][lo:lo + 8192], g["u"][lo:lo + 8192], sl).all()
    idx = np.arange(0, B, 32)
    sub = P.ProblemBatch(pb.G[idx], pb.a[idx], pb.C[idx], pb.bl[idx], pb.bu[idx], pb.xl[idx], pb.xu[idx])
    ref = _oracle(sub)
    assert_parity({k: v[idx] for k, v in g.items() if isinstance(v, np.ndarray)}, ref)


@pytest.mark.parametrize("span,ulps", [(0, 0), (30, 3), (300, 3), (30, 200), (600, 1)])
def test_exact_arithmetic_primitives(span, ulps):
    """fp64_exact.cuh: a quotient that div_rcp declares proven IS x / y bit for bit, whatever reciprocal
    approximation it was given; the restated square-root sequence IS sqrt() bit for bit on [1, 2]."""
    proven, wrong, unproven, sqrt_bad, rsqrt_far = S.selftest_arith(1 << 26, seed=span * 1000 + ulps, exponent_span=span, rcp_ulps=ulps)
    assert wrong == 0 and sqrt_bad == 0 and rsqrt_far == 0
    total = proven + unproven
    if span <= 30 and ulps <= 3:  # beyond +-400 binades of range the proof is declined by design
        assert unproven <= 1e-3 * total  # the stock-division fallback is rare


# ---------------------------------------------------------------------------------------------
# warm start: experimental::GoldfarbIdnaniSolver
# ---------------------------------------------------------------------------------------------
def _gpu_warm(pb, as_in, warm=True, max_iter=None):
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, pb.xl is not None, pb.batch)
    o = S.SolverOptions().warmStart(warm)
    if max_iter is not None:
        o.maxIter(max_iter)
    sv.options(o)
    before = S.launch_count()
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, experimental=True, as_in=as_in)
    assert S.launch_count() > before
    return sv.last


def _oracle_warm(pb, as_in, warm=True, max_iter=500):
    return po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count(), experimental=True,
                          warm_start=warm, as_in=as_in, max_iter=max_iter)


WARM_CHARACS = [  # tests/GoldfarbIdnaniSolverTest.cpp:139-143 (the reference's warm-start test) + the bench shapes
    P.ProblemCharacteristics(5), P.ProblemCharacteristics(5, nEq=2),
    P.ProblemCharacteristics(5, nIneq=8, nStrongActIneq=4), P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=3),
    P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=1, bounds=True, nStrongActBounds=2),
]


WIDE = P.ProblemCharacteristics(70, 5, 150, nStrongActIneq=20, bounds=True, nStrongActBounds=5, doubleSidedIneq=True)  # W = 3, mc > 128


@pytest.mark.parametrize("ch", WARM_CHARACS + [P.config_B(), P.config_A(), WIDE], ids=lambda c: f"n{c.nVar}e{c.nEq}i{c.nIneq}b{int(c.bounds)}")
def test_warm_start_exact_guess_takes_zero_iterations(ch):
    """tests/GoldfarbIdnaniSolverTest.cpp:127-181: warm start from the active set of the cold solve gives
    SUCCESS, iterations() == 0, KKT, the planted solution; and the CUDA path equals the oracle bit for bit."""
    B = 2000 if ch.nVar <= 20 else 512
    pb = P.random_problems(ch, B, seed=77)
    cold = _gpu(pb)
    assert (cold["status"] == 0).all()
    g = _gpu_warm(pb, cold["active_set"])
    ref = _oracle_warm(pb, cold["active_set"])
    assert_parity(g, ref)
    assert (g["status"] == 0).all()
    assert (g["iterations"] == 0).all()
    assert np.array_equal(g["active_set"], cold["active_set"])
    assert P.test_kkt(g["x"], g["u"], pb).all()
    assert P.is_approx(g["x"], pb.x, 1e-6).mean() > 0.999  # the reference tolerates 0.1 % of failures here
    for k in ("x", "u", "f"):
        assert np.allclose(g[k], cold[k], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("ch", WARM_CHARACS[2:] + [P.config_B(), WIDE], ids=lambda c: f"n{c.nVar}e{c.nEq}i{c.nIneq}b{int(c.bounds)}")
def test_warm_start_wrong_guesses(ch):
    """rubbish / partially wrong guesses (tests/GoldfarbIdnaniSolverTest.cpp:183-216): still SUCCESS + KKT,
    same trajectory as the oracle (drops of negative multipliers counted in iterations())."""
    pb = P.random_problems(ch, 1000, seed=5)
    cold = _gpu(pb)
    rng = np.random.default_rng(3)
    m = cold["active_set"].shape[1]
    guess = cold["active_set"].copy()
    flip = rng.random(guess.shape) < 0.25
    # swap sides, activate inactive ones, deactivate active ones, sprinkle invalid statuses
    swapped = guess.copy()
    swapped[guess == 1], swapped[guess == 2], swapped[guess == 4], swapped[guess == 5] = 2, 1, 5, 4
    kind_lower = np.where(np.arange(m) < pb.mc, 1, 4).astype(np.int8)
    swapped[guess == 0] = np.broadcast_to(kind_lower, guess.shape)[guess == 0]
    guess = np.where(flip, swapped, guess)
    guess[rng.random(guess.shape) < 0.02] = 6  # FIXED guesses are ignored
    guess[rng.random(guess.shape) < 0.02] = 4  # bound status on general constraints: ignored there
    g = _gpu_warm(pb, guess)
    ref = _oracle_warm(pb, guess)
    assert_parity(g, ref)
    ok = g["status"] == 0
    assert ok.mean() > 0.9  # the reference itself notes that rubbish guesses can fail (its test disables that part)
    # the algorithm itself (reference included, see the FIXME at tests/GoldfarbIdnaniSolverTest.cpp:186) ends on a
    # non-optimal point for a few percent of such guesses; what is checked exactly is the parity above
    assert P.test_kkt(g["x"], g["u"], pb)[ok].mean() > (0.95 if ch.nVar <= 50 else 0.9)  # 94.7 % at n = 70 (oracle and GPU alike)


def test_experimental_cold_and_shared_guess():
    """warmStart off: the guess is ignored (equalities of the data are still pre-factorised); a guess
    shared by the whole batch (stride 0) is accepted."""
    ch = P.config_B()
    pb = P.random_problems(ch, 512, seed=9)
    g = _gpu_warm(pb, None, warm=False)
    ref = _oracle_warm(pb, None, warm=False)
    assert_parity(g, ref)
    assert (g["status"] == 0).all() and P.test_kkt(g["x"], g["u"], pb).all()
    shared = np.zeros(pb.mc + pb.n, dtype=np.int8)
    shared[pb.mc:pb.mc + 3] = 4
    g2 = _gpu_warm(pb, shared)
    ref2 = _oracle_warm(pb, np.broadcast_to(shared, (pb.batch, shared.size)).copy())
    assert_parity(g2, ref2)


def test_experimental_overconstrained_and_max_iter():
    n = 4
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, n, n))
    G = A @ A.transpose(0, 2, 1) + np.eye(n)
    a = rng.normal(size=(3, n))
    Cm = rng.normal(size=(3, 6, n))
    bl = rng.normal(size=(3, 6))
    bu = bl.copy()            # six equalities on four variables
    bu[1, 4:] += 1.0          # instance 1: only four equalities
    sv = S.BatchedGoldfarbIdnaniSolver(n, 6, False, 3)
    sv.options(S.SolverOptions().warmStart(True))
    sv.solve(G, a, Cm, bl, bu, experimental=True)
    ref = po.solve_batch(G, a, Cm, bl, bu, experimental=True, warm_start=True)
    assert sv.last["status"].tolist() == ref["status"].tolist()
    assert sv.last["status"][0] == 6 and sv.last["status"][2] == 6  # OVERCONSTRAINED_PROBLEM


def test_experimental_class_mirror():
    """the reference's calling sequence (tests/GoldfarbIdnaniSolverTest.cpp:146-181) on the class mirror"""
    ch = P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=1, bounds=True, nStrongActBounds=2)
    pb = P.random_problems(ch, 8, seed=123)
    for k in range(pb.batch):
        G = pb.G[k].T.copy()
        cold = S.GoldfarbIdnaniSolver(pb.n, pb.mc, True)
        assert cold.solve(G.copy(), pb.a[k], pb.C[k].T, pb.bl[k], pb.bu[k], pb.xl[k], pb.xu[k]) == S.TerminationStatus.SUCCESS
        ws = S.experimental.GoldfarbIdnaniSolver(pb.n, pb.mc, True)
        ws.options(S.SolverOptions().warmStart(True))
        assert ws.solve(G.copy(), pb.a[k], pb.C[k].T, pb.bl[k], pb.bu[k], pb.xl[k], pb.xu[k], cold.activeSet()) == S.TerminationStatus.SUCCESS
        assert ws.iterations() == 0
        assert np.allclose(ws.solution(), pb.x[k], atol=1e-6)
        # reusing the previous active set
        assert ws.solve(G.copy(), pb.a[k], pb.C[k].T, pb.bl[k], pb.bu[k], pb.xl[k], pb.xu[k]) == S.TerminationStatus.SUCCESS
        assert ws.iterations() == 0
