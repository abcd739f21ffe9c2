"""GPU parity tests of the global-workspace kernel (jrl-qp_b200/csrc/gi_large.cuh: 128 < n <= 1024, or any n
when selected with jrlqp_set_kernel_path(2)) against the CPU oracle, through the C-ABI.

Covers BASELINE.json config 3: the reference's MultiIK fixtures (tests/MultiIK.zip, committed losslessly
under tests/golden/ by tests/golden/make_golden.py; n = 387 / m = 1621 and n = 210 / m = 25 + 210 bounds,
tests/BlockGISolverTest.in.cpp:172-188,273-284) replicated with perturbed linear terms, G and C shared by the
batch (stride 0), cold and warm-started (pattern of benchmarks/SolversWarmStart.cpp:254-276).

Bar: status, iterations, active set identical; x, u, f within 1e-9 relative — and, because the kernel follows
the oracle's canonical operation order, bit-identical."""
import os

import numpy as np
import pytest

import pyoracle as po
import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import problems as P, solver as S
from test_gpu_parity import assert_parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle(pb, **kw):
    return po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count(), **kw)


def _solver(pb, path=2, warm=None, max_iter=None):
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, pb.xl is not None, pb.batch)
    sv.set_kernel_path(path)
    o = S.SolverOptions()
    if warm is not None:
        o.warmStart(warm)
    if max_iter is not None:
        o.maxIter(max_iter)
    sv.options(o)
    return sv


def _gpu(pb, path=2, want_L=False, **kw):
    sv = _solver(pb, path, **kw)
    before = S.launch_count()
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, want_L=want_L)
    assert S.launch_count() > before, "no CUDA kernel was launched"
    assert sv.kernel_info()["threads_per_qp"] == (256 if path == 2 or pb.n > 128 else 32 * ((pb.n + 31) // 32))
    return sv.last


def _gpu_warm(pb, as_in, warm=True, path=2, max_iter=None):
    sv = _solver(pb, path, warm=warm, max_iter=max_iter)
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, experimental=True, as_in=as_in)
    return sv.last


def _oracle_warm(pb, as_in, warm=True, max_iter=500):
    return po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count(), experimental=True,
                          warm_start=warm, as_in=as_in, max_iter=max_iter)


# ---------------------------------------------------------------------------------------------
# the two kernel families on the same small problems
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,B", [("config_B", 1024), ("config_A", 512), ("config_D", 64)])
def test_global_workspace_kernel_on_the_baseline_shapes(cfg, B):
    pb = P.random_problems(getattr(P, cfg)(), B)
    g = _gpu(pb, want_L=True)
    ref = _oracle(pb)
    assert_parity(g, ref)
    assert (g["status"] == 0).all() and P.test_kkt(g["x"], g["u"], pb).all()
    if pb.n <= 128:
        s = _gpu(pb, path=1, want_L=True)  # shared-memory kernel: bit-identical, Cholesky factor included
        for k in ("x", "u", "f", "iterations", "active_set", "L"):
            assert np.array_equal(g[k], s[k]), k


@pytest.mark.parametrize("n", [1, 2, 5, 33, 100, 129, 160, 257, 300])
def test_sizes_across_thread_and_tile_boundaries(n):
    ne = n // 5
    ni = n - ne
    ch = P.ProblemCharacteristics(n, ne, ni, min(ni, max(0, n // 4)), 0, min(n // 10, n - ne - min(ni, n // 4)), 0, True, True)
    B = 48 if n <= 160 else 12
    pb = P.random_problems(ch, B, seed=n)
    g = _gpu(pb, path=0 if n > 128 else 2)
    assert_parity(g, _oracle(pb))
    assert (g["status"] == 0).all()


def test_failure_statuses_fixed_variables_and_inconsistent_bounds():
    pb = P.random_problems(P.config_B(), 64, seed=5)
    pb.G[3] = -pb.G[3]  # not positive definite
    pb.C[30, 6] = pb.C[30, 5]
    pb.bl[30, 5], pb.bu[30, 5] = 1.0, 2.0
    pb.bl[30, 6], pb.bu[30, 6] = -2.0, -1.0  # contradictory pair -> INFEASIBLE
    pb.xl[40:, 3] = pb.xu[40:, 3] = pb.x[40:, 3]  # FIXED variables
    pb.bl[50:, 7], pb.bu[50:, 7] = pb.bu[50:, 7] + 0.5, pb.bl[50:, 7] - 0.5  # bl > bu: exact sequential scan
    g = _gpu(pb, max_iter=60)
    ref = _oracle(pb, max_iter=60)
    assert_parity(g, ref)
    assert g["status"][3] == S.TerminationStatus.NON_POS_HESSIAN and g["status"][30] == S.TerminationStatus.INFEASIBLE


def test_many_drops_partial_steps():
    ch = P.ProblemCharacteristics(12, 2, 14, 3, 3, 1, 2, True, True)  # weakly active constraints: partial steps, drops
    pb = P.random_problems(ch, 512, seed=3)
    g = _gpu(pb)
    assert_parity(g, _oracle(pb))


# ---------------------------------------------------------------------------------------------
# BASELINE config 3: MultiIK fixtures
# ---------------------------------------------------------------------------------------------
def _multiik_sequential(B, seed=0, scale=1e-3):
    d = np.load(os.path.join(GOLDEN, "multiik_sequential.npz"))
    rng = np.random.default_rng(seed)
    a = d["a"][None] * (1.0 + scale * rng.standard_normal((B, d["a"].size)))
    a[0] = d["a"]
    bu = d["u"]
    return P.ProblemBatch(np.ascontiguousarray(d["G"].T), a, np.ascontiguousarray(d["C"]), np.full_like(bu, -np.inf), bu, None, None), d


def _multiik_simultaneous(B, seed=0, scale=1e-3):
    d = np.load(os.path.join(GOLDEN, "multiik_simultaneous.npz"))
    rng = np.random.default_rng(seed)
    a = d["a"][None] * (1.0 + scale * rng.standard_normal((B, d["a"].size)))
    a[0] = d["a"]
    bu = d["u"]
    return P.ProblemBatch(np.ascontiguousarray(d["G"].T), a, np.ascontiguousarray(d["C"]), np.full_like(bu, -np.inf), bu, d["xl"], d["xu"]), d


def _solve_shared(pb, **kw):
    sv = _solver(pb, path=0, **{k: v for k, v in kw.items() if k in ("warm", "max_iter")})
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, experimental=kw.get("experimental", False), as_in=kw.get("as_in"))
    return sv.last


def test_multiik_sequential_cold_and_warm():
    pb, d = _multiik_sequential(40)
    ref = _oracle(pb)
    g = _solve_shared(pb)
    assert_parity(g, ref)
    assert (g["status"] == 0).all()
    # the reference's own check (tests/BlockGISolverTest.in.cpp:184-188): solution of the stored problem at 1e-4
    assert np.abs(g["x"][0] - d["sol"]).max() < 1e-4
    assert P.test_kkt(g["x"], g["u"], pb).all()
    # warm start from the previous instance's active set (the pattern of benchmarks/SolversWarmStart.cpp:254-276)
    guess = np.roll(g["active_set"], 1, axis=0)
    gw = _solve_shared(pb, warm=True, experimental=True, as_in=guess)
    rw = _oracle_warm(pb, guess)
    assert_parity(gw, rw)
    assert (gw["status"] == 0).all() and P.test_kkt(gw["x"], gw["u"], pb).all()
    same = (guess == g["active_set"]).all(axis=1)
    assert (gw["iterations"][same] == 0).all()  # an exact guess takes no iteration (tests/GoldfarbIdnaniSolverTest.cpp:176-181)
    assert np.allclose(gw["x"], g["x"], rtol=1e-6, atol=1e-8)


def test_multiik_simultaneous_cold_and_warm():
    pb, d = _multiik_simultaneous(64)
    ref = _oracle(pb)
    g = _solve_shared(pb)
    assert_parity(g, ref)
    assert (g["status"] == 0).all() and P.test_kkt(g["x"], g["u"], pb).all()
    gw = _solve_shared(pb, warm=True, experimental=True, as_in=g["active_set"])
    rw = _oracle_warm(pb, g["active_set"])
    assert_parity(gw, rw)
    assert (gw["iterations"] == 0).all()
    assert np.allclose(gw["x"], g["x"], rtol=1e-6, atol=1e-8)


def test_shared_G_is_factorised_once_and_matches_per_instance_factorisation():
    """G shared by the batch (stride 0) on the global-workspace kernel: one CTA factorises it once (same code, same
    bits) and the solver CTAs copy L^-T. Results must equal those of the same problems with G replicated per instance
    (every CTA factorising its own copy), the oracle's, and the factor returned must be the same."""
    torch = pytest.importorskip("torch")
    pb = P.random_problems(P.config_A(), 300, seed=5)
    rng = np.random.default_rng(1)
    a = pb.a[0] + 0.05 * rng.standard_normal(pb.a.shape)
    sh = P.ProblemBatch(pb.G[0], a, pb.C[0], pb.bl[0], pb.bu[0], pb.xl[0], pb.xu[0])
    rep = P.ProblemBatch(np.repeat(pb.G[:1], 300, axis=0), a, np.repeat(pb.C[:1], 300, axis=0), np.repeat(pb.bl[:1], 300, axis=0),
                         np.repeat(pb.bu[:1], 300, axis=0), np.repeat(pb.xl[:1], 300, axis=0), np.repeat(pb.xu[:1], 300, axis=0))
    ref = _oracle(rep)
    before = S.launch_count()
    g_sh = _gpu(sh, want_L=True)  # host entry: prefactor once, chunks re-use it
    assert S.launch_count() >= before + 2, "prefactor kernel + solver kernel"
    g_rep = _gpu(rep, want_L=True)
    assert_parity(g_sh, ref)
    assert_parity(g_rep, ref)
    assert np.array_equal(np.tril(g_sh["L"][7].T), np.tril(g_rep["L"][7].T))
    # warm start from the cold active sets: zero iterations, same points
    gw = _gpu_warm(sh, g_sh["active_set"])
    assert_parity(gw, _oracle_warm(rep, g_sh["active_set"]))
    assert (gw["iterations"] == 0).all()
    # device entry point (one launch = one prefactor + one solver kernel) and a warm-started sequence (prefactor once)
    sv = _solver(sh)
    dev = torch.device("cuda:0")
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev)  # noqa: E731
    x = torch.empty((300, pb.n), dtype=torch.float64, device=dev)
    it = torch.empty(300, dtype=torch.int32, device=dev)
    sv.solve_device(300, t(sh.G), t(sh.a), t(sh.C), t(sh.bl), t(sh.bu), t(sh.xl), t(sh.xu), x, iterations=it,
                    shared=("G", "C", "bl", "bu", "xl", "xu"))
    torch.cuda.synchronize()
    assert np.array_equal(x.cpu().numpy(), ref["x"]) and np.array_equal(it.cpu().numpy(), ref["iterations"])
    a_seq = np.stack([a, a * 1.01, a * 0.99])
    sv2 = _solver(sh, warm=True)
    sv2.solve_sequence(sh.G, a_seq, sh.C, sh.bl, sh.bu, sh.xl, sh.xu, warm=True)
    last = po.solve_batch(rep.G, a_seq[2], rep.C, rep.bl, rep.bu, rep.xl, rep.xu, nthreads=os.cpu_count())
    assert np.allclose(sv2.last["x"], last["x"], rtol=1e-9, atol=1e-11) and (sv2.last["status_worst"] == 0).all()
    # a shared G that is not positive definite: every instance reports NON_POS_HESSIAN
    bad = P.ProblemBatch(-sh.G, a, sh.C, sh.bl, sh.bu, sh.xl, sh.xu)
    gb = _gpu(bad)
    assert (gb["status"] == S.TerminationStatus.NON_POS_HESSIAN).all()


# ---------------------------------------------------------------------------------------------
# warm start on the global-workspace kernel
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,B", [("config_B", 1024), ("config_A", 256)])
def test_warm_start_exact_and_wrong_guesses(cfg, B):
    pb = P.random_problems(getattr(P, cfg)(), B, seed=77)
    cold = _gpu(pb)
    g = _gpu_warm(pb, cold["active_set"])
    assert_parity(g, _oracle_warm(pb, cold["active_set"]))
    assert (g["status"] == 0).all() and (g["iterations"] == 0).all()
    s = _gpu_warm(pb, cold["active_set"], path=1)
    for k in ("x", "u", "f", "iterations", "active_set"):
        assert np.array_equal(g[k], s[k]), k
    rng = np.random.default_rng(3)
    m = cold["active_set"].shape[1]
    guess = cold["active_set"].copy()
    swapped = guess.copy()
    swapped[guess == 1], swapped[guess == 2], swapped[guess == 4], swapped[guess == 5] = 2, 1, 5, 4
    kind_lower = np.where(np.arange(m) < pb.mc, 1, 4).astype(np.int8)
    swapped[guess == 0] = np.broadcast_to(kind_lower, guess.shape)[guess == 0]
    guess = np.where(rng.random(guess.shape) < 0.25, swapped, guess)
    g = _gpu_warm(pb, guess)
    assert_parity(g, _oracle_warm(pb, guess))
    g = _gpu_warm(pb, None, warm=False)
    assert_parity(g, _oracle_warm(pb, None, warm=False))


def test_warm_start_overconstrained():
    n = 4
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, n, n))
    G = A @ A.transpose(0, 2, 1) + np.eye(n)
    a = rng.normal(size=(3, n))
    Cm = rng.normal(size=(3, 6, n))
    bl = rng.normal(size=(3, 6))
    bu = bl.copy()
    bu[1, 4:] += 1.0
    sv = S.BatchedGoldfarbIdnaniSolver(n, 6, False, 3)
    sv.set_kernel_path(2)
    sv.options(S.SolverOptions().warmStart(True))
    sv.solve(G, a, Cm, bl, bu, experimental=True)
    ref = po.solve_batch(G, a, Cm, bl, bu, experimental=True, warm_start=True)
    assert sv.last["status"].tolist() == ref["status"].tolist()
    assert sv.last["status"][0] == 6 and sv.last["status"][2] == 6


def test_concurrent_launches_on_two_streams_do_not_share_a_workspace():
    import torch
    pb = P.random_problems(P.config_A(), 2048, seed=4)
    ref = _oracle(pb)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(getattr(pb, k)).to(dev) for k in ("G", "a", "C", "bl", "bu", "xl", "xu")}
    sv = _solver(pb)
    outs = []
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    half = pb.batch // 2
    torch.cuda.synchronize()
    for rep in range(3):
        for k, st in enumerate(streams):
            sl = slice(k * half, (k + 1) * half)
            x = torch.empty((half, pb.n), dtype=torch.float64, device=dev)
            it = torch.empty(half, dtype=torch.int32, device=dev)
            sv.solve_device(half, d["G"][sl], d["a"][sl], d["C"][sl], d["bl"][sl], d["bu"][sl], d["xl"][sl], d["xu"][sl], x,
                            iterations=it, stream=st.cuda_stream)
            outs.append((sl, x, it))
    torch.cuda.synchronize()
    for sl, x, it in outs:
        assert np.array_equal(x.cpu().numpy(), ref["x"][sl]) and np.array_equal(it.cpu().numpy(), ref["iterations"][sl])


# ---------------------------------------------------------------------------------------------
# TMA column ring (gi_large.cuh ring_*): the rotation sweep and z = J2 d2 read J through bulk copies
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [20, 50, 131, 212, 300, 512])
@pytest.mark.parametrize("warm", [False, True])
def test_tma_column_ring_is_invisible(n, warm):
    """Same bits with the ring forced on (JRLQP_LARGE_RING=2: every n <= 512 that fits, cold and warm kernels) and off
    (=0), and equal to the oracle: row counts below / at / above one and two rows per thread, column counts that are not
    a multiple of the stage width, sweeps shorter than one stage (many pre-activated equalities)."""
    ne = n // 3
    ch = P.ProblemCharacteristics(n, ne, n - ne, nStrongActIneq=max(1, n // 6), bounds=True, nStrongActBounds=n // 8,
                                  doubleSidedIneq=True)
    pb = P.random_problems(ch, 24 if n > 256 else 96, seed=700 + n)
    old = os.environ.get("JRLQP_LARGE_RING")
    old_slab = os.environ.get("JRLQP_LARGE_SLAB")
    out = {}
    try:
        for ring in ("2", "0"):
            os.environ["JRLQP_LARGE_RING"] = ring
            os.environ["JRLQP_LARGE_SLAB"] = "1" if ring == "2" else "0"  # (warm start: J = J Q on shared-memory slabs, on / off)
            if warm:
                cold = _oracle(pb)
                guess = np.roll(cold["active_set"], 1, axis=0)  # a neighbour's active set: a few iterations to repair it
                out[ring] = _gpu_warm(pb, guess)
            else:
                out[ring] = _gpu(pb)
    finally:
        for name, val in (("JRLQP_LARGE_RING", old), ("JRLQP_LARGE_SLAB", old_slab)):
            if val is None:
                os.environ.pop(name)
            else:
                os.environ[name] = val
    for k in ("x", "u", "f", "iterations", "status", "active_set"):
        assert np.array_equal(out["2"][k], out["0"][k]), k
    ref = _oracle_warm(pb, np.roll(_oracle(pb)["active_set"], 1, axis=0)) if warm else _oracle(pb)
    assert_parity(out["2"], ref)
