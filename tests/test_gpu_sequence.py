"""GPU tests of the warm-started sequence mode (jrlqp_solve_sequence_host / _device): the loop of
benchmarks/SolversWarmStart.cpp:234-276 (same G, C, bounds; a(t) = cos(t w) p1 + sin(t w) v; every solve warm-started
from the previous active set) against the oracle driven step by step on the CPU. Bit-exact, every step."""
import os

import numpy as np
import pytest
import torch

import pyoracle as po
import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import problems as P, solver as S

pytestmark = pytest.mark.gpu


def _trajectory(pb, steps, seed):
    """geta() of the reference benchmark (benchmarks/SolversWarmStart.cpp:162-169)."""
    rng = np.random.default_rng(seed)
    B, n = pb.a.shape
    v = rng.standard_normal((B, n)) * np.linalg.norm(pb.a, axis=1, keepdims=True) / np.sqrt(n) * 0.25
    w = rng.uniform(0.5, 2.0, size=(B, 1))
    t = (np.arange(steps) / steps)[:, None, None]
    return np.cos(t * w) * pb.a + np.sin(t * w) * v  # [T, B, n]


def _oracle_sequence(pb, a_seq, warm, as0=None):
    out = []
    prev = as0
    for t in range(a_seq.shape[0]):
        if warm:
            r = po.solve_batch(pb.G, a_seq[t], pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count(),
                               experimental=True, warm_start=True, as_in=prev)
            prev = r["active_set"]
        else:
            r = po.solve_batch(pb.G, a_seq[t], pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count())
        out.append(r)
    return out


@pytest.mark.parametrize("cfg,B,T", [("config_B", 512, 12), ("config_A", 256, 8)])
@pytest.mark.parametrize("warm", [True, False])
def test_sequence_matches_stepwise_oracle(cfg, B, T, warm):
    pb = P.random_problems(getattr(P, cfg)(), B, seed=77)
    a_seq = _trajectory(pb, T, 5)
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, B)
    before = S.launch_count()
    st = sv.solve_sequence(pb.G, a_seq, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, warm=warm, keep_steps=True)
    assert S.launch_count() >= before + T
    g = sv.last
    ref = _oracle_sequence(pb, a_seq, warm)
    for t in range(T):
        for k in ("x", "u", "f", "iterations", "status"):
            assert np.array_equal(g[k][t], ref[t][k]), f"step {t}: {k} differs from the oracle"
    for k in ("active_set", "active_list", "n_active"):
        assert np.array_equal(g[k], ref[-1][k])
    assert np.array_equal(g["iterations_total"], sum(r["iterations"] for r in ref))
    assert np.array_equal(g["status_worst"], np.max([r["status"] for r in ref], axis=0))
    assert int(st) == int(g["status_worst"].max())
    if warm:
        cold = _oracle_sequence(pb, a_seq, False)
        # the point of warm starting (BENCH_GI_EX vs BENCH_GI): fewer iterations along a smooth trajectory
        assert g["iterations_total"].sum() < 0.8 * sum(r["iterations"].sum() for r in cold)


def test_sequence_last_step_only_and_initial_guess():
    pb = P.random_problems(P.config_B(), 300, seed=3)
    a_seq = _trajectory(pb, 6, 9)
    cold0 = po.solve_batch(pb.G, a_seq[0], pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count())
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, 300)
    sv.solve_sequence(pb.G, a_seq, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, warm=True, as_in=cold0["active_set"])
    g = sv.last
    ref = _oracle_sequence(pb, a_seq, True, as0=cold0["active_set"])
    assert ref[0]["iterations"].max() == 0  # exact guess for step 0 (tests/GoldfarbIdnaniSolverTest.cpp:176)
    for k in ("x", "u", "f", "iterations", "status", "active_set"):
        assert np.array_equal(g[k], ref[-1][k]), k
    assert np.array_equal(g["iterations_total"], sum(r["iterations"] for r in ref))


def test_sequence_device_pointers_shared_matrices():
    """Device entry point, G / C / bounds shared by the batch (one robot, many linear terms)."""
    pb1 = P.random_problems(P.config_B(), 1, seed=21)
    B, T, n, m = 256, 5, pb1.n, pb1.mc + pb1.nb
    rng = np.random.default_rng(2)
    a0 = pb1.a[0] * (1 + 0.05 * rng.standard_normal((B, n)))
    pb = P.ProblemBatch(pb1.G[0], a0, pb1.C[0], pb1.bl[0], pb1.bu[0], pb1.xl[0], pb1.xu[0])
    a_seq = _trajectory(pb, T, 4)
    dev = torch.device("cuda:0")
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev)
    G, Cm, bl, bu, xl, xu, aseq = map(t, (pb.G, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, a_seq))
    x = torch.empty((T, B, n), dtype=torch.float64, device=dev)
    u = torch.empty((B, m), dtype=torch.float64, device=dev)
    act = torch.empty((B, m), dtype=torch.int8, device=dev)
    tot = torch.empty(B, dtype=torch.int32, device=dev)
    worst = torch.empty(B, dtype=torch.int32, device=dev)
    sv = S.BatchedGoldfarbIdnaniSolver(n, pb1.mc, True, 1)
    sv.solve_sequence_device(B, T, G, aseq, Cm, bl, bu, xl, xu, x, act, u=u, iterations_total=tot, status_worst=worst,
                             warm=True, stream=torch.cuda.current_stream().cuda_stream,
                             shared=("G", "C", "bl", "bu", "xl", "xu"), step_strides={"x": B * n})
    torch.cuda.synchronize()
    ref = _oracle_sequence(pb, a_seq, True)
    for s in range(T):
        assert np.array_equal(x[s].cpu().numpy(), ref[s]["x"]), f"step {s}"
    assert np.array_equal(u.cpu().numpy(), ref[-1]["u"])
    assert np.array_equal(act.cpu().numpy(), ref[-1]["active_set"])
    assert np.array_equal(tot.cpu().numpy(), sum(r["iterations"] for r in ref))
    assert int(worst.max()) == 0


def _seq_with_env(value, fn):
    old = os.environ.get("JRLQP_SEQ_FCACHE")
    os.environ["JRLQP_SEQ_FCACHE"] = value
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop("JRLQP_SEQ_FCACHE")
        else:
            os.environ["JRLQP_SEQ_FCACHE"] = old


def test_sequence_factor_cache_is_invisible():
    """The steps t > 0 of a warm sequence re-read the factor step 0 stored in HBM (same G, same bits) instead of
    factorising again: every output of every step is identical with the cache on and off (JRLQP_SEQ_FCACHE is read
    when a handle is created), also after the handle has been used for a larger and a smaller batch."""
    pb = P.random_problems(P.config_A(), 192, seed=11)
    a_seq = _trajectory(pb, 7, 13)

    def run():
        sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, 192)
        out = []
        for nb in (64, 192, 100):  # grow, then shrink: the cache is sized by the largest call
            sl = slice(0, nb)
            sv.solve_sequence(pb.G[sl], a_seq[:, sl], pb.C[sl], pb.bl[sl], pb.bu[sl], pb.xl[sl], pb.xu[sl], warm=True, keep_steps=True)
            out.append({k: np.array(v, copy=True) for k, v in sv.last.items() if isinstance(v, np.ndarray)})
        return out

    on = _seq_with_env("1", run)
    off = _seq_with_env("0", run)
    for a, b in zip(on, off):
        assert a.keys() == b.keys()
        for k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
    ref = _oracle_sequence(pb, a_seq, True)
    for t in range(a_seq.shape[0]):
        assert np.array_equal(on[1]["x"][t], ref[t]["x"]) and np.array_equal(on[1]["iterations"][t], ref[t]["iterations"])


def test_sequence_factor_cache_failed_first_step_and_indefinite_G():
    """Instances whose step 0 stops before the factorisation (a guess with more than n equalities: OVERCONSTRAINED) store
    nothing and factorise at the first step that gets that far; an indefinite G is remembered as such."""
    ch = P.ProblemCharacteristics(8, 2, 10, nStrongActIneq=2, bounds=True, nStrongActBounds=1)
    pb = P.random_problems(ch, 48, seed=5)
    T = 5
    a_seq = _trajectory(pb, T, 3)
    m = pb.mc + pb.n
    as0 = np.zeros((48, m), dtype=np.int8)
    as0[::3, :pb.mc] = 3  # twelve EQUALITY guesses on eight variables
    G = pb.G.copy()
    G[1::7] = -G[1::7]  # not positive definite
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, 48)
    sv.solve_sequence(G, a_seq, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, warm=True, as_in=as0, keep_steps=True)
    g = sv.last
    prev = as0
    for t in range(T):
        r = po.solve_batch(G, a_seq[t], pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count(), experimental=True,
                           warm_start=True, as_in=prev)
        assert np.array_equal(g["status"][t], r["status"]), f"step {t}"
        ok = r["status"] == 0
        for k in ("x", "u", "f", "iterations"):
            assert np.array_equal(g[k][t][ok], r[k][ok]), f"step {t}: {k}"
        prev = np.where(ok[:, None], r["active_set"], 0).astype(np.int8)  # a failed solve hands over an empty set (write_failure)
        if t == 0:
            assert (r["status"][::3] == 6).all() and (r["status"][1::7][np.arange(1, 48, 7) % 3 != 0] == 2).all()
        else:
            assert ok[[i for i in range(48) if i % 7 != 1]].all()
