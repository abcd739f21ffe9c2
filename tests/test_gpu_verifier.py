"""GPU tests of the batch verifier (jrl-qp_b200/csrc/kkt.cu) through the C-ABI
(jrlqp_kkt_check_host / _device): flags and residuals bit-identical to the oracle restatement of
src/test/kkt.cpp, on planted solutions, on solver output, on perturbed points and on device-resident batches."""
import numpy as np
import pytest
import torch

import pyoracle as po
import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import problems as P, solver as S

pytestmark = pytest.mark.gpu

CHARACS = [
    P.ProblemCharacteristics(5),
    P.ProblemCharacteristics(5, nEq=2),
    P.ProblemCharacteristics(5, nIneq=8, nStrongActIneq=4),
    P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=1, bounds=True, nStrongActBounds=2),
    P.config_B(),
    P.config_A(),
    P.config_D(),
    P.ProblemCharacteristics(150, 10, 290, nStrongActIneq=20, bounds=True, nStrongActBounds=5, doubleSidedIneq=True),
]


def _both(pb, x, u, x_ref):
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, pb.xl is not None, 1)
    before = S.launch_count()
    g = sv.test_kkt(x, u, pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, x_ref=x_ref)
    assert S.launch_count() > before
    o = po.kkt_check_batch(x, u, pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, x_ref=x_ref, nthreads=8)
    return g, o


@pytest.mark.parametrize("ch", CHARACS)
def test_verifier_matches_oracle_bitwise(ch):
    B = 96 if ch.nVar > 100 else 512
    pb = P.random_problems(ch, B, seed=31)
    rng = np.random.default_rng(1)
    x, u = pb.x.copy(), pb.lam.copy()
    x[0::4] += 1e-4 * rng.standard_normal(x[0::4].shape)
    u[1::4] += 1e-4 * rng.standard_normal(u[1::4].shape)
    x[2::4] *= 1 + 1e-7  # inside the thresholds
    g, o = _both(pb, x, u, pb.x)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and g[2] == o[2]
    assert ((g[0][3::4]) == 7).all()
    assert g[2] == int((g[0] != 7).sum()) and g[2] > 0


def test_verifier_on_solver_output_and_without_reference():
    pb = P.random_problems(P.config_A(), 1024, seed=8)
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, pb.batch)
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    r = sv.last
    flags, resid, nfail = sv.test_kkt(r["x"], r["u"], pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    assert nfail == 0 and (flags == 3).all()
    assert P.test_kkt(r["x"], r["u"], pb).all()
    flags, resid, nfail = sv.test_kkt(r["x"], r["u"], pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, x_ref=pb.x)
    assert nfail == 0 and (flags == 7).all()  # tests/GoldfarbIdnaniSolverTest.cpp:94: x.isApprox(pb.x, 1e-6)
    o = po.kkt_check_batch(r["x"], r["u"], pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, x_ref=pb.x, nthreads=8)
    assert np.array_equal(resid, o[1])


def test_verifier_shared_arrays_and_no_constraints():
    # G, C and the bounds shared by the batch (stride 0); and a problem without general constraints
    pb = P.random_problems(P.config_B(), 1, seed=2)
    B = 64
    rng = np.random.default_rng(0)
    a = pb.a[0] + 1e-3 * rng.standard_normal((B, pb.n))
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, B)
    sv.solve(pb.G[0], a, pb.C[0], pb.bl[0], pb.bu[0], pb.xl[0], pb.xu[0])
    r = sv.last
    g = sv.test_kkt(r["x"], r["u"], pb.G[0], a, pb.C[0], pb.bl[0], pb.bu[0], pb.xl[0], pb.xu[0])
    o = po.kkt_check_batch(r["x"], r["u"], pb.G[0], a, pb.C[0], pb.bl[0], pb.bu[0], pb.xl[0], pb.xu[0])
    assert g[2] == 0 and np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
    pb2 = P.random_problems(P.ProblemCharacteristics(7, bounds=True, nStrongActBounds=2), 33, seed=4)
    g, o = _both(pb2, pb2.x, pb2.lam, pb2.x)
    assert g[2] == 0 and np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])


def test_verifier_device_resident_batch():
    """The use case: solve and verify on the device, only the failure count comes back."""
    pb = P.random_problems(P.config_B(), 4096, seed=12)
    dev = torch.device("cuda:0")
    t = lambda v: torch.from_numpy(v).to(dev)
    G, a, Cm, bl, bu, xl, xu = map(t, (pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu))
    B, n, m = pb.batch, pb.n, pb.mc + pb.nb
    x = torch.empty((B, n), dtype=torch.float64, device=dev)
    u = torch.empty((B, m), dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    flags = torch.empty(B, dtype=torch.int32, device=dev)
    nfail = torch.zeros(1, dtype=torch.int64, device=dev)
    xref = t(pb.x)
    sv = S.BatchedGoldfarbIdnaniSolver(n, pb.mc, True, 1)
    stream = torch.cuda.current_stream().cuda_stream
    sv.solve_device(B, G, a, Cm, bl, bu, xl, xu, x, u=u, status=status, stream=stream)
    sv.test_kkt_device(B, x, u, G, a, Cm, bl, bu, xl, xu, flags, n_fail=nfail, x_ref=xref, stream=stream)
    torch.cuda.synchronize()
    assert int(status.max()) == 0 and int(nfail.item()) == 0 and bool((flags == 7).all())
    # corrupt one solution on the device: exactly that instance is reported
    x[17, 3] += 1e-3
    nfail.zero_()
    sv.test_kkt_device(B, x, u, G, a, Cm, bl, bu, xl, xu, flags, n_fail=nfail, x_ref=xref, stream=stream)
    torch.cuda.synchronize()
    assert int(nfail.item()) == 1 and int(flags[17]) != 7 and int((flags != 7).sum()) == 1
