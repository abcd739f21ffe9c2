"""CPU tests of the KKT-checker restatement (oracle/kkt_oracle.cpp) against the numpy test support
(jrl-qp_b200/problems.py::test_kkt, itself a restatement of src/test/kkt.cpp) and against the reference's own
use of it: the generator's planted (x, lambda) satisfy testKKT (tests/RandomProblemsTest.cpp:122-140)."""
import numpy as np
import pytest

import pyoracle as po
import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import problems as P

CHARACS = [
    P.ProblemCharacteristics(5),
    P.ProblemCharacteristics(5, nEq=2),
    P.ProblemCharacteristics(5, nIneq=8, nStrongActIneq=4),
    P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=1, bounds=True, nStrongActBounds=2),
    P.config_B(),
    P.config_A(),
]


@pytest.mark.parametrize("ch", CHARACS)
def test_planted_solution_satisfies_kkt(ch):
    pb = P.random_problems(ch, 64, seed=99)
    flags, resid, nfail = po.kkt_check_batch(pb.x, pb.lam, pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, x_ref=pb.x)
    assert nfail == 0 and (flags == 7).all()
    assert (resid[:, 0] <= resid[:, 1]).all() and (resid[:, 3] == 0).all()


@pytest.mark.parametrize("ch", CHARACS[2:])
def test_matches_numpy_restatement_on_perturbed_points(ch):
    pb = P.random_problems(ch, 128, seed=7)
    rng = np.random.default_rng(5)
    x = pb.x.copy()
    u = pb.lam.copy()
    # a third of the instances get a primal perturbation, a third a dual one
    x[0::3] += 1e-3 * rng.standard_normal(x[0::3].shape)
    u[1::3] += 1e-3 * rng.standard_normal(u[1::3].shape)
    flags, resid, nfail = po.kkt_check_batch(x, u, pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, x_ref=pb.x)
    ok = P.test_kkt(x, u, pb)
    assert np.array_equal((flags & 3) == 3, ok)
    assert nfail == int((flags != 7).sum())
    assert ((flags[2::3] & 7) == 7).all()  # untouched instances pass everything
    assert ((flags[0::3] & 4) == 0).all()  # 1e-3 perturbation is far from isApprox(1e-6)
    # residual against an independent dense evaluation
    m = pb.mc
    dL = np.einsum("bij,bj->bi", pb.G, x) + pb.a + np.einsum("bci,bc->bi", pb.C, u[:, :m])
    if pb.xl is not None:
        dL = dL + u[:, m:]
    assert np.allclose(resid[:, 0], np.abs(dL).max(axis=1), rtol=1e-9, atol=1e-12)


def test_threads_do_not_change_results():
    pb = P.random_problems(P.config_B(), 256, seed=3)
    r1 = po.kkt_check_batch(pb.x, pb.lam, pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, x_ref=pb.x, nthreads=1)
    r4 = po.kkt_check_batch(pb.x, pb.lam, pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, x_ref=pb.x, nthreads=4)
    assert np.array_equal(r1[0], r4[0]) and np.array_equal(r1[1], r4[1]) and r1[2] == r4[2]
