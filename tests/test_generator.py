"""CPU tests of the seeded problem generator (restatement of src/test/randomProblems.cpp), mirroring
tests/RandomProblemsTest.cpp:122-140: problems are well formed and the planted (x, lambda) satisfy KKT."""
import numpy as np
import pytest

import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import problems as P

CHARACS = [
    P.ProblemCharacteristics(5),
    P.ProblemCharacteristics(5, nEq=2),
    P.ProblemCharacteristics(5, nIneq=8, nStrongActIneq=4),
    P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=3),
    P.ProblemCharacteristics(5, 2, 6, nStrongActIneq=1, bounds=True, nStrongActBounds=2),
    P.ProblemCharacteristics(8, 2, 10, 2, 2, 1, 2, True, True),   # weakly active constraints and bounds
    P.ProblemCharacteristics(8, 0, 10, 3, 1, 0, 0, False, False),
    P.config_A(), P.config_B(),
]


@pytest.mark.parametrize("ch", CHARACS)
def test_well_formed_and_planted_kkt(ch):
    pb = P.random_problems(ch, 50, seed=42)
    n, mc = ch.nVar, ch.nEq + ch.nIneq
    assert pb.G.shape == (50, n, n) and pb.C.shape == (50, mc, n)
    assert np.array_equal(pb.G, pb.G.transpose(0, 2, 1))  # exactly symmetric
    assert (pb.bl <= pb.bu).all()
    assert (pb.bl[:, :ch.nEq] == pb.bu[:, :ch.nEq]).all()  # equalities first (src/test/problems.cpp:21-25)
    if ch.nIneq:
        assert (pb.bl[:, ch.nEq:] < pb.bu[:, ch.nEq:]).all()
    if ch.bounds:
        assert (pb.xl <= pb.xu).all()
    if not ch.doubleSidedIneq and ch.nIneq:
        assert np.isneginf(pb.bl[:, ch.nEq:]).all()
    assert np.linalg.eigvalsh(pb.G).min() > 0
    assert P.test_kkt(pb.x, pb.lam, pb).all()
    nact = (np.abs(pb.lam) > 0).sum(axis=1)
    assert (nact == ch.nEq + ch.nStrongActIneq + ch.nStrongActBounds).all()


def test_streams_are_reproducible_and_shardable():
    ch = P.config_B()
    full = P.random_problems(ch, 40, seed=5, nthreads=3)
    again = P.random_problems(ch, 40, seed=5, nthreads=1)
    assert np.array_equal(full.G, again.G) and np.array_equal(full.bl, again.bl)
    part = P.random_problems(ch, 15, seed=5, first_index=25)
    assert np.array_equal(full.G[25:], part.G) and np.array_equal(full.x[25:], part.x)
    other = P.random_problems(ch, 4, seed=6)
    assert not np.array_equal(full.G[:4], other.G)


def test_inconsistent_characteristics_rejected():
    with pytest.raises(ValueError):
        P.random_problems(P.ProblemCharacteristics(5, nEq=6), 1)
    with pytest.raises(ValueError):
        P.random_problems(P.ProblemCharacteristics(5, nIneq=2, nStrongActIneq=3), 1)
    with pytest.raises(ValueError):
        P.random_problems(P.ProblemCharacteristics(5, nStrongActBounds=1), 1)
