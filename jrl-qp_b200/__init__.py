"""B200-native batched Goldfarb-Idnani QP solver (drop-in for jrl-qp's GoldfarbIdnaniSolver path).

The package directory is `jrl-qp_b200/`; import it as `jrl_qp_b200` through the shim module
`jrl_qp_b200.py` at the repository root.
"""
from . import build  # noqa: F401
