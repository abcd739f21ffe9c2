#!/usr/bin/env python
"""FP64 tensor-core (DMMA) experiment: python scripts/dmma_probe.py  — see jrl-qp_b200/csrc/dmma_probe.cu."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jrl_qp_b200  # noqa: F401,E402
from jrl_qp_b200 import solver as S  # noqa: E402

r = S.probe_dmma(0, int(sys.argv[1]) if len(sys.argv) > 1 else 64)
r["fp64_peak_tflops_measured"] = S.measure_fp64_tflops(0)
r["dmma_over_pipe"] = r["dmma_gflops"] / r["fp64_pipe_gflops"]
print(json.dumps(r, indent=1))
