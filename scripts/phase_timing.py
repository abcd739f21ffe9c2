#!/usr/bin/env python
"""Per-phase cycle accounting of the dense kernel (debug build with -DJRLQP_PHASE_TIMING).
Builds a separate library (_build/libjrlqp_b200_timing.so), solves one batch, prints, per warp,
the share of wall time spent in every phase (barrier waits are charged to the phase that ends with them).

    python scripts/phase_timing.py [--config A] [--batch 16384]
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import jrl_qp_b200  # noqa: F401,E402
from jrl_qp_b200 import build as B, problems as P, solver as S  # noqa: E402

PHASES = ["init", "select", "fetch normal", "d = J^T n", "z = J2 d2", "own recurrence", "wait other recurrence",
          "step length", "take step", "add (apply rotations)", "remove", "loop control", "write result"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="A")
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--no-build", action="store_true", help="use the _build/libjrlqp_b200_timing.so built beforehand (no nvcc run on the GPU box)")
    args = ap.parse_args()
    out = os.path.join(B.OUT, "libjrlqp_b200_timing.so")
    srcs = B._listdir(B.CSRC, (".cu",))
    if not (args.no_build and os.path.exists(out)):
        subprocess.run([B.NVCC] + B.NVCC_FLAGS + ["-DJRLQP_PHASE_TIMING"] + srcs + ["-o", out], check=True, capture_output=True)
    S.library_path = lambda: out
    S._lib = None
    ch = {"A": P.config_A, "B": P.config_B, "D": P.config_D}[args.config]()
    pb = P.random_problems(ch, args.batch, seed=P.DEFAULT_SEED)
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, pb.xl is not None, pb.batch)
    lib = S.load_library()
    buf = (C.c_uint64 * 64)()
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    lib.jrlqp_debug_phase_cycles(sv._h, buf)
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    assert lib.jrlqp_debug_phase_cycles(sv._h, buf) == 0
    c = np.array(list(buf), dtype=np.float64).reshape(4, 16) / args.batch
    info = sv.kernel_info()
    print(f"config {args.config}: threads/QP {info['threads_per_qp']}, cycles per QP per warp:")
    for w in range(info["threads_per_qp"] // 32):
        tot = c[w].sum()
        print(f" warp {w}: total {tot:10.0f} cycles")
        for k, name in enumerate(PHASES):
            print(f"    {name:26s} {c[w, k]:10.0f}  {100 * c[w, k] / tot:5.1f}%")


if __name__ == "__main__":
    main()
