#!/usr/bin/env python
"""Development aid: A/B timing of variant builds (scripts/build_variant.py) in ONE process on the same device-resident
batch, rounds interleaved so that box-to-box and drift noise cancels:
    python scripts/ab_variants.py --config A --batch 65536 v0 v1 v3 ...   (names of _build/libjrlqp_b200_<name>.so; 'main')"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import jrl_qp_b200  # noqa: F401,E402
from jrl_qp_b200 import build as B, problems as P, solver as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="A")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--scan", type=int, default=-1)
    ap.add_argument("--n", type=int, default=0, help="instead of --config: the test3 family of scripts/shape_sweep.py at this n")
    ap.add_argument("names", nargs="+")
    args = ap.parse_args()
    if args.n > 0:
        n_, ne = args.n, args.n * 20 // 100
        ch = P.ProblemCharacteristics(n_, ne, n_, min(n_ - ne, n_) * 30 // 100, 0, n_ * 10 // 100, 0, True, True)
        args.config = f"n{n_}"
    else:
        ch = {"A": P.config_A, "B": P.config_B, "D": P.config_D}[args.config]()
    pb = P.random_problems(ch, args.batch, seed=P.DEFAULT_SEED)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(getattr(pb, k)).to(dev) for k in ("G", "a", "C", "bl", "bu", "xl", "xu")}
    Bn, n, m = pb.batch, pb.n, pb.mc + pb.n
    solvers, outs = {}, {}
    S._check_abi = False
    for name in args.names:
        path = os.path.join(B.OUT, "libjrlqp_b200.so" if name == "main" else f"libjrlqp_b200_{name}.so")
        S._lib = None
        S.library_path = lambda p=path: p
        sv = S.BatchedGoldfarbIdnaniSolver(n, pb.mc, True, Bn)
        if args.scan >= 0:
            sv.set_scan_transposed(bool(args.scan))
        solvers[name] = sv
        outs[name] = (torch.empty((Bn, n), dtype=torch.float64, device=dev), torch.empty(Bn, dtype=torch.int32, device=dev))
    stream = torch.cuda.current_stream()

    def step(name):
        x, it = outs[name]
        solvers[name].solve_device(Bn, d["G"], d["a"], d["C"], d["bl"], d["bu"], d["xl"], d["xu"], x, iterations=it,
                                   stream=stream.cuda_stream)

    times = {k: [] for k in args.names}
    for name in args.names:
        for _ in range(2):
            step(name)
    torch.cuda.synchronize()
    for r in range(args.rounds):
        for name in args.names:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                step(name)
            e1.record(stream)
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1) * 1e-3 / args.steps)
    ref = outs[args.names[0]]
    for name in args.names:
        t = np.array(times[name])
        same = bool(torch.equal(outs[name][0], ref[0]) and torch.equal(outs[name][1], ref[1]))
        print(f"{args.config} {name:8s} {Bn / t.min():12.0f} QP/s best  {Bn / np.median(t):12.0f} median  rounds {np.round(Bn / t / 1e3).astype(int).tolist()} k  "
              f"identical_to_{args.names[0]}={same} regs={solvers[name].kernel_info()['regs_per_thread']} qps_per_sm={solvers[name].kernel_info()['qps_per_sm']}")


if __name__ == "__main__":
    main()
