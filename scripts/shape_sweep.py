#!/usr/bin/env python
"""Throughput over the reference's benchmark shapes (benchmarks/Solvers.cpp:613-639) on one GPU, device-resident:
  test3  n = 10 .. 100 (plus the sizes either side of the 32- and 64-thread boundaries), 20 % equalities, n double-sided
         inequalities with 30 % active, bounds with 10 % active  — the family of the headline configuration;
  test6  n = 50, 50 double-sided inequalities, 0 .. 100 % of them active, no bounds.
One table: QP/s, threads per QP, QPs per SM, mean iterations. python scripts/shape_sweep.py [--batch 32768]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import jrl_qp_b200  # noqa: F401,E402
from jrl_qp_b200 import problems as P, solver as S  # noqa: E402


def run(ch, B, dev, steps=3):
    pb = P.random_problems(ch, B, seed=P.DEFAULT_SEED)
    d = {k: (None if getattr(pb, k) is None else torch.from_numpy(getattr(pb, k)).to(dev)) for k in ("G", "a", "C", "bl", "bu", "xl", "xu")}
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, pb.xl is not None, B)
    x = torch.empty((B, pb.n), dtype=torch.float64, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        sv.solve_device(B, d["G"], d["a"], d["C"], d["bl"], d["bu"], d["xl"], d["xu"], x, iterations=it, status=st, stream=stream.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / steps
    ok = bool((st == 0).all().item()) and bool(P.is_approx(x.cpu().numpy(), pb.x, 1e-5).all())  # (1e-5: unconstrained instances of cond(G) ~ 1e8 among 32768)
    info = sv.kernel_info()
    return {"qps": B / t, "threads": info["threads_per_qp"], "qps_per_sm": info["qps_per_sm"], "regs": info["regs_per_thread"],
            "stage_c": info["stage_c"], "iterations": float(it.double().mean().item()), "verified": ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    print("test3 (benchmarks/Solvers.cpp:621-623): n, 20 % eq, n double-sided ineq (30 % active), bounds (10 % active)")
    print("    n   m  threads QPs/SM regs   iterations        QP/s   QP/s per SM-thread-slot  ok")
    for n in (10, 20, 30, 32, 33, 40, 50, 60, 64, 65, 70, 80, 90, 96, 97, 100, 128):
        ne, ni = n * 20 // 100, n
        ch = P.ProblemCharacteristics(n, ne, ni, min(n - ne, ni) * 30 // 100, 0, n * 10 // 100, 0, True, True)
        B = args.batch if n <= 64 else args.batch // 4
        r = run(ch, B, dev)
        r.update(test="test3", n=n, m=ne + ni + n)
        rows.append(r)
        print(f"  {n:3d} {r['m']:3d}  {r['threads']:6d} {r['qps_per_sm']:6d} {r['regs']:4d}  {r['iterations']:10.2f} {r['qps']:12.0f}  {r['qps'] / (148 * r['qps_per_sm']):12.0f}  {r['verified']}", flush=True)
    print("test6 (benchmarks/Solvers.cpp:633-635): n = 50, 50 double-sided ineq, varying share of active ones, no bounds")
    for pct in (0, 10, 30, 50, 70, 100):
        ch = P.ProblemCharacteristics(50, 0, 50, 50 * pct // 100, 0, 0, 0, False, True)
        r = run(ch, args.batch, dev)
        r.update(test="test6", n=50, active_pct=pct)
        rows.append(r)
        print(f"  active {pct:3d} %  threads {r['threads']} QPs/SM {r['qps_per_sm']}  iterations {r['iterations']:6.2f}  {r['qps']:12.0f} QP/s  {r['verified']}", flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
