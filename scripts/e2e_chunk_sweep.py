#!/usr/bin/env python
"""End-to-end host entry (jrlqp_solve_batch_host, pinned caller buffers) against the chunk size of its H2D / kernel / D2H
pipeline and the way G crosses the link (read in place by the kernels vs DMA upload of the lower triangle): one batch of
config A generated once, every setting timed on it (3 calls after one warm-up), results checked identical.
    python scripts/e2e_chunk_sweep.py [--batch 131072] [--chunks 1776,3552,7104,14208,28416]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jrl_qp_b200  # noqa: E402,F401
from jrl_qp_b200 import problems as P, solver as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=131072)
    ap.add_argument("--chunks", default="0,1776,3552,7104,14208,28416")
    ap.add_argument("--config", default="A")
    args = ap.parse_args()
    B = args.batch
    ch = {"A": P.config_A, "B": P.config_B}[args.config]()
    n, mc = ch.nVar, ch.nEq + ch.nIneq
    m = mc + n

    pinned = []

    def alloc(*shape):
        t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        pinned.append(t)
        return t.numpy()

    pb = P.random_problems(ch, B, seed=P.DEFAULT_SEED, out=alloc)
    hx = torch.empty((B, n), dtype=torch.float64, pin_memory=True)
    hu = torch.empty((B, m), dtype=torch.float64, pin_memory=True)
    hf = torch.empty(B, dtype=torch.float64, pin_memory=True)
    hit = torch.empty(B, dtype=torch.int32, pin_memory=True)
    hst = torch.empty(B, dtype=torch.int32, pin_memory=True)
    hact = torch.empty((B, m), dtype=torch.int8, pin_memory=True)
    lib = S.load_library()
    ref_x = None
    rows = []
    for zc in ("1", "0"):
        os.environ["JRLQP_G_ZEROCOPY"] = zc
        sv = S.BatchedGoldfarbIdnaniSolver(n, mc, True, B, device=0)
        prob = sv._problem(B, pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, set())
        res = S._Result(hx.data_ptr(), hu.data_ptr(), hf.data_ptr(), hit.data_ptr(), hst.data_ptr(), hact.data_ptr(), None, None, None)
        for ck in [int(c) for c in args.chunks.split(",")]:
            if ck:
                os.environ["JRLQP_HOST_CHUNK"] = str(ck)
            else:
                os.environ.pop("JRLQP_HOST_CHUNK", None)
            hx.zero_()
            assert lib.jrlqp_solve_batch_host(sv._h, C.byref(prob), C.byref(res)) == 0
            torch.cuda.synchronize()
            if ref_x is None:
                ref_x = hx.numpy().copy()
            same = bool(np.array_equal(ref_x, hx.numpy()))
            t0 = time.perf_counter()
            for _ in range(3):
                lib.jrlqp_solve_batch_host(sv._h, C.byref(prob), C.byref(res))
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            rows.append({"zero_copy_G": zc, "chunk": ck or "default", "ms_per_call": 1e3 * dt, "qp_per_s": B / dt, "identical": same})
            print(json.dumps(rows[-1]), flush=True)
        del sv
    os.environ.pop("JRLQP_HOST_CHUNK", None)
    os.environ.pop("JRLQP_G_ZEROCOPY", None)


if __name__ == "__main__":
    main()
