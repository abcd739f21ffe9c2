#!/usr/bin/env python
"""End-to-end multi-GPU measurement through jrlqp_multi_solve_batch_host (ONE process, one host batch), beside the
platform ceiling of the host link (jrlqp_measure_host_link: concurrent pinned-host <-> device copies).

    python scripts/multi_gpu_e2e.py [--config A] [--per-gpu 131072] [--gpus 1,2,4,8] [--steps 3]

For every GPU count g: host -> device, device -> host and bidirectional aggregate GB/s of plain cudaMemcpyAsync
(the ceiling), then QP/s of the multi-GPU host entry with G read in place over the link (zero-copy, the default for
n <= 64) and with G uploaded by DMA (JRLQP_G_ZEROCOPY=0) — the same batch of g * per_gpu QPs in pinned host memory,
results verified bit for bit against the one-GPU host entry on the first shard."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import jrl_qp_b200  # noqa: F401,E402
from jrl_qp_b200 import problems as P, solver as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="A")
    ap.add_argument("--per-gpu", type=int, default=131072)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--link-bytes", type=int, default=1 << 30)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    ndev = torch.cuda.device_count()
    gs = [g for g in (int(v) for v in args.gpus.split(",")) if g <= ndev]
    ch = {"A": P.config_A, "B": P.config_B, "D": P.config_D}[args.config]()
    n, mc = ch.nVar, ch.nEq + ch.nIneq
    m = mc + n
    Bmax = max(gs) * args.per_gpu
    pinned = []

    def alloc(*shape):
        t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        pinned.append(t)
        return t.numpy()

    t0 = time.time()
    # the first per_gpu QPs are generated, the other shards are copies of them (throughput does not care, and every
    # shard's result can then be checked against the first one's)
    p1 = P.random_problems(ch, args.per_gpu, seed=P.DEFAULT_SEED)
    rep = lambda v: None if v is None else np.copyto(alloc(Bmax, *v.shape[1:]).reshape(max(gs), *v.shape), v[None]) or pinned[-1].numpy()  # noqa: E731
    pb = P.ProblemBatch(rep(p1.G), rep(p1.a), rep(p1.C), rep(p1.bl), rep(p1.bu), rep(p1.xl), rep(p1.xu))
    print(f"# {Bmax} QPs (config {args.config}: {args.per_gpu} generated, replicated per shard) in pinned host memory in {time.time() - t0:.1f} s, {pb.input_bytes() / 1e9:.1f} GB", flush=True)
    hx = torch.empty((Bmax, n), dtype=torch.float64, pin_memory=True)
    hu = torch.empty((Bmax, m), dtype=torch.float64, pin_memory=True)
    hf = torch.empty(Bmax, dtype=torch.float64, pin_memory=True)
    hit = torch.empty(Bmax, dtype=torch.int32, pin_memory=True)
    hst = torch.empty(Bmax, dtype=torch.int32, pin_memory=True)
    hact = torch.empty((Bmax, m), dtype=torch.int8, pin_memory=True)
    lib = S.load_library()
    import ctypes as C
    rows = []
    ref_x = None
    for g in gs:
        link = {}
        for name, direction in (("h2d", 0), ("d2h", 1), ("bidir", 2)):
            agg, per = S.measure_host_link(g, nbytes=args.link_bytes, reps=4, direction=direction)
            link[name] = {"aggregate_gbs": round(agg, 1), "per_device_gbs": [round(v, 1) for v in per]}
        B = g * args.per_gpu
        row = {"gpus": g, "batch": B, "host_link": link}
        for mode, env in (("zero_copy_G", "-1"), ("dma_G", "0")):
            os.environ["JRLQP_G_ZEROCOPY"] = env  # read by jrlqp_create
            sv = S.MultiGpuGoldfarbIdnaniSolver(n, mc, True, B, n_devices=g)
            prob = sv._problem(B, pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, set())
            res = S._Result(hx.data_ptr(), hu.data_ptr(), hf.data_ptr(), hit.data_ptr(), hst.data_ptr(), hact.data_ptr(), None, None, None)
            for _ in range(4):  # warm-up: device staging is allocated, the shares of the shards settle
                rc = sv._host_call(prob, res, False)
                assert rc == 0, (rc, sv._last_error())
            t0 = time.perf_counter()
            for _ in range(args.steps):
                rc = sv._host_call(prob, res, False)
            dt = (time.perf_counter() - t0) / args.steps
            assert rc == 0
            if ref_x is None:
                ref_x = hx.numpy()[: args.per_gpu].copy()
            same = all(bool(np.array_equal(hx.numpy()[k * args.per_gpu:(k + 1) * args.per_gpu], ref_x)) for k in range(g))
            gbytes = sv.host_g_bytes(pinned=True)
            h2d = pb.input_bytes() / Bmax * B + B * (gbytes - 8 * n * n)
            row[mode] = {"qps": round(B / dt), "ms_per_step": round(1e3 * dt, 2), "h2d_gbs": round(h2d / dt / 1e9, 1), "all_shards_identical_to_1gpu": same,
                         "shares": [round(v, 3) for v in sv.weights()]}
            del sv
        rows.append(row)
        print(json.dumps(row), flush=True)
    base = rows[0]
    print("\n gpus |  H2D ceiling GB/s |  D2H | bidir | multi entry, G in place: QP/s  (GB/s, eff.) | G by DMA: QP/s (GB/s, eff.)")
    for r in rows:
        g = r["gpus"]
        z, dmm = r["zero_copy_G"], r["dma_G"]
        print(f" {g:4d} | {r['host_link']['h2d']['aggregate_gbs']:17.1f} | {r['host_link']['d2h']['aggregate_gbs']:5.1f} | {r['host_link']['bidir']['aggregate_gbs']:5.1f} |"
              f" {z['qps']:12d} ({z['h2d_gbs']:6.1f}, {z['qps'] / (g * base['zero_copy_G']['qps']):.2f}) | {dmm['qps']:10d} ({dmm['h2d_gbs']:6.1f}, {dmm['qps'] / (g * base['dma_G']['qps']):.2f})")
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
