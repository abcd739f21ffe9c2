#!/usr/bin/env python
"""End-to-end structured QP solves/sec through the BlockGISolver path (BASELINE.json configs[4]: MPC
horizon QP, 32 blocks of 12 x 12, batch 128k): jrlqp_blockgi_solve_device on
device-resident inputs, CUDA events on the launching stream, beside (i) the same problems through
the dense kernel path (jrlqp_solve_batch_device on the dense n x n matrices, a bounded sample) and
(ii) the oracle port of experimental::BlockGISolver on the host cores.

    python scripts/bench_blockgi.py [--type tri|down|up] [--blocks 32] [--size 12] [--batch 131072]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import jrl_qp_b200  # noqa: F401,E402
from jrl_qp_b200 import solver as S  # noqa: E402
from jrl_qp_b200.blockgi import BatchedBlockGISolver  # noqa: E402
from jrl_qp_b200.structured import Type  # noqa: E402


def main(argv=None, emit=True):
    ap = argparse.ArgumentParser()
    ap.add_argument("--type", default="tri", choices=["tri", "down", "up"])
    ap.add_argument("--blocks", type=int, default=32)
    ap.add_argument("--size", type=int, default=12)
    ap.add_argument("--cstr", type=int, default=12, help="double-sided inequalities per block")
    ap.add_argument("--batch", type=int, default=131072)
    ap.add_argument("--base", type=int, default=512, help="distinct generated problems (tiled to the batch)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--dense-sample", type=int, default=2048)
    args = ap.parse_args(argv)
    import block_cases as bc
    import pyoracle as po
    type = {"tri": Type.TriBlockDiagonal, "down": Type.BlockArrowDown, "up": Type.BlockArrowUp}[args.type]
    sizes, mi = [args.size] * args.blocks, [args.cstr] * args.blocks
    t0 = time.perf_counter()
    pb = bc.random_block_problem(type, sizes, mi, args.base, seed=20261017, shift=0.05)
    gen_s = time.perf_counter() - t0
    B, n, mc = args.batch, pb.n, pb.mc
    dev = torch.device("cuda", 0)
    reps = (B + args.base - 1) // args.base

    def tile(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev).repeat(reps, 1)[:B].contiguous()

    # instance k = base problem (k mod base) scaled by s_k > 0 on (a, bl, bu): same G and C, solution s_k x*
    # (distinct data and distinct floating-point trajectories, same active sets)
    G0, a, Cd, bl, bu = tile(pb.Gdata), tile(pb.a), tile(pb.Cdata), tile(pb.bl), tile(pb.bu)
    s = 1.0 + 0.5 * torch.rand(B, 1, dtype=torch.float64, device=dev)
    a *= s
    bl *= s
    bu *= s
    G = G0.clone()
    x = torch.empty((B, n), dtype=torch.float64, device=dev)
    u = torch.empty((B, mc), dtype=torch.float64, device=dev)
    f = torch.empty(B, dtype=torch.float64, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    act = torch.empty((B, mc), dtype=torch.int8, device=dev)
    sv = BatchedBlockGISolver(pb.stG, pb.stC, False, B)
    stream = torch.cuda.current_stream()
    launches0 = S.launch_count()

    def step():
        sv.solve_device(B, G, a, Cd, bl, bu, None, None, x, u=u, f=f, iterations=it, status=status, active_set=act,
                        stream=stream.cuda_stream)

    ts = []
    for i in range(args.warmup + args.steps):
        G.copy_(G0)  # the device entry factorises G in place (what the reference leaves in G)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= args.warmup:
            ts.append(e0.elapsed_time(e1) * 1e-3)
    t_gpu = float(np.mean(ts))
    launches = (S.launch_count() - launches0) // (args.warmup + args.steps)
    all_ok = bool((status == 0).all().item())
    mean_it = float(it.double().mean().item())
    # planted solution (scaled) and oracle parity on a sample
    cs = min(B, args.cpu_sample)
    xs = x[:cs].cpu().numpy()
    sc = s[:cs].cpu().numpy()
    idx = np.arange(cs) % args.base
    planted = bool(np.abs(xs - sc * pb.x_planted[idx]).max() <= 1e-6 * max(1.0, np.abs(pb.x_planted).max()))
    cores = os.cpu_count() or 1
    ah, blh, buh = a[:cs].cpu().numpy(), bl[:cs].cpu().numpy(), bu[:cs].cpu().numpy()
    Gh, Ch = pb.Gdata[idx], pb.Cdata[idx]
    po.block_solve_batch(pb.stG, pb.stC, Gh[:64], ah[:64], Ch[:64], blh[:64], buh[:64], nthreads=cores)
    t0 = time.perf_counter()
    ref = po.block_solve_batch(pb.stG, pb.stC, Gh, ah, Ch, blh, buh, nthreads=cores)
    t_cpu = time.perf_counter() - t0
    parity = bool(np.array_equal(ref["x"], xs) and np.array_equal(ref["iterations"], it[:cs].cpu().numpy())
                  and np.array_equal(ref["active_set"], act[:cs].cpu().numpy()))
    # the same problems as dense n x n QPs through the dense kernel path (large-n kernel for n > 128)
    ds = min(B, args.dense_sample)
    dense = None
    try:
        idd = np.arange(ds) % args.base
        Gd = torch.from_numpy(np.ascontiguousarray(pb.Gdense[idd])).to(dev)
        Cdn = torch.from_numpy(np.ascontiguousarray(pb.Cdense[idd])).to(dev)
        dsv = S.BatchedGoldfarbIdnaniSolver(n, mc, False, ds)
        xd = torch.empty((ds, n), dtype=torch.float64, device=dev)
        std = torch.empty(ds, dtype=torch.int32, device=dev)
        td = []
        for i in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            dsv.solve_device(ds, Gd, a[:ds], Cdn, bl[:ds], bu[:ds], None, None, xd, status=std, stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            td.append(e0.elapsed_time(e1) * 1e-3)
        dense = {"value": ds / min(td[1:]), "unit": "QP/s", "sample": f"first {ds} instances as dense {n}x{n} QPs",
                 "max_abs_dx_vs_structured": float((xd - x[:ds]).abs().max().item()), "all_success": bool((std == 0).all().item())}
    except Exception as e:  # noqa: BLE001 - the comparison arm is optional
        dense = {"unavailable": repr(e)[:200]}
    info = sv.info()
    bytes_in = 8 * (info["g_elements_per_instance"] + n + pb.stC.stride + 2 * mc)
    bytes_out = 8 * (n + mc + 1) + 8 + mc
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    line = {
        "metric": "structured GI QP solves/sec (BlockGISolver)", "value": B / t_gpu, "unit": "QP/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_gpu, "higher_is_better": True, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"config E: {args.type} G, {args.blocks} blocks of {args.size}x{args.size} (n={n}), block-diagonal C with "
                               f"{args.cstr} double-sided inequalities per block (mc={mc}), batch {B} ({args.base} generated problems, "
                               "each instance scaled by its own factor)",
                   "kernel": info, "generator_s": gen_s},
        "roofline": {"bound": "hbm", "achieved": (bytes_in + bytes_out) * B / t_gpu / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": (bytes_in + bytes_out) * B / t_gpu / 1e9 / hbm, "traffic": None, "bytes_per_qp": bytes_in + bytes_out,
                     "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "dense_path": dense,
        "cpu_baseline": {"value": cs / t_cpu, "unit": "QP/s", "cores": cores, "kind": "port",
                         "sample": f"first {cs} instances, oracle port of experimental::BlockGISolver, one solver per thread"},
        "gpu_launches": int(launches),
        "verified": {"all_success": all_ok, "planted_solution": planted, "oracle_bit_exact_sample": parity, "mean_iterations": mean_it},
    }
    if emit:
        print(json.dumps(line))
    return line


if __name__ == "__main__":
    main()
