"""Quick GPU-vs-oracle parity probe (development aid; the real tests are under tests/)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import pyoracle as po
import jrl_qp_b200
from jrl_qp_b200 import problems as P, solver as S

def run(ch, B, stage=-1):
    pb = P.random_problems(ch, B)
    t = time.time(); ref = po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count()); tc = time.time() - t
    sv = S.BatchedGoldfarbIdnaniSolver(ch.nVar, ch.nEq + ch.nIneq, ch.bounds, B)
    sv.set_stage_c(stage)
    print(sv.kernel_info())
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    t = time.time(); sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu); tg = time.time() - t
    g = sv.last
    for k in ("status", "iterations", "n_active", "active_set", "active_list"):
        neq = (g[k] != ref[k]).reshape(B, -1).any(axis=1).sum()
        print(f"  {k}: mismatching instances {neq}/{B}")
    for k in ("x", "u", "f"):
        d = np.abs(g[k] - ref[k]).max(); bit = (g[k] == ref[k]).all()
        print(f"  {k}: max abs diff {d:.3e} bit-exact {bit}")
    print(f"  n={ch.nVar} B={B}: oracle {B/tc:.0f} QP/s ({os.cpu_count()} thr), gpu host-API {B/tg:.0f} QP/s; kkt ok {P.test_kkt(g['x'], g['u'], pb).mean():.4f}")
    bad = np.nonzero((g["iterations"] != ref["iterations"]) | (g["status"] != ref["status"]))[0]
    if len(bad):
        b = bad[0]; print("  first bad", b, g["status"][b], ref["status"][b], g["iterations"][b], ref["iterations"][b], g["active_list"][b], ref["active_list"][b])

if __name__ == "__main__":
    print("fp64 probe TFLOP/s:", S.measure_fp64_tflops())
    run(P.ProblemCharacteristics(5, 2, 6, 1, 0, 2, 0, True, False), 64)
    run(P.config_B(), 4096, 0); run(P.config_B(), 4096, 1)
    run(P.config_A(), 4096, 0); run(P.config_A(), 4096, 1)
    run(P.config_D(), 512)
