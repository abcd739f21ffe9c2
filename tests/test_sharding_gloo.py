"""CPU test of the N>1 path: world_size-2 gloo processes shard a batch by contiguous ranges with no
data-path collective; the only communication is the barrier + MAX/SUM reductions the benchmark uses."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import sharding


def test_shard_ranges_cover_batch_exactly():
    for B in (0, 1, 7, 8, 1000, 1 << 20):
        for w in (1, 2, 3, 4, 8):
            sh = sharding.all_shards(B, w)
            assert sh[0][0] == 0 and sh[-1][1] == B
            assert all(sh[i][1] == sh[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in sh]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import torch.distributed as dist
    import jrl_qp_b200  # noqa: F401
    from jrl_qp_b200 import problems as P, sharding as sh
    import pyoracle as po
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = sh.shard_range(B, rank, world)
    pb = P.random_problems(P.config_B(), hi - lo, seed=77, first_index=lo, nthreads=1)
    r = po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    dist.barrier()
    tmax = sh.reduce_max_time(1.0 + rank)
    total = sh.reduce_sum(hi - lo)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=r["x"], it=r["iterations"], lo=lo, hi=hi, tmax=tmax, total=total)
    dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_process(tmp_path):
    B, world = 37, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    import pyoracle as po
    from jrl_qp_b200 import problems as P
    pb = P.random_problems(P.config_B(), B, seed=77)
    ref = po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    xs, its = [], []
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert d["tmax"] == 2.0 and d["total"] == B
        xs.append(d["x"])
        its.append(d["it"])
    assert np.array_equal(np.concatenate(xs), ref["x"])
    assert np.array_equal(np.concatenate(its), ref["iterations"])
