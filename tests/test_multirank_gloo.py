"""CPU, world_size 2 over gloo: the one-rank-per-GPU launch path of bench.py -- every rank takes
its rank_slice of the Gray index space, computes its partial sum (here with the CPU oracle standing
in for the device), the partials are gathered and added in rank order, and the result equals the
single-rank permanent.  No data-path collective is involved."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from _oracle import Oracle
    orc = Oracle()
    A = bench.synthetic_matrix(n, 0.5)
    total = 1 << (n - 1)
    lo, hi = bench.rank_slice(total, rank, world, align_log2=6)
    part = torch.tensor([orc.ryser_range_f64(A, lo, hi)], dtype=torch.float64)
    parts = [torch.zeros_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    s = 0.0
    for p in parts:           # fixed rank order
        s += float(p.item())
    t = torch.tensor([float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        perm = (-2.0 if n % 2 == 0 else 2.0) * s
        np.save(out_path, np.array([perm, orc.perm_ld(A), float(t.item()), float(total)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rank_sliced_permanent(tmp_path, world):
    n = 15
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    perm, want, covered, total = np.load(out)
    assert covered == total
    assert perm == pytest.approx(want, rel=1e-10)
