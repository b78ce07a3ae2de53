"""CPU, world_size=2 over gloo: the N>1 host logic (batch sharding, gather, global max).  The solve
itself needs no collective: a rank's oracle solution of its shard equals the slice of the full solve."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import gen_problem


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import lqr_oracle as orc
        from mpc.pytorch_b200.parallel import shard_problem, shard_range, gather_batch, global_max
        T, n, m = 6, 4, 2
        C, c, F, f, x0 = gen_problem(99, B, T, n, m, torch.float64)
        u = torch.zeros(T, B, m, dtype=torch.float64)
        x = orc.get_traj(T, u, x0, F, f)
        full = orc.lqr_step_forward(n, m, T, x0, C, c, F, f, x, u, u_lower=-0.2, u_upper=0.2, coupled=False)
        sh = shard_problem(rank, world, x0, C, c, F, f, cur_x=x, cur_u=u)
        lo, hi = shard_range(B, rank, world)
        assert sh["C"].is_contiguous() and sh["C"].shape[1] == hi - lo
        mine = orc.lqr_step_forward(n, m, T, sh["x_init"], sh["C"], sh["c"], sh["F"], sh["f"], sh["cur_x"],
                                    sh["cur_u"], u_lower=-0.2, u_upper=0.2, coupled=False)
        assert torch.equal(mine.new_u, full.new_u[:, lo:hi])          # bit-identical: no coupling
        got_u = gather_batch(mine.new_u, 1, B)
        got_c = gather_batch(mine.costs, 0, B)
        assert torch.equal(got_u, full.new_u) and torch.equal(got_c, full.costs)
        true_norm = (u - full.new_u).pow(2).sum((0, 2)).sqrt()
        gm = global_max(true_norm[lo:hi].max())
        assert float(gm) == float(true_norm.max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_two_rank_batch_sharding(B):
    mp.spawn(_worker, args=(2, _free_port(), B), nprocs=2, join=True)


def test_shard_ranges_partition_the_batch():
    from mpc.pytorch_b200.parallel import shard_range
    for B in (1, 7, 8, 4096, 32768):
        for W in (1, 2, 4, 8):
            r = [shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
