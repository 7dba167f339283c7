"""CPU, world_size 2 over gloo: the N>1 host logic of the data-parallel runner (sharding, the single all-gather,
rank-count-invariant metrics).  The per-rank compute is CUDA-only and is covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_total, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from safediffcon_b200 import runner
    lo, hi = runner.shard_range(n_total, rank, ws)
    g = torch.Generator().manual_seed(0)
    J = torch.rand(n_total, generator=g)
    pts = torch.randint(0, 50, (n_total,), generator=g)
    tms = torch.randint(0, 11, (n_total,), generator=g)
    flg = (tms > 0).long()
    packed = torch.stack([J, pts.float(), tms.float(), flg.float()], dim=1)
    got = runner.all_gather_concat(packed[lo:hi].clone(), n_total)
    assert torch.equal(got, packed), "gathered vectors must come back in global index order"
    m = runner.metrics_from_vectors(got[:, 0], got[:, 1].long(), got[:, 2].long(), got[:, 3].long(), 11, 128)
    torch.save(m, os.path.join(out_dir, f"m{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 7])
def test_all_gather_and_metrics_rank_invariant(tmp_path, n_total):
    from safediffcon_b200 import runner
    ws = 2
    mp.spawn(_worker, args=(ws, _free_port(), n_total, str(tmp_path)), nprocs=ws, join=True)
    m0, m1 = torch.load(tmp_path / "m0.pt"), torch.load(tmp_path / "m1.pt")
    assert m0 == m1
    g = torch.Generator().manual_seed(0)
    J = torch.rand(n_total, generator=g)
    pts = torch.randint(0, 50, (n_total,), generator=g)
    tms = torch.randint(0, 11, (n_total,), generator=g)
    single = runner.metrics_from_vectors(J, pts, tms, (tms > 0).long(), 11, 128)
    assert single == m0


def test_shard_range_partitions():
    from safediffcon_b200 import runner
    for n in (0, 1, 7, 8192, 100000):
        for ws in (1, 2, 3, 8):
            spans = [runner.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
