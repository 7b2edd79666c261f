"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: bank broadcast, batch sharding and
the ragged match gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from onepose_plus_plus_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    ref = {k: torch.randn(s, generator=g) for k, s in
           (("keypoints3d", (1, 50, 3)), ("descriptors3d_db", (1, 128, 50)), ("descriptors3d_coarse_db", (1, 256, 50)))}
    bank = {k: (v.clone() if rank == 0 else torch.zeros_like(v)) for k, v in ref.items()}
    parallel.broadcast_bank(bank, src=0)
    ok = all(torch.equal(bank[k], ref[k]) for k in ref)
    lo, hi = parallel.shard_range(9, rank, world)
    m = 3 + 2 * rank
    data = {"m_bids": torch.arange(m) % (hi - lo), "mkpts_3d_db": torch.full((m, 3), float(rank)),
            "mkpts_query_f": torch.full((m, 2), 10.0 + rank), "mconf": torch.full((m,), 0.5)}
    allm = parallel.gather_matches(data, lo)
    torch.save({"ok": ok, "range": (lo, hi), "all": allm}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_broadcast_shard_gather_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert all(r["ok"] for r in res)
    assert res[0]["range"] == (0, 5) and res[1]["range"] == (5, 9)
    assert torch.equal(res[0]["all"], res[1]["all"])
    allm = res[0]["all"]
    assert allm.shape == (3 + 5, 7)
    assert torch.equal(allm[:3, 1], torch.zeros(3)) and torch.equal(allm[3:, 1], torch.ones(5))
    assert allm[3:, 0].min().item() >= 5  # global image ids of rank 1 start at its offset


def test_shard_range_covers_everything():
    for n in (1, 7, 64, 513):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
