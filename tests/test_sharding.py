"""Host-side multi-GPU logic on CPU: locus partition and gather over gloo at world size 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mchap_b200.sharding import gather_by_locus, locus_block


def test_locus_block_matches_array_split():
    for n in (0, 1, 7, 10, 10000):
        for world in (1, 2, 3, 8):
            parts = np.array_split(np.arange(n), world)
            for r in range(world):
                start, stop = locus_block(n, r, world)
                np.testing.assert_array_equal(np.arange(start, stop), parts[r])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_loci, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop = locus_block(n_loci, rank, world)
    # a stand-in for per-locus device results: (locus id, a small array derived from it)
    local = [(i, np.full(3, i, dtype=np.int8)) for i in range(start, stop)]
    full = gather_by_locus(local, n_loci)
    ok = [x[0] for x in full] == list(range(n_loci)) and all((x[1] == x[0]).all() for x in full)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t[0]) == float(world)
    open(os.path.join(out_dir, "rank%d.%s" % (rank, "ok" if ok else "bad")), "w").close()
    dist.destroy_process_group()


def test_gather_by_locus_world_size_2(tmp_path):
    world, n_loci = 2, 11
    mp.spawn(_worker, args=(world, _free_port(), n_loci, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]
