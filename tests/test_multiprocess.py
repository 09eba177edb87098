"""world_size-2 gloo test of the multi-GPU host logic (sharding by index, table broadcast, bitmap all-gather).
The evaluator on each rank is the C oracle (there is no GPU here); on the GPU box the same plumbing runs
over NCCL in bench.py."""
import os
import socket

import numpy as np
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


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import workloads as W
    from cerbos_b200.dist import all_gather_bitmaps, broadcast_blob, shard_range
    from cerbos_b200.encode import Encoder, manifest_from_blob
    from oracle import cref
    w = W.C2()
    blob = None
    if rank == 0:
        _, ft, _ = W.build(w)
        blob = ft.blob
    blob = broadcast_blob(blob, "cpu")
    enc = Encoder(manifest_from_blob(blob))
    lo, hi = shard_range(n_total, rank, world)
    b = w.columns(w.fields(hi - lo, start=lo), enc)
    eff = cref.check(blob, b.columns, b.n, b.max_actions)
    bits = np.packbits((eff == 1).astype(np.uint8), axis=1, bitorder="little").reshape(-1)
    full = all_gather_bitmaps(torch.from_numpy(bits.copy()))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    dist.destroy_process_group()


def test_sharded_evaluation_matches_single_process(tmp_path):
    n_total, world = 4096, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    import workloads as W
    from oracle import cref
    w = W.C2()
    _, ft, enc = W.build(w)
    b = w.columns(w.fields(n_total), enc)
    eff = cref.check(ft.blob, b.columns, b.n, b.max_actions)
    want = np.packbits((eff == 1).astype(np.uint8), axis=1, bitorder="little").reshape(-1)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        assert (got == want).all(), f"rank {r} does not hold the full result"


def test_shard_ranges_cover_everything():
    from cerbos_b200.dist import shard_range
    for n in (1, 7, 1024, 1 << 20):
        for world in (1, 2, 3, 8):
            pieces = [shard_range(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(world - 1))
