"""Multi-GPU check of the fused all-gather (run under torchrun, one rank per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_gather_check.py

Every rank evaluates its shard of a C2 batch twice: through cgpu_check_device (local bitmap, then NCCL all-gather as the
reference exchange) and through cgpu_check_device_gather (results stored straight into every rank's gather buffer over
peer memory).  Both gathered images must be identical on every rank, for several steps and rotating buffers."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cerbos_b200 import capi  # noqa: E402
import workloads as W
from cerbos_b200.device import DeviceBatch  # noqa: E402
from cerbos_b200.dist import PeerGather, all_gather_bitmaps  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    w = W.C2()
    _, ft, enc = W.build(w)
    ctx = capi.Context(local)
    table = ctx.load_table(ft.blob)
    table.wait_ready()
    n = 1 << 16
    n_buf = 3
    pg = PeerGather(ctx, n * 1, n_buf)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sh = stream.cuda_stream
    bad = 0
    for step in range(7):
        b = w.columns(w.fields(n, start=(step * world + rank) * n), enc)
        db = DeviceBatch(b, f"cuda:{local}")
        db.run(table, stream=sh)
        ctx.sync(sh)
        ref = all_gather_bitmaps(db.bitmap[: n]).cpu().numpy()
        j = step % n_buf
        call = table.prepared_gather_call(db.ptrs, db.sizes, db.n, db.max_actions, pg.bufs[j], pg.lane_flags(0), rank, n)
        call(step + 1, sh, step if step >= 1 else 0)   # also exercises the in-launch wait for the previous step
        pg.wait(step + 1, sh)
        ctx.sync(sh)
        got = pg.read(j)
        if not np.array_equal(got, ref):
            bad += 1
            print(f"[rank {rank}] step {step}: gathered image differs in {(got != ref).sum()} bytes", flush=True)
        dist.barrier()
    t = torch.tensor([bad], device=f"cuda:{local}")
    dist.all_reduce(t)
    if rank == 0:
        print("MGPU_GATHER_OK" if int(t.item()) == 0 else f"MGPU_GATHER_FAILED {int(t.item())}", flush=True)
    pg.close()
    table.release()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
