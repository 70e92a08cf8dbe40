"""Row (e) on real GPUs: two processes, one GPU each, the library's own NCCL gather (mb200_gather_results) of variable-length
device buffers to rank 0, including the grow-and-retry path when the root's buffer is too small.  Needs >= 2 GPUs
(`gpurun --gpus 2`); skipped on a one-GPU box.  The N>1 host logic is also covered on CPU (tests/test_shard_gloo.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    import manta_b200
    from manta_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # bootstrap only: carries the 128-byte NCCL id
    torch.cuda.set_device(rank)
    ctx = manta_b200.Context(rank)

    def bcast(uid):
        box = [uid]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    g = shard.NcclGather(ctx, rank, world, "cuda:%d" % rank, bcast)
    ok = True
    for rnd, n in enumerate((1000 + 37 * rank, 0 if rank == 1 else 5, 3_000_000 + rank)):  # 3rd round outgrows the root buffer
        payload = (np.arange(n, dtype=np.uint32) * (rank + 1) + rnd).astype(np.uint32)
        got = g.gather([payload[:n // 2], payload[n // 2:]], dst=0)
        if rank == 0:
            for r in range(world):
                nr = (1000 + 37 * r, 0 if r == 1 else 5, 3_000_000 + r)[rnd]
                want = (np.arange(nr, dtype=np.uint32) * (r + 1) + rnd).astype(np.uint32)
                ok &= got[r].view(np.uint32).tolist() == want.tolist() if nr < 10000 else bool((got[r].view(np.uint32) == want).all())
    g.close()
    ctx.close()
    if rank == 0:
        ret.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_gather_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
