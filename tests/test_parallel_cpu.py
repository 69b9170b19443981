"""World-size-2 gloo test (CPU) of the multi-GPU host logic: image sharding + the single all_gather of restored
shards.  The per-rank compute is stood in for by a deterministic function of the image index (the CUDA path is
covered by the -m gpu tests); what is checked is that every rank ends with the full batch in image order, for
even and ragged batch sizes."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from daclip_b200.parallel import gather_restored, shard_range


def test_shard_range_partitions():
    for n in (1, 2, 7, 16, 128, 129):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_images, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_range(n_images, rank, world)
        idx = torch.arange(b, e, dtype=torch.float32)
        local = idx[:, None, None, None] * torch.ones(1, 3, 4, 5) + 0.25      # "restored" shard of this rank
        full = gather_restored(local, n_images)
        expect = torch.arange(n_images, dtype=torch.float32)[:, None, None, None] * torch.ones(1, 3, 4, 5) + 0.25
        out[rank] = bool(torch.equal(full, expect))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [8, 7])
def test_gather_restored_gloo_world2(n_images):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out[0] and out[1]
