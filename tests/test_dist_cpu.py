"""world_size-2 gloo test (CPU) of the sharding + gather plumbing used by the multi-GPU path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from viddet_b200 import dist as vdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = vdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    b, e = vdist.shard_range(total, rank, world)
    full = torch.arange(total * 4 * 6, dtype=torch.float32).reshape(total, 4, 6)
    counts = [vdist.shard_range(total, i, world)[1] - vdist.shard_range(total, i, world)[0] for i in range(world)]
    got = vdist.gather_detections(full[b:e].clone(), counts=counts)
    q.put((rank, torch.equal(got, full), tuple(got.shape)))
    dist.barrier()
    dist.destroy_process_group()


def _run(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_gather_equal_and_ragged_shards():
    for total in (8, 7):
        for rank, ok, shape in _run(total):
            assert ok and shape == (total, 4, 6)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 129):
        for w in (1, 2, 4, 8):
            spans = [vdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert vdist.shard_clips([10, 20, 30, 40, 50], 1, 2) == [1, 3]


def test_mirror_layout_addresses():
    """Host-side address arithmetic of the fused detection gather (PeerGather): rank r's slot sits at r * slot_bytes in every
    buffer; the deltas lead from the own slot to the same slot of every peer buffer, in rank order without the own rank."""
    bases = [0x10000000, 0x7f0000000000, 0x20000100]
    for rank in range(3):
        own, deltas = vdist.mirror_layout(3, rank, 4096, bases)
        assert own == bases[rank] + rank * 4096
        peers = [p for p in range(3) if p != rank]
        assert [own + d for d in deltas] == [bases[p] + rank * 4096 for p in peers]
    import pytest
    with pytest.raises(AssertionError):
        vdist.mirror_layout(2, 0, 4096, [0x1000, 0x2004])        # misaligned peer buffer
