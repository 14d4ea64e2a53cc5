"""Multi-GPU plumbing for the head path: frames / clips / training images are independent, so the
batch is sharded across ranks with no data-path collective (the reference shards the same way with
gluon.utils.split_and_load, detect_yolo3.py:211-213); the only exchange is the final gather of the
packed (frames, post_nms, 6) detections.  One process per GPU, torch.distributed (NCCL on GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns
    (rank, world_size, local_rank); a no-op single-process setup when WORLD_SIZE is absent or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous split of n_items over `world` ranks (split_and_load(even_split=False) semantics:
    the first n % world ranks get one extra item).  Returns (begin, end)."""
    base, rem = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_clips(clip_lengths, rank, world):
    """Round-robin assignment of whole clips to ranks (windows never cross ranks; edge clamping is
    per clip, datasets/imgnetvid.py:480-506).  Returns the clip indices owned by `rank`."""
    return [i for i in range(len(clip_lengths)) if i % world == rank]


def gather_detections(packed_local, counts=None, out=None):
    """All-gather of the per-rank packed detections (frames_local, post_nms, 6) -> (sum frames, post_nms, 6).

    With equal shards this is one all_gather_into_tensor; ragged shards (counts = frames per rank)
    are padded to the largest shard and trimmed after the exchange."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return packed_local
    world = dist.get_world_size()
    if counts is None:
        counts = [packed_local.shape[0]] * world
    fmax = max(counts)
    send = packed_local
    if packed_local.shape[0] != fmax:
        send = packed_local.new_full((fmax,) + tuple(packed_local.shape[1:]), -1.0)
        send[:packed_local.shape[0]] = packed_local
    send = send.contiguous()
    if out is None:
        out = send.new_empty((world * fmax,) + tuple(send.shape[1:]))
    dist.all_gather_into_tensor(out, send)
    if all(c == fmax for c in counts):
        return out
    parts = [out[r * fmax:r * fmax + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
