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


# ------------------------------------------------------------------------------------------------
# Final detection gather fused into the NMS kernel's sink: every rank owns one
# gather buffer (world x slot floats); rank r's detections live in slot r of EVERY buffer.  The
# fused head writes its outputs into its own buffer's slot and mirrors each store into the same
# slot of the peers' buffers over NVLink (VdHeadParams::mirror_delta), so when the step's kernels
# have finished the gathered result is already everywhere: no staging copy, no collective.
# ------------------------------------------------------------------------------------------------
def mirror_layout(world, rank, slot_bytes, bases):
    """Host-side address arithmetic of the peer gather (pure function, CPU-testable).

    bases[p] = address, in THIS process, of rank p's gather buffer (own buffer for p == rank).  Returns
    (own_slot_address, [byte deltas to add to an address inside the own slot to reach the same element in every peer
    buffer]) -- the deltas are what VdHeadParams::mirror_delta takes."""
    assert len(bases) == world and 0 <= rank < world and slot_bytes % 16 == 0
    own = bases[rank] + rank * slot_bytes
    deltas = [(bases[p] + rank * slot_bytes) - own for p in range(world) if p != rank]
    assert all(d % 16 == 0 for d in deltas), "gather buffers must be 16-byte aligned"
    return own, deltas


class _DevMem:
    """Raw device memory exposed through __cuda_array_interface__ so torch can view it (the buffer comes from the C ABI's
    cudaMalloc, not from torch's caching allocator: CUDA IPC handles must name whole allocations)."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (int(nfloats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerGather:
    """world x slot_floats fp32 gather buffer per rank, peer-mapped over CUDA IPC.

    `slot` (torch view, slot_floats) is this rank's slot of its own buffer: bind the head's outputs to slices of it;
    `deltas` goes into VdHeadParams::mirror_delta (HeadSession(mirrors=...)); `gathered` views the whole own buffer
    (world, slot_floats) -- complete once every rank's kernels have finished (stream sync + barrier)."""

    def __init__(self, slot_floats, group=None):
        import ctypes
        from . import _lib
        assert dist.is_initialized(), "PeerGather needs an initialised process group"
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert self.world - 1 <= _lib.VD_MAX_MIRRORS, "at most %d peers" % _lib.VD_MAX_MIRRORS
        self.slot_floats = (int(slot_floats) + 3) // 4 * 4              # 16-byte slots
        lib = _lib.load()
        self._lib, self._own, self._peers = lib, ctypes.c_void_p(), {}
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(lib.vd_ipc_alloc(self.world * self.slot_floats * 4, ctypes.byref(self._own), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        bases, err = [], None
        for p in range(self.world):
            if p == self.rank:
                bases.append(self._own.value)
                continue
            ptr = ctypes.c_void_p()
            h = (ctypes.c_ubyte * 64).from_buffer_copy(handles[p])
            rc = lib.vd_ipc_open(h, ctypes.byref(ptr))
            if rc != 0:
                err = lib.vd_last_error().decode("utf-8", "replace")
                break
            self._peers[p] = ptr
            bases.append(ptr.value)
        # every rank must take the same branch: agree on success before anyone relies on the mapping
        flag = torch.tensor([0 if err else 1], dtype=torch.int32, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            for ptr in self._peers.values():
                lib.vd_ipc_close(ptr)
            self._peers = {}
            dist.barrier(group=group)
            lib.vd_ipc_free(self._own)
            self._own = None
            raise RuntimeError("PeerGather: mapping a peer buffer failed on some rank (%s)" % (err or "another rank"))
        own_slot, self.deltas = mirror_layout(self.world, self.rank, self.slot_floats * 4, bases)
        self._mem = _DevMem(self._own.value, self.world * self.slot_floats)
        self.gathered = torch.as_tensor(self._mem, device="cuda").view(self.world, self.slot_floats)
        self.slot = self.gathered[self.rank]
        assert self.slot.data_ptr() == own_slot
        dist.barrier(group=group)                                        # every peer has mapped every buffer

    def close(self):
        for ptr in self._peers.values():
            self._lib.vd_ipc_close(ptr)
        self._peers = {}
        if self._own:
            dist.barrier()                                               # nobody still writes into a buffer about to be freed
            self._lib.vd_ipc_free(self._own)
            self._own = None
