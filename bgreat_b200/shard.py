"""Multi-GPU plumbing of the mapping path: reads shard across ranks, the index is
replicated. One process per GPU; torch.distributed supplies the transport (NCCL over NVLink on the GPU
box, gloo in the CPU tests). There is no data-path collective: one broadcast of the flat index blob
before mapping, one all-reduce of the counters after it."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous slice [lo, hi) of n_items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_blob(blob, src=0, device=None):
    """Replicate a flat uint8 tensor (the index blob) from `src` to every rank.

    blob: uint8 tensor on `src` (ignored elsewhere). Returns the tensor on every rank. The size
    travels first so that receivers can allocate."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return blob
    rank = dist.get_rank()
    if device is None:
        device = blob.device if blob is not None else torch.device("cpu")
    size = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        size[0] = blob.numel()
    dist.broadcast(size, src)
    if rank != src:
        blob = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(blob, src)
    return blob


def replicate_index(aligner, build_fn, src=0):
    """Build the index once (build_fn() on rank `src`), broadcast the blob device-to-device, import it on
    the other ranks. `aligner` is a bgreat_b200.Aligner whose context lives on the current CUDA device."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == src:
        build_fn()
    if world == 1:
        return
    blob = None
    if rank == src:
        n = aligner.index_blob_size()
        blob = torch.empty(n, dtype=torch.uint8, device="cuda")
        aligner.export_index_device(blob.data_ptr(), n)
        torch.cuda.synchronize()
    blob = broadcast_blob(blob, src, device=torch.device("cuda"))
    if rank != src:
        aligner.import_index_device(blob.data_ptr(), blob.numel())
    del blob


def reduce_counters(counters, device=None):
    """Sum a dict of integer counters over all ranks (the reference's atomic<uint> totals, aligner.h:68)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(counters)
    keys = sorted(counters)
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(keys, t.tolist())}
