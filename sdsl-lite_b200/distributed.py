"""Multi-GPU plumbing for the query engine: replicated index, query batch sharded over ranks, optional
all-gather of the results.  One process per GPU, torch.distributed (NCCL on GPUs, gloo in
the CPU tests).  Queries are independent, so there is NO data-path collective while answering; the only
collective is the final all-gather of results, and only when the caller wants every rank to hold the whole
answer (north_star: "replicated index, NCCL all-gather of results only").
"""
import numpy as np


def shard_range(n, rank, world):
    """contiguous slice [lo, hi) of n queries owned by `rank`: sizes differ by at most one"""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def _gather_device(group=None):
    """where collective buffers must live: the current CUDA device under the NCCL backend, the host under gloo"""
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def sharded_query(fn, columns, group=None, gather=True):
    """Answer a batch split across the ranks of `group`.

    fn(*local_columns) -> torch tensor (or numpy array) of one result per local query;
    columns: torch tensors / numpy arrays of equal length n, identical on every rank.
    Returns the full result (n entries, original order) on every rank if gather, else this rank's slice.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(columns[0])
    lo, hi = shard_range(n, rank, world)
    local = fn(*[c[lo:hi] for c in columns])
    if not gather or world == 1:
        return local
    was_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local).view(np.int64)) if was_numpy else local
    dev = _gather_device(group)
    if t.device.type != dev.type:  # host results under NCCL (or device results under gloo): the collective decides
        t = t.to(dev)
    sizes = shard_sizes(n, world)
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=t.dtype, device=t.device)
    buf[: hi - lo] = t
    parts = [torch.empty(pad, dtype=t.dtype, device=t.device) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    full = torch.cat([p[:s] for p, s in zip(parts, sizes)])
    return full.cpu().numpy().view(np.uint64) if was_numpy else full


def _to_numpy_u64(x):
    if isinstance(x, np.ndarray):
        return x
    return x.cpu().numpy().view(np.uint64)


def sharded_locate(count_fn, locate_fn, flat, off, group=None):
    """locate over a sharded pattern set: per-rank (occ_off, occ) -> global CSR on every rank.
    count/locate work on patterns [lo, hi) given as (flat bytes, offsets rebased to 0)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(off) - 1
    lo, hi = shard_range(n, rank, world)
    b, e = int(off[lo]), int(off[hi])
    local_off = (off[lo : hi + 1] - off[lo]).astype(np.uint64)
    local_flat = flat[b:e] if e > b else np.zeros(1, np.uint8)
    occ_off, occ = locate_fn(local_flat, local_off)
    if world == 1:
        return occ_off, occ
    sizes = shard_sizes(n, world)
    # 1) all-gather the per-pattern counts (padded), 2) all-gather the occurrences (padded to the largest shard)
    dev = _gather_device(group)
    occ_off, occ = _to_numpy_u64(occ_off), _to_numpy_u64(occ)
    cnt = np.diff(occ_off.astype(np.int64))
    cpad = max(sizes)
    cbuf = torch.zeros(cpad, dtype=torch.int64, device=dev)
    cbuf[: hi - lo] = torch.from_numpy(cnt).to(dev)
    cparts = [torch.empty(cpad, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(cparts, cbuf, group=group)
    all_cnt = torch.cat([p[:s] for p, s in zip(cparts, sizes)]).cpu().numpy()
    totals = [int(p[:s].sum()) for p, s in zip(cparts, sizes)]
    opad = max(max(totals), 1)
    obuf = torch.zeros(opad, dtype=torch.int64, device=dev)
    obuf[: len(occ)] = torch.from_numpy(np.ascontiguousarray(occ).view(np.int64)).to(dev)
    oparts = [torch.empty(opad, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(oparts, obuf, group=group)
    full_occ = torch.cat([p[:t] for p, t in zip(oparts, totals)]).cpu().numpy().view(np.uint64)
    full_off = np.zeros(n + 1, dtype=np.uint64)
    full_off[1:] = np.cumsum(all_cnt).astype(np.uint64)
    return full_off, full_occ
