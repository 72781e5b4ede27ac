"""Multi-GPU plumbing: one process per GPU, QPs sharded by contiguous index range, no collective
inside the solve.  torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests) only
gathers the per-rank results and reduces a handful of statistics."""
import numpy as np


def shard_range(batch, rank, world):
    """Contiguous range [lo, hi) of rank `rank`: sizes differ by at most one, union = [0, batch)."""
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local, batch, group=None, dst=None):
    """All-gather (dst=None) or gather-to-dst of a dict of per-QP arrays sharded with shard_range.

    `local` maps name -> torch tensor whose first dimension is the local shard.  Shards may differ in
    length by one, so they are padded to the longest before the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    out = {}
    for name, t in local.items():
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        if dst is None:
            bufs = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(bufs, pad, group=group)
        else:
            bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
            dist.gather(pad, bufs, dst=dst, group=group)
        if bufs is not None:
            out[name] = torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
    return out


_PACK_CACHE = {}


def gather_packed(local, batch, group=None, dst=None, to_host=False):
    """gather_results with ONE collective: the per-QP result columns (first input, objective, iterations, status ... any
    1-D / (n,k) arrays; int32 columns are exact in float64) are packed into one float64 matrix per rank, gathered once and
    unpacked with their dtypes.  At a few KB per rank the cost of a gather is launch latency, so what matters is the NUMBER of
    small operations around it (NVLink bandwidth is irrelevant here): the send matrix and the receive block are
    allocated once per shape and reused, the columns are written into the send matrix by converting copies (no concatenation),
    and the receive block is gathered into views of one tensor (no concatenation on the receiving side when the shards have
    equal length).  to_host=True: the receiving rank copies the gathered block to the host in ONE device-to-host copy (into a
    page-locked buffer that is reused) and returns numpy arrays — what a caller that wants the results on the host should use
    instead of one `.cpu()` per column."""
    import torch
    import torch.distributed as dist
    names = list(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    widths = [int(local[n].numel() // max(1, local[n].shape[0])) if local[n].shape[0] else int(np.prod(local[n].shape[1:], dtype=np.int64)) for n in names]
    w = sum(widths)
    dev = local[names[0]].device
    key = (world, mx, w, str(dev), dst, id(group))
    if key not in _PACK_CACHE:
        send = torch.zeros((mx, w), dtype=torch.float64, device=dev)
        recv = torch.empty((world, mx, w), dtype=torch.float64, device=dev) if (dst is None or rank == dst) else None
        _PACK_CACHE.clear()          # one shape at a time is all a loop of identical steps needs (host block included)
        _PACK_CACHE[key] = (send, recv)
    send, recv = _PACK_CACHE[key]
    n_loc, o = sizes[rank], 0
    for n, wd in zip(names, widths):
        send[:n_loc, o:o + wd].copy_(local[n].reshape(n_loc, wd))
        o += wd
    if dst is None:
        dist.all_gather(list(recv.unbind(0)), send, group=group)
    else:
        dist.gather(send, list(recv.unbind(0)) if rank == dst else None, dst=dst, group=group)
    if recv is None:
        return {}
    full = recv.reshape(world * mx, w) if min(sizes) == mx else torch.cat([recv[r, :n] for r, n in enumerate(sizes)], dim=0)
    if to_host:
        hk = ("host",) + key
        if hk not in _PACK_CACHE or _PACK_CACHE[hk].shape[0] < full.shape[0]:
            pin = dev.type == "cuda"
            _PACK_CACHE[hk] = torch.empty((world * mx, w), dtype=torch.float64, pin_memory=pin)
        host = _PACK_CACHE[hk][: full.shape[0]]
        host.copy_(full)                      # one D2H (synchronous for the caller: the results are on the host when it returns)
        arr, out, o = host.numpy(), {}, 0
        for n, wd in zip(names, widths):
            dt = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64}[local[n].dtype]
            out[n] = arr[:, o:o + wd].astype(dt).reshape((arr.shape[0],) + tuple(local[n].shape[1:]))
            o += wd
        return out
    out, o = {}, 0
    for n, wd in zip(names, widths):
        part = full[:, o:o + wd].to(local[n].dtype, copy=True)   # never a view of the reused receive block
        out[n] = part.reshape((part.shape[0],) + tuple(local[n].shape[1:]))
        o += wd
    return out


def reduce_stats(status, iters, obj, group=None):
    """{n_optimal, n_maxiter, n_infeasible, n_numerical, sum_iters, max_iters, sum_obj_optimal} over all ranks."""
    import torch
    import torch.distributed as dist
    st = status.to(torch.int64)
    sums = torch.stack([(st == k).sum() for k in range(4)] + [iters.to(torch.int64).sum()]).to(torch.float64)
    sums = torch.cat([sums, torch.where(st == 0, obj, torch.zeros_like(obj)).sum().reshape(1)])
    mx = iters.max().to(torch.float64).reshape(1) if iters.numel() else torch.zeros(1, dtype=torch.float64, device=obj.device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    s = sums.tolist()
    return dict(n_optimal=int(s[0]), n_maxiter=int(s[1]), n_infeasible=int(s[2]), n_numerical=int(s[3]),
                sum_iters=int(s[4]), max_iters=int(mx.item()), sum_obj_optimal=s[5])


def sample_initial_states(batch, seed, lo=(-0.40, -0.45, -0.05, -1.0), hi=(0.10, 0.10, 0.05, 1.0), first=None):
    """i.i.d. uniform initial conditions of the bench configs; element 0 is the
    reference's canonical start [-0.35;-0.4;0;0] (LBMPC_RunExample.m:41-44).  Deterministic in the GLOBAL
    index, so a shard is a slice of the same array regardless of the world size."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    x = lo + (hi - lo) * rng.random((int(batch), lo.size))
    if batch > 0:
        x[0] = [-0.35, -0.4, 0.0, 0.0] if first is None else first
    return x
