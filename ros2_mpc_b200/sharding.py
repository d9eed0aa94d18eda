"""Multi-GPU host logic: independent problems shard by batch index, one process per GPU, no collective on the
solve path.  The only communication is the final host-side gather of the per-rank output
slices, done with torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def contiguous_shard(B, rank, world):
    """[lo, hi) of rank's contiguous slice of a batch of B problems (sizes differ by at most one)."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def solve_sharded(solve_fn, arrays, B, rank, world, dist=None, device=None, shared=()):
    """Each rank calls solve_fn(slice of every array) on its shard; rank 0 receives the concatenated outputs.

    solve_fn maps a dict of input arrays to a dict of output arrays with the shard's batch on axis 0.  Every numpy array
    in `arrays` is batch-shaped (leading dimension B) and is sliced, EXCEPT the keys named in `shared` (one obstacle list
    for all problems, tables, ...) and None entries, which are passed through whole.  A batched array whose leading
    dimension is not B is an error — nothing is inferred from shapes.  Returns the full-batch dict on rank 0 and None
    elsewhere.  With world == 1 no process group is needed."""
    lo, hi = contiguous_shard(B, rank, world)
    local_in = {}
    for k, v in arrays.items():
        if v is None or k in shared or not isinstance(v, np.ndarray):
            local_in[k] = v
            continue
        if v.ndim < 1 or v.shape[0] != B:
            raise ValueError(f"array {k!r} has leading dimension {v.shape[:1]}, expected the batch size {B} "
                             "(name it in `shared` if it is not per problem)")
        local_in[k] = v[lo:hi]
    local_out = solve_fn(local_in)
    if world == 1:
        return local_out
    import torch  # noqa: PLC0415
    sizes = [contiguous_shard(B, r, world) for r in range(world)]
    nmax = max(h - l for l, h in sizes)
    result = {}
    for key in sorted(local_out):
        a = np.ascontiguousarray(local_out[key])
        # dist.gather needs equally sized tensors: pad the shard to the largest shard and trim after the gather
        pad = np.zeros((nmax,) + a.shape[1:], dtype=a.dtype)
        pad[:a.shape[0]] = a
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, gather_list=bufs, dst=0)
        if rank == 0:
            result[key] = np.concatenate([bufs[r][:h - l].cpu().numpy() for r, (l, h) in enumerate(sizes)], axis=0)
    return result if rank == 0 else None
