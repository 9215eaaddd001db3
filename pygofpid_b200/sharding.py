"""Stream sharding across GPUs / ranks.

Camera streams are independent (each owns its temporal background model), so the multi-GPU
scheme is pure data parallelism without any data-path collective: rank r owns a contiguous
block of streams, processes them on its own GPU, and the only cross-rank steps are the final
host gather of per-stream detections and, in bench.py, the max-over-ranks of the elapsed time.
"""


def streams_for_rank(n_streams, rank, world):
    """Half-open range [lo, hi) of the streams rank `rank` owns (block partition, sizes differ
    by at most one)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(local, dist=None, dst=0):
    """Final host gather: `local` maps stream index -> detections (any picklable object).
    Returns the merged dict on rank `dst` (None elsewhere); with dist=None it is the identity."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(dict(local), parts, dst=dst)
    if parts is None:
        return None
    merged = {}
    for p in parts:
        overlap = set(merged) & set(p)
        if overlap:
            raise ValueError("streams owned by two ranks: %s" % sorted(overlap))
        merged.update(p)
    return merged


def max_over_ranks(values, dist=None, device=None):
    """Element-wise maximum of a list of floats over all ranks (timing aggregation)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
