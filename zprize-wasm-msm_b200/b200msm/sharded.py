"""Point-range sharding of one MSM over the ranks of a torch.distributed job.

MSM is a sum over independent (point, scalar) pairs: rank g owns the contiguous slice [lo_g, hi_g) of the points,
runs the complete single-GPU pipeline on it and produces ONE partial G1 point (3*n8 bytes).  The only exchange is
an all_gather of those world_size partials (NCCL over NVLink on GPUs; gloo in the CPU tests), after which every
rank sums them (g1m_add chain) -- the same combination ffjavascript's worker pool does with the results of
g1m_multiexpAffine_chunk / slices of g1m_multiexpAffine (wasmcurves/src/build_multiexp.js:319-369).

The module is pure plumbing: `local_msm` and `combine` are injected (the GPU engine in production, the oracle in the
CPU tests), so the partition / gather logic is testable without a GPU.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced partition of [0, n): the first n % world ranks get one extra point."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_partials(partial: bytes, device=None, group=None):
    """all_gather of one fixed-size byte string per rank -> list of world_size byte strings (rank order)."""
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(partial), dtype=torch.uint8)
    if device is not None: t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    raw = bytes(out.cpu().numpy())
    k = len(partial)
    return [raw[i * k:(i + 1) * k] for i in range(world)]


def sharded_msm(local_msm, combine, bases: bytes, scalars: bytes, scalar_size: int, n: int, point_bytes: int, device=None, group=None):
    """local_msm(bases_slice, scalars_slice, scalar_size, count) -> partial (Jacobian bytes);
    combine(list_of_partials) -> result bytes.  `bases` / `scalars` are the FULL inputs (each rank slices its own range);
    returns the same result on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n, rank, world)
    part = local_msm(bases[lo * point_bytes: hi * point_bytes], scalars[lo * scalar_size: hi * scalar_size], scalar_size, hi - lo)
    return combine(gather_partials(part, device=device, group=group))


def engine_sharded_msm(engine, curve, bases, scalars, scalar_size, n, device=None, group=None):
    """Production form: the per-rank MSM and the final sum both run on this rank's GPU through the C ABI."""
    from ._lib import N8
    return sharded_msm(lambda b, s, ss, m: engine.multiexp_affine(curve, b, s, ss, m),
                       lambda parts: engine.sum_points(curve, b"".join(parts), len(parts)),
                       bases, scalars, scalar_size, n, 2 * N8[curve], device=device, group=group)
