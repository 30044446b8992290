"""Trajectory-index sharding across the GPUs of one box (one process per GPU).

The reference's only parallelism is Rayon `par_iter_traj_id` over independent trajectories
(obs_dataset_api.rs:191-206): no data-path collective exists, so the multi-GPU design is the same
decomposition -- contiguous trajectory ranges per rank, balanced by a work estimate -- followed by
ONE gather of the per-trajectory results (the only cross-GPU step).
"""
import numpy as np


def work_estimate(n_obs, max_triplets, n_noise):
    """Relative cost of a trajectory: candidates x (Gauss solve + arc scoring)."""
    n = np.asarray(n_obs, dtype=np.float64)
    feasible = n * (n - 1) * (n - 2) / 6.0
    k = np.minimum(float(max_triplets), feasible)
    return k * (1.0 + n_noise) * (60.0 + n) + 0.05 * feasible + 1.0


def shard_ranges(traj_offset, world_size, max_triplets=10, n_noise=20):
    """Split [0, T) into `world_size` contiguous ranges with near-equal estimated work.

    Returns a list of (t_begin, t_end) per rank (possibly empty ranges when T < world_size).
    """
    off = np.asarray(traj_offset, dtype=np.int64)
    T = len(off) - 1
    if T <= 0:
        return [(0, 0)] * world_size
    cost = np.cumsum(work_estimate(np.diff(off), max_triplets, n_noise))
    total = cost[-1]
    cuts = [0]
    for r in range(1, world_size):
        cuts.append(int(np.searchsorted(cost, total * r / world_size, side="left")))
    cuts.append(T)
    cuts = np.maximum.accumulate(np.minimum(cuts, T))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world_size)]


def slice_batch(batch, t_begin, t_end):
    """Host-side view of the trajectories [t_begin, t_end) of a synth/host batch (re-based offsets)."""
    off = np.asarray(batch["traj_offset"], dtype=np.int64)
    o0, o1 = int(off[t_begin]), int(off[t_end])
    out = {"traj_offset": np.ascontiguousarray((off[t_begin:t_end + 1] - o0).astype(np.uint64))}
    for k in ("mjd_tt", "ra", "dec", "sigma_ra", "sigma_dec", "mjd_ut1"):
        if batch.get(k) is not None:
            out[k] = np.ascontiguousarray(batch[k][o0:o1])
    for k in ("helio_equ", "geo_ecl", "body_fixed"):
        if batch.get(k) is not None:
            out[k] = np.ascontiguousarray(batch[k][:, o0:o1])
    for k in ("noise_z", "traj_seed"):  # per-trajectory inputs travel with their shard
        if batch.get(k) is not None:
            out[k] = np.ascontiguousarray(batch[k][t_begin:t_end])
    if t_end > t_begin:
        out["max_obs_per_traj"] = int(np.diff(off[t_begin:t_end + 1]).max())
    for k in ("table", "max_triplets", "n_noise"):
        if k in batch:
            out[k] = batch[k]
    return out


def gather_results(local, ranges, rank, world_size, dist=None, device=None):
    """Gather per-rank result arrays (numpy structured, outfit_b200.RESULT_DTYPE) on rank 0.

    Uses torch.distributed (`gloo` on CPU, `nccl` on GPU) when world_size > 1; the payload is the
    raw bytes of the fixed-size result records, padded to the largest shard.
    """
    if world_size == 1 or dist is None:
        return local
    import torch
    itemsize = local.dtype.itemsize
    n_max = max(e - b for b, e in ranges)
    buf = torch.zeros(n_max * itemsize, dtype=torch.uint8, device=device)
    if len(local):
        raw = torch.from_numpy(np.frombuffer(local.tobytes(), dtype=np.uint8).copy())
        buf[: raw.numel()] = raw.to(buf.device)
    bufs = [torch.zeros_like(buf) for _ in range(world_size)]
    dist.all_gather(bufs, buf)
    if rank != 0:
        return None
    # join the raw records (np.concatenate on aligned structured dtypes does not preserve padding bytes)
    raw = b"".join(bufs[r][: (e - b) * itemsize].cpu().numpy().tobytes() for r, (b, e) in enumerate(ranges))
    return np.frombuffer(raw, dtype=local.dtype)
