"""Frame sharding across the GPUs of one box (BASELINE.json configs[4]).

Frames are independent: ORBextractor::operator() keeps no cross-frame state (ORBextractor.cc:1060-1141), so the batch is
split by frame index with NO data-path collective.  torch.distributed is used only by the benchmark / callers for the
start barrier, the max-over-ranks timing and an optional host-side gather of small per-frame summaries."""


def shard_range(n_frames, rank, world):
    """Contiguous, balanced block of frame indices of `rank`: sizes differ by at most one, lower ranks get the extra."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def weak_scaling_frames(pairs_per_rank, rank):
    """Global frame ids of a weak-scaling run: every rank owns `pairs_per_rank` distinct frames."""
    return rank * pairs_per_rank, (rank + 1) * pairs_per_rank


def frame_seed(frame_id):
    """orbx-synth-v1 seed of a global frame id (left and right image share it)."""
    return frame_id & 0xFFFFFFFF


def gather_frame_summaries(local, dist=None):
    """All ranks' per-frame summaries (any picklable list) concatenated in rank order on every rank; no-op without a group."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, list(local))
    return [x for p in parts for x in p]
