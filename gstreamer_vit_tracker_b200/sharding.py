"""Stream-level data parallelism across the GPUs of one box.

Video streams are independent (the reference creates one TrackerContext per pipeline,
/root/reference/src/pipeline.rs:55) and the path has no exchange step, so the only multi-GPU
"strategy" is: stream i -> rank i mod world, one process per GPU, NO data-path collective.
torch.distributed (nccl on GPUs, gloo in the CPU tests) is used for the start/stop barrier and
for combining the per-rank timings: whole-job value = frames of all ranks / max-over-ranks time.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence


def streams_of_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Global stream ids owned by `rank` (round-robin: stream i lives on rank i % world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_streams, world))


def rank_of_stream(stream: int, world: int) -> int:
    return stream % world


@dataclass
class JobTiming:
    ms_max: List[float]      # per timed leg: max over ranks of the device time (ms)
    frames: float            # frames processed by all ranks
    launches: float          # kernels launched by all ranks


def combine_timings(ms_local: Sequence[float], frames_local: float, launches_local: float, device=None) -> JobTiming:
    """MAX over ranks of every leg time, SUM of frames / launches.  Works with or without an initialised process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return JobTiming(list(map(float, ms_local)), float(frames_local), float(launches_local))
    t = torch.tensor(list(ms_local), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor([frames_local, launches_local], dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return JobTiming([float(x) for x in t], float(c[0]), float(c[1]))


def whole_job_fps(t: JobTiming, leg: int = 0) -> float:
    return t.frames / (t.ms_max[leg] * 1e-3) if t.ms_max[leg] > 0 else 0.0
