"""N>1 path on CPU: two gloo ranks shard 64 streams, run the scripted host state machine of every stream they own,
and combine timings exactly as bench.py does (max over ranks of the time, sum of the frames).  No GPU involved."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gstreamer_vit_tracker_b200 import sharding  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.streams_of_rank(64, rank, world)
        # every stream runs the reference's state machine (TrackerContext, scripted VitTrack results): independent per stream
        from gstreamer_vit_tracker_b200 import api
        states = []
        for sid in mine:
            ctx = api.TrackerContext.scripted(1920, 1080)
            ctx.handle_command(api.UserCommand.Confirm)
            ctx.process_scripted(None)
            ctx.handle_command(api.UserCommand.MoveRight, True)
            ctx.handle_command(api.UserCommand.MoveDown, True)
            ctx.handle_command(api.UserCommand.Confirm)
            ok = (sid % 3) != 0  # every third stream fails the acquisition gate (score <= 0.25, src/tracker_context.rs:93)
            ctx.process_scripted(api.TrackResult(True, 0.9 if ok else 0.1, (10 + sid, 20, 30, 40)))
            states.append((sid, ctx.state_name()))
        ms_local = [100.0 + 10.0 * rank, 50.0 - rank]
        tm = sharding.combine_timings(ms_local, frames_local=float(len(mine) * 10), launches_local=float(len(mine)))
        gathered = [None] * world
        dist.all_gather_object(gathered, states)
        if rank == 0:
            out.put((tm.ms_max, tm.frames, tm.launches, gathered))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_timing_combination():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    ms_max, frames, launches, gathered = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert ms_max == [110.0, 50.0]          # max over ranks, per leg
    assert frames == 640.0 and launches == 64.0  # sum over ranks
    seen = sorted(s for part in gathered for s, _ in part)
    assert seen == list(range(64))          # disjoint and complete
    for part_rank, part in enumerate(gathered):
        for sid, name in part:
            assert sharding.rank_of_stream(sid, world) == part_rank
            assert name == ("TRACKING" if sid % 3 else "SELECT START")
    assert sharding.whole_job_fps(sharding.JobTiming(ms_max, frames, launches), 0) == pytest.approx(640.0 / 0.110)


def test_single_process_fallback():
    tm = sharding.combine_timings([5.0], 10, 3)
    assert tm.ms_max == [5.0] and tm.frames == 10 and tm.launches == 3
    assert sharding.streams_of_rank(5, 1, 2) == [1, 3]
    with pytest.raises(ValueError):
        sharding.streams_of_rank(5, 2, 2)
