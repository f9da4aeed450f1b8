#!/usr/bin/env python
"""Kernel timeline of training updates on several GPUs (torchrun): rank 0 prints start, duration and name of every kernel of
three updates -- shows how long the exchange + Adam kernel waits for its peers.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tools/timeline_dist_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from dqnflappybird_b200 import dist as fdist  # noqa: E402
from dqnflappybird_b200.brains import BrainDQNNature  # noqa: E402
from dqnflappybird_b200.game import GameState  # noqa: E402


def main():
    rank, world, local = fdist.init()
    dev = f"cuda:{local}"
    N, B, C = 4096, 256, 28
    brain = BrainDQNNature(2, "bird", num_envs=N, device=dev, replay_memory_per_env=C, batch_size=B * world, observe=1e18, seed=0, first_env_id=rank * N)
    gs = GameState(num_envs=N, device=dev, seed=42, history=C + 4, ring=brain.ring, first_env_id=rank * N)
    obs, *_ = gs.frame_step(torch.zeros(N, dtype=torch.uint8, device=dev))
    brain.setInitState(obs)
    for k in range(1, C + 9):
        a_row, r_row, t_row = brain.replayMemory.rows(k)
        gs.step_random(1, 0.5, 1234, a_row, r_row, t_row, None)
        brain._k = k
        brain.replayMemory.appended(k)
    brain.timeStep = 1
    for _ in range(30):
        brain._trainQNetwork(); brain.timeStep += 1
    torch.cuda.synchronize()
    torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        brain._trainQNetwork(); brain.timeStep += 1
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"world {world}: {e0.elapsed_time(e1) * 10:.1f} us per update")
    # phase stamps of the exchange kernel (block 0 of every rank): wait for peers' gradients, own slice, slice barrier, Adam
    if brain.net.exchange is not None:
        import ctypes as C
        from dqnflappybird_b200 import _lib
        L = _lib.lib()
        buf = (C.c_ulonglong * 5)()
        L.fb_dist_debug_stamps(brain.net.exchange._h, buf)
        rows = []
        for _ in range(6):
            for _ in range(40):
                brain._trainQNetwork(); brain.timeStep += 1
            L.fb_dist_debug_stamps(brain.net.exchange._h, buf)
            t = [int(v) for v in buf]
            rows.append([(t[k + 1] - t[k]) / 1e3 for k in range(4)])
        med = [sorted(r[k] for r in rows)[len(rows) // 2] for k in range(4)]
        out = torch.tensor(med, device=dev)
        allm = [torch.empty_like(out) for _ in range(world)]
        torch.distributed.all_gather(allm, out)
        if rank == 0:
            print("exchange kernel phases, us (median of 6 samples): wait for gradients | reduce own slice | fence + grid barrier + slice flags | gather + Adam")
            for r, m in enumerate(allm):
                print(f"  rank {r}: " + "  ".join(f"{v:7.1f}" for v in m.tolist()))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            brain._trainQNetwork(); brain.timeStep += 1
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        for e in evs:
            print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:70]}")
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
