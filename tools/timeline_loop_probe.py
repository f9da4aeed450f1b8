#!/usr/bin/env python
"""Kernel timeline of the closed loop (getAction -> frame_step -> setPerception) through torch.profiler: per kernel
start / duration, gaps on the device, and the CPU time of each step."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from dqnflappybird_b200.brains import BrainDQNNature  # noqa: E402
from dqnflappybird_b200.game import GameState  # noqa: E402


def main():
    dev = "cuda:0"
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    observe = float(sys.argv[2]) if len(sys.argv) > 2 else 1e18
    brain = BrainDQNNature(2, "bird", num_envs=N, device=dev, replay_memory_per_env=28, batch_size=256, observe=observe, seed=0)
    gs = GameState(num_envs=N, device=dev, seed=42, history=32, ring=brain.ring)
    obs, *_ = gs.frame_step(torch.zeros(N, dtype=torch.uint8, device=dev))
    brain.setInitState(obs)

    def step():
        a = brain.getAction()
        o, r, t, s = gs.frame_step(a, out=brain.next_rows()[1:])
        brain.setPerception(o, a, r, t, s)

    for _ in range(40):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        step()
    cpu_issue = (time.perf_counter() - t0) / 50
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 50
    print(f"envs {N}: host issue {cpu_issue * 1e6:.0f} us/step, wall {wall * 1e6:.0f} us/step")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    end = 0
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = s - end
        print(f"{s:9.1f} us  +{d:7.1f} us  {'GAP %5.1f' % gap if gap > 3 else '         '}  {e.name[:60]}")
        end = max(end, s + d)
    print("total span us", end)


if __name__ == "__main__":
    main()
