#!/usr/bin/env python
"""Kernel timeline of a few training updates through torch.profiler (CUPTI): start offset, duration, stream, name."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from dqnflappybird_b200.brains import MODELS  # noqa: E402
from dqnflappybird_b200.game import GameState  # noqa: E402


def main():
    dev = "cuda:0"
    N, B, C = 4096, 256, 28
    brain = MODELS[sys.argv[1] if len(sys.argv) > 1 else "dqnnature"](2, "bird", num_envs=N, device=dev, replay_memory_per_env=C, batch_size=B, observe=1e18, seed=0, max_act_batch=2048)
    gs = GameState(num_envs=N, device=dev, seed=42, history=C + 4, ring=brain.ring)
    obs, *_ = gs.frame_step(torch.zeros(N, dtype=torch.uint8, device=dev))
    brain.setInitState(obs)
    for k in range(1, C + 9):
        a_row, r_row, t_row = brain.replayMemory.rows(k)
        gs.step_random(1, 0.5, 1234, a_row, r_row, t_row, None)
        brain._k = k
        brain.replayMemory.appended(k)
    brain.timeStep = 1
    for _ in range(6):
        brain._trainQNetwork(); brain.timeStep += 1
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(4):
            brain._trainQNetwork(); brain.timeStep += 1
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    last_end = {}
    for e in evs:
        name = e.name[:70]
        print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:7.1f} us  {name}")
    print("total span us", evs[-1].time_range.end - t0)


if __name__ == "__main__":
    main()
