#!/usr/bin/env python
"""Kernel timeline (torch.profiler / CUPTI) of loss_backward + Adam on a fixed minibatch of any size."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from dqnflappybird_b200 import qnet  # noqa: E402
from dqnflappybird_b200.game import GameState  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
gs = GameState(num_envs=B, seed=3, history=5)
gs.step_random(60, 0.3, 7)
order = [(gs.slot - 4 + k) % 5 for k in range(5)]
frames = gs.ring[:, order].contiguous()
net = qnet.QNetwork(max_batch=B, precision="bf16")
act = torch.randint(0, 2, (B,), dtype=torch.uint8, device="cuda")
rew = torch.full((B,), 0.1, device="cuda"); term = torch.zeros(B, dtype=torch.uint8, device="cuda")
for _ in range(5):
    net.loss_backward("nature", frames, act, rew, term); net.adam_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        net.loss_backward("nature", frames, act, rew, term); net.adam_step()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
n = len(evs) // 3
for e in evs[-n:]:
    print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:78]}")
print("span of the last update us", evs[-1].time_range.end - evs[-n].time_range.start)
