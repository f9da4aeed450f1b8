#!/usr/bin/env python
"""Per-kernel time of ONE training step at a large minibatch (default 4096) through torch.profiler: where the throughput goes."""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from dqnflappybird_b200.game import GameState  # noqa: E402
from dqnflappybird_b200.qnet import QNetwork  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    gs = GameState(num_envs=B, seed=3, history=5)
    gs.step_random(45, 0.3, 11)
    order = [(gs.slot - 4 + k) % 5 for k in range(5)]
    frames = gs.ring[:, order].contiguous()
    a = (torch.rand(B, device="cuda") < 0.5).to(torch.uint8)
    r = torch.full((B,), 0.1, device="cuda")
    term = torch.zeros(B, dtype=torch.uint8, device="cuda")
    net = QNetwork(max_batch=B, precision=prec, seed=0)
    for _ in range(4):
        net.train_step("nature", frames, a, r, term)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        net.train_step("nature", frames, a, r, term)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flop = (2 * 11675648 + 16797696) * B
    print(f"B={B} {prec}: {ms * 1e3:.1f} us / update, {flop / (ms * 1e-3) / 1e12:.1f} TFLOP/s")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            net.train_step("nature", frames, a, r, term)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    half = evs[len(evs) // 2:]
    t0 = half[0].time_range.start
    tot = defaultdict(float)
    for e in half:
        print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:90]}")
        tot[e.name[:60]] += e.time_range.end - e.time_range.start
    print("span us", half[-1].time_range.end - t0)


if __name__ == "__main__":
    main()
