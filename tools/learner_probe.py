#!/usr/bin/env python
"""Learner-only probe for profiling: BrainDQNNature, minibatch 256, replay pre-filled from random rollouts.

    python tools/learner_probe.py [--updates 20] [--envs 4096] [--batch 256] [--precision bf16] [--graph]

Prints ms/update measured with CUDA events (full _trainQNetwork, and loss_backward + Adam alone on a fixed minibatch).
Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from dqnflappybird_b200.brains import BrainDQNNature, BrainDoubleDQN  # noqa: E402
from dqnflappybird_b200.game import GameState  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--updates", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--double", action="store_true")
    ap.add_argument("--max-act-batch", type=int, default=2048)
    a = ap.parse_args()
    dev = "cuda:0"
    N, B, C = a.envs, a.batch, 28
    cls = BrainDoubleDQN if a.double else BrainDQNNature
    brain = cls(2, "bird", num_envs=N, device=dev, replay_memory_per_env=C, batch_size=B, observe=1e18, seed=0,
                max_act_batch=a.max_act_batch, precision=a.precision)
    gs = GameState(num_envs=N, device=dev, seed=42, history=C + 4, ring=brain.ring)
    obs, *_ = gs.frame_step(torch.zeros(N, dtype=torch.uint8, device=dev))
    brain.setInitState(obs)
    for k in range(1, C + 9):
        a_row, r_row, t_row = brain.replayMemory.rows(k)
        gs.step_random(1, 0.5, 1234, a_row, r_row, t_row, None)
        brain._k = k
        brain.replayMemory.appended(k)
    brain.timeStep = 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(a.warmup):
        brain._trainQNetwork(); brain.timeStep += 1
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.updates):
        brain._trainQNetwork(); brain.timeStep += 1
    e1.record(); torch.cuda.synchronize()
    ms_full = e0.elapsed_time(e1) / a.updates
    mb = brain.replayMemory.sample(B)
    net = brain.net
    torch.cuda.synchronize()
    import time
    e0.record()
    t0 = time.perf_counter()
    for _ in range(a.updates):
        net.loss_backward(brain.variant, mb.frames, mb.actions, mb.rewards, mb.terminals)
        net.adam_step()
    host_issue_ms = (time.perf_counter() - t0) * 1e3 / a.updates      # host time to ISSUE an update (no sync): CPU-bound if ~ ms_net
    e1.record(); torch.cuda.synchronize()
    ms_net = e0.elapsed_time(e1) / a.updates
    e0.record()
    for _ in range(a.updates):
        net.loss_backward(brain.variant, mb.frames, mb.actions, mb.rewards, mb.terminals)
    e1.record(); torch.cuda.synchronize()
    ms_lb = e0.elapsed_time(e1) / a.updates
    for _ in range(3):
        net.train_step(brain.variant, mb.frames, mb.actions, mb.rewards, mb.terminals)
    e0.record()
    for _ in range(a.updates):
        net.train_step(brain.variant, mb.frames, mb.actions, mb.rewards, mb.terminals)
    e1.record(); torch.cuda.synchronize()
    ms_fused = e0.elapsed_time(e1) / a.updates
    e0.record()
    for _ in range(a.updates):
        brain.getAction()
    e1.record(); torch.cuda.synchronize()
    ms_act = e0.elapsed_time(e1) / a.updates
    print(json.dumps({"ms_per_update_full": ms_full, "ms_per_update_net_only": ms_net, "ms_loss_backward_only": ms_lb, "ms_train_step_fused": ms_fused, "host_issue_ms_per_update": host_issue_ms, "updates_per_s": 1e3 / ms_full,
                      "ms_per_act": ms_act, "act_envs_per_s": N / (ms_act * 1e-3), "envs": N, "batch": B,
                      "precision": a.precision, "variant": brain.variant}))


if __name__ == "__main__":
    main()
