#!/usr/bin/env python
"""How long do the two halves of the step kernel take on their own?  physics only (frame_step without drawing) and drawing only
(fb_env_draw: the current state of every env), 131,072 envs, CUDA events."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from dqnflappybird_b200 import _lib  # noqa: E402
from dqnflappybird_b200.game import GameState, _stream_ptr  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
gs = GameState(num_envs=E, seed=42, history=4)
rew = torch.empty(E, dtype=torch.float32, device="cuda"); term = torch.empty(E, dtype=torch.uint8, device="cuda")
score = torch.empty(E, dtype=torch.int32, device="cuda")
gs.step_random(1000, 0.5, 1234, draw=False)


def t(fn, reps=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


out = {"envs": E}
out["step_and_draw_us"] = t(lambda: gs.step_random(1, 0.5, 1234, None, rew, term, score))
out["physics_only_us"] = t(lambda: gs.step_random(1, 0.5, 1234, None, rew, term, score, draw=False))
out["draw_only_us"] = t(lambda: _lib.check(gs._L.fb_env_draw(gs._h, gs.ring.data_ptr(), gs.history, 0, _stream_ptr(gs.device)), "draw"))
a = (torch.rand(E, device="cuda") < 0.5).to(torch.uint8)
out["step_and_draw_given_actions_us"] = t(lambda: gs.frame_step(a))
print(json.dumps(out))
