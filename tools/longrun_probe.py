import sys, time, torch
sys.path.insert(0, "/root/repo")
from dqnflappybird_b200 import play
for model in ("dqnnature", "prioritydqn", "ddqn"):
    free0, _ = torch.cuda.mem_get_info()
    t0 = time.time()
    brain, gs, stats = play.playFlappyBird(model, num_envs=4096, steps=6000, replay_memory_per_env=60, batch_size=256, observe=100, lr=1e-5,
                                           record=True, log_capacity=1 << 22)
    torch.cuda.synchronize()
    brain.flush_logs()
    free1, _ = torch.cuda.mem_get_info()
    import numpy as np
    lh = np.array(brain.lost_hist)
    print(model, stats, "time %.1fs" % (time.time() - t0), "finite", bool(torch.isfinite(brain.net.params).all()), "loss first/last 100 mean %.4f %.4f" % (lh[:100].mean(), lh[-100:].mean()),
          "episodes", len(brain.score_every_episode), "max score", max(brain.score_every_episode), "mean reward last 100 steps %.4f" % np.mean(brain.reward_every_time_step[-100:]),
          "mem delta MB %.1f" % ((free0 - free1) / 1e6))
    del brain, gs
