"""torchrun worker (one process per GPU): peer-memory gradient exchange + Adam against NCCL all-reduce + Adam, then a
short data-parallel BrainDQNNature run whose parameters must stay bitwise identical on every rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from dqnflappybird_b200 import dist as fdist, qnet  # noqa: E402
from dqnflappybird_b200.brains import BrainDQNNature  # noqa: E402
from dqnflappybird_b200.game import GameState  # noqa: E402


def main():
    rank, world, local = fdist.init("nccl")
    dev = torch.device("cuda", local)
    net = qnet.QNetwork(device=dev, max_batch=8, seed=3, precision="fp32")
    ref = qnet.QNetwork(device=dev, max_batch=8, seed=3, precision="fp32")
    x = net.enable_peer_exchange()
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for step in range(8):
        x.set_two_shot(step >= 4)                   # both forms of the exchange, whatever the world size
        mine = torch.randn(net.n_params, device=dev, generator=g) * 0.01
        net.grads.copy_(mine)
        net.adam_step()
        tot = mine.clone()
        dist.all_reduce(tot)
        ref.grads.copy_(tot)
        ref.adam_step()
        np.testing.assert_allclose(net.params.cpu().numpy(), ref.params.cpu().numpy(), rtol=0, atol=3e-7)
    everyone = [torch.empty_like(net.params) for _ in range(world)]
    dist.all_gather(everyone, net.params)
    assert all(torch.equal(everyone[0], e) for e in everyone), "parameters differ between ranks"

    # data-parallel training: envs / replay sharded, per-rank minibatch, gradients summed inside the Adam kernel
    N = 64
    brain = BrainDQNNature(2, "bird", num_envs=N, device=dev, replay_memory_per_env=12, observe=6., batch_size=32 * world, seed=1,
                           first_env_id=rank * N, replace_target_iter=4)
    assert brain.net.exchange is not None and brain.local_batch == 32
    gs = GameState(num_envs=N, device=dev, seed=2, history=16, first_env_id=rank * N, ring=brain.ring)
    obs, *_ = gs.frame_step(torch.zeros(N, dtype=torch.uint8, device=dev))
    brain.setInitState(obs)
    p0 = brain.net.params.clone()
    for _ in range(20):
        a = brain.getAction()
        obs, r, t, s = gs.frame_step(a)
        brain.setPerception(obs, a, r, t, s)
    assert brain.net.adam_steps == 20 - 7 and not torch.equal(brain.net.params, p0)
    dist.all_gather(everyone[:world], brain.net.params) if everyone[0].shape == brain.net.params.shape else None
    ps = [torch.empty_like(brain.net.params) for _ in range(world)]
    dist.all_gather(ps, brain.net.params)
    assert all(torch.equal(ps[0], e) for e in ps), "replicated learners diverged"
    assert torch.isfinite(brain.net.params).all()
    dist.barrier()
    if rank == 0:
        print("PEER_EXCHANGE_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
