"""torchrun worker (one process per GPU): peer-memory gradient exchange + Adam against NCCL all-reduce + Adam, then a
short data-parallel BrainDQNNature run whose parameters must stay bitwise identical on every rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from dqnflappybird_b200 import dist as fdist, qnet  # noqa: E402
from dqnflappybird_b200.brains import BrainDQNNature, BrainPrioritizedReplyDQN  # noqa: E402
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

    # ---- the exchange INSIDE the training step's graph (tensor-core path): per-rank minibatches, W_fc1's bucket summed and
    # Adam-updated beside the convolution gradients, the rest at the tail -- against loss_backward + NCCL all-reduce + Adam
    B = 32
    gsx = GameState(num_envs=B, device=dev, seed=50 + rank, history=5)
    gsx.step_random(45, 0.3, 11 + rank)
    order = [(gsx.slot - 4 + k) % 5 for k in range(5)]
    frames = gsx.ring[:, order].contiguous()
    gr = torch.Generator(device=dev).manual_seed(7 + rank)
    for precision in ("fp16", "bf16"):
        fused = qnet.QNetwork(device=dev, max_batch=B, seed=5, precision=precision, lr=1e-3)
        plain = qnet.QNetwork(device=dev, max_batch=B, seed=5, precision=precision, lr=1e-3)
        for nn in (fused, plain):
            nn.params.mul_(4.0); nn.target.mul_(4.0)
        fused.enable_peer_exchange()
        assert fused.exchange_in_step
        for step in range(6):                          # eager, eager, then graph replays (two graphs: one per exchange buffer)
            a = (torch.rand(B, device=dev, generator=gr) < 0.5).to(torch.uint8)
            r = torch.where(torch.rand(B, device=dev, generator=gr) < 0.2, torch.tensor(3.0, device=dev), torch.tensor(0.1, device=dev))
            term = torch.zeros(B, dtype=torch.uint8, device=dev)
            fused.train_step("nature", frames, a, r, term, global_batch=B * world)
            plain.loss_backward("nature", frames, a, r, term, global_batch=B * world)
            # the sum in rank order, as the exchange kernels form it (NCCL's order differs from 4 ranks up, and the first Adam
            # steps, lr * g / (|g| + eps), amplify last-bit differences of near-zero gradients: 1.5e-5 at 8 ranks)
            gl = [torch.empty_like(plain.grads) for _ in range(world)]
            dist.all_gather(gl, plain.grads)
            tot = gl[0].clone()
            for gk in gl[1:]:
                tot += gk
            nccl = plain.grads.clone()
            dist.all_reduce(nccl)
            assert torch.allclose(nccl, tot, rtol=1e-5, atol=1e-6 * tot.abs().max().item())
            plain.grads.copy_(tot)
            plain.adam_step()
            assert torch.allclose(fused.params, plain.params, rtol=0, atol=2e-6), (precision, step, (fused.params - plain.params).abs().max().item())
            assert fused.beta1_power == plain.beta1_power and fused.adam_steps == step + 1
            if step == 3:
                fused.sync_target(); plain.sync_target()
        ps = [torch.empty_like(fused.params) for _ in range(world)]
        dist.all_gather(ps, fused.params)
        assert all(torch.equal(ps[0], e) for e in ps), "in-step exchange: replicas diverged"
        moved = (fused.params - plain.params).abs().max().item()
        upd = (fused.params.mean() * 0 + (plain.adam_m.abs().max())).item()
        assert upd > 0, "no update happened"
        # the next forward uses the operand copies the bucket kernels rewrote
        qf = fused.forward(qnet.FrameBatch.from_stack(frames, 0)); qp = plain.forward(qnet.FrameBatch.from_stack(frames, 0))
        assert torch.allclose(qf, qp, rtol=0, atol=2e-2 * qp.abs().max().item()), moved
        del fused, plain

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

    # ---- prioritized replay sharded over the ranks: Memory.sample's min_prob is the minimum over EVERY shard's leaves (one MIN
    # all-reduce of a scalar), so ISWeights are those of one global memory: (p_i / min_p_global)^-beta
    pb = BrainPrioritizedReplyDQN(2, "bird", num_envs=N, device=dev, replay_memory_per_env=12, observe=6., batch_size=32 * world, seed=4,
                                  first_env_id=rank * N, replace_target_iter=4)
    mem = pb.replayMemory
    assert mem._gmin is not None
    gs2 = GameState(num_envs=N, device=dev, seed=6, history=16, first_env_id=rank * N, ring=pb.ring)
    obs, *_ = gs2.frame_step(torch.zeros(N, dtype=torch.uint8, device=dev))
    pb.setInitState(obs)
    for _ in range(16):
        a = pb.getAction()
        obs, r, t, s = gs2.frame_step(a)
        pb.setPerception(obs, a, r, t, s)
    assert pb.net.adam_steps == 16 - 7
    ps = [torch.empty_like(pb.net.params) for _ in range(world)]
    dist.all_gather(ps, pb.net.params)
    assert all(torch.equal(ps[0], e) for e in ps), "prioritized learners diverged"
    # ranks hold different priorities by now; make the minimum live on the LAST rank only, then sample
    cap = mem.N * mem.C
    leaves = mem.tree()[cap - 1:]
    local_min = leaves[leaves > 0].min()
    mins = [torch.empty_like(local_min) for _ in range(world)]
    dist.all_gather(mins, local_min)
    if rank == world - 1:
        idx = torch.tensor([cap - 1 + 5], dtype=torch.int32, device=dev)
        mem.batch_update(idx, priorities=torch.tensor([float(min(m.item() for m in mins)) * 0.25], dtype=torch.float64, device=dev))
    leaves = mem.tree()[cap - 1:]
    local_min = leaves[leaves > 0].min()
    dist.all_gather(mins, local_min)
    gmin = min(m.item() for m in mins)
    assert (rank == world - 1) == (local_min.item() == gmin)
    mb = mem.sample(32)
    p = mem.tree()[mb.tree_idx.long()]
    want = (p / gmin) ** (-mem.beta)
    assert torch.allclose(mb.is_weights, want, rtol=1e-12, atol=0), (mb.is_weights - want).abs().max().item()
    if rank != world - 1:
        assert (mb.is_weights < (p / local_min.item()) ** (-mem.beta)).all()       # the local minimum would have given larger weights
    dist.barrier()
    if rank == 0:
        print("PEER_EXCHANGE_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
