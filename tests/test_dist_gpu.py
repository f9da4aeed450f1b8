"""Gradient exchange fused with Adam over peer memory (csrc/fb_dist.cu).

On one GPU: two "ranks" inside one process (fb_dist_connect_local, handshake skipped because the two kernels run one
after the other).  On >= 2 GPUs: real processes, CUDA IPC + NVLink, the publish / wait handshake
(tests/dist_peer_exchange_worker.py under torchrun) -- skipped when the box has a single GPU.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dqnflappybird_b200 import dist, qnet
    return dist, qnet


def test_two_ranks_in_one_process_sum_and_adam(mods):
    dist, qnet = mods
    nets = [qnet.QNetwork(max_batch=8, seed=3, precision="fp32") for _ in range(2)]
    ref = qnet.QNetwork(max_batch=8, seed=3, precision="fp32")
    xs = [dist.PeerGradExchange(nets[0].n_params, "cuda:0", rank=r, world=2, connect=False) for r in range(2)]
    xs[0].connect_local(1, xs[1]); xs[1].connect_local(0, xs[0])
    for r in range(2):
        xs[r].no_wait = True
        nets[r].enable_peer_exchange(xs[r])
    g = torch.Generator(device="cuda").manual_seed(0)
    for step in range(4):
        gs = [torch.randn(ref.n_params, device="cuda", generator=g) * 10 ** torch.empty(ref.n_params, device="cuda").uniform_(-5, 0, generator=g)
              for _ in range(2)]
        bufs = [nets[r].grads for r in range(2)]
        assert bufs[0].data_ptr() != bufs[1].data_ptr()
        for r in range(2):
            nets[r].grads.copy_(gs[r])
        for r in range(2):
            nets[r].adam_step()
        assert nets[0].grads.data_ptr() != bufs[0].data_ptr()          # the exchange buffers alternate by step parity
        ref.grads.copy_(gs[0] + gs[1])
        ref.adam_step()
        assert torch.equal(nets[0].params, nets[1].params)             # same order of summation on every rank
        np.testing.assert_allclose(nets[0].params.cpu().numpy(), ref.params.cpu().numpy(), rtol=0, atol=2e-7)
        np.testing.assert_allclose(nets[0].adam_v.cpu().numpy(), ref.adam_v.cpu().numpy(), rtol=1e-5, atol=1e-12)


def test_peer_exchange_across_processes():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_peer_exchange_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PEER_EXCHANGE_OK" in r.stdout


def test_exchange_adam_refreshes_the_bf16_operand_copies(mods):
    """tensor-core path: the exchange + Adam kernel rewrites the bf16 copies of every parameter it updates, so the next step
    starts without a pack kernel -- Q-values after it equal those of a net whose copies were rebuilt from the same weights"""
    dist, qnet = mods
    from dqnflappybird_b200 import game
    nets = [qnet.QNetwork(max_batch=16, seed=3, precision="bf16") for _ in range(2)]
    xs = [dist.PeerGradExchange(nets[0].n_params, "cuda:0", rank=r, world=2, connect=False) for r in range(2)]
    xs[0].connect_local(1, xs[1]); xs[1].connect_local(0, xs[0])
    for r in range(2):
        xs[r].no_wait = True
        nets[r].enable_peer_exchange(xs[r])
        nets[r].lr = np.float32(1e-3)
    gs = game.GameState(num_envs=16, seed=5, history=6)
    gs.step_random(20, 0.4, 3)
    fb = qnet.FrameBatch.from_ring(gs.ring, gs.slot)
    q0 = nets[0].forward(fb).clone()                               # packs the operand copies of the initial weights
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(3):
        for r in range(2):
            nets[r].grads.copy_(torch.randn(nets[r].n_params, device="cuda", generator=g))
        for r in range(2):
            nets[r].adam_step()
    q1 = nets[0].forward(fb)
    fresh = qnet.QNetwork(max_batch=16, seed=9, precision="bf16")
    fresh.params.copy_(nets[0].params)
    assert not torch.equal(q1, q0)
    assert torch.equal(q1, fresh.forward(fb))
