"""GPU parity of the Q-network path (csrc/fb_qnet.cu, strict fp32) against the float64 oracle.

Stated tolerances (fp32 CUDA-core path, fp32 FMA accumulation vs float64 reference):
  Q-values          |dq| <= 1e-5 + 1e-4 |q|
  loss              relative 1e-4
  gradients         per tensor  ||g - g_ref||_2 <= 2e-4 ||g_ref||_2   (deterministic split-K)
  Adam step         parameters after 3 steps: |dp| <= 3e-7 (lr 1e-6 => updates are ~1e-6 per step)
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import qnet_oracle as qo  # noqa: E402


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dqnflappybird_b200 import game, qnet
    return game, qnet


def _env_frames(game, B, seed):
    """[B][5][80][80] real observations: 5 consecutive frames of B envs after a random warm-up"""
    gs = game.GameState(num_envs=B, seed=seed, history=5)
    gs.step_random(40 + seed % 7, 0.3, 99 + seed)
    gs.step_random(5, 0.3, 99 + seed)
    torch.cuda.synchronize()
    L = 5
    order = [(gs.slot - 4 + k) % L for k in range(5)]
    return gs.ring[:, order].contiguous()


def _set_params(net, flat, target=None):
    net.params.copy_(torch.from_numpy(flat))
    net.target.copy_(torch.from_numpy(target if target is not None else flat))


@pytest.mark.parametrize("dueling", [False, True])
def test_forward_matches_oracle(mods, dueling):
    game, qnet = mods
    B = 37
    frames = _env_frames(game, B, 3)
    net = qnet.QNetwork(precision="fp32", hidden=512, dueling=dueling, max_batch=16)          # 37 > 16: exercises chunking
    flat = qo.init_params(512, dueling, seed=5) * np.float32(4.0)           # larger weights -> non-trivial activations
    _set_params(net, flat)
    q = net.forward(qnet.FrameBatch.from_stack(frames, 0)).cpu().numpy()
    q2 = net.forward(qnet.FrameBatch.from_stack(frames, 1)).cpu().numpy()
    x = frames.cpu().numpy()
    ref = qo.forward(torch.tensor(flat.astype(np.float64)), x[:, 0:4], 512, dueling).numpy()
    ref2 = qo.forward(torch.tensor(flat.astype(np.float64)), x[:, 1:5], 512, dueling).numpy()
    assert np.abs(ref).max() > 0.05
    np.testing.assert_allclose(q, ref, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(q2, ref2, rtol=1e-4, atol=1e-5)
    # noise frames too (dense inputs)
    noise = (torch.rand((9, 5, 80, 80), device="cuda") < 0.3).to(torch.uint8) * 255
    qn = net.forward(qnet.FrameBatch.from_stack(noise.contiguous(), 0)).cpu().numpy()
    refn = qo.forward(torch.tensor(flat.astype(np.float64)), noise.cpu().numpy()[:, 0:4], 512, dueling).numpy()
    np.testing.assert_allclose(qn, refn, rtol=1e-4, atol=1e-4)


def test_forward_from_ring_view(mods):
    """acting reads the ring in place: channel order follows the slot rotation (newest last, BrainDQN.py:68)"""
    game, qnet = mods
    N = 21
    gs = game.GameState(num_envs=N, seed=1, history=7)
    gs.step_random(23, 0.4, 5)
    net = qnet.QNetwork(precision="fp32", max_batch=32)
    q = net.forward(qnet.FrameBatch.from_ring(gs.ring, gs.slot)).cpu().numpy()
    st = gs.stacked_state().permute(0, 3, 1, 2).contiguous().cpu().numpy()      # [N,4,80,80], newest last
    ref = qo.forward(torch.tensor(net.params.cpu().numpy().astype(np.float64)), st).numpy()
    np.testing.assert_allclose(q, ref, rtol=1e-4, atol=1e-5)


CASES = [("vanilla", False, True, False), ("nature", False, False, False), ("double", False, False, False),
         ("nature", True, False, False), ("nature", False, False, True), ("double", True, False, True)]


@pytest.mark.parametrize("variant,dueling,loss_sum,per", CASES)
def test_loss_and_gradients_match_oracle(mods, variant, dueling, loss_sum, per):
    game, qnet = mods
    B = 32
    frames = _env_frames(game, B, 11)
    net = qnet.QNetwork(precision="fp32", hidden=512, dueling=dueling, max_batch=B)
    p = qo.init_params(512, dueling, seed=1) * np.float32(3.0)
    t = qo.init_params(512, dueling, seed=2) * np.float32(3.0)
    _set_params(net, p, t)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2, B).astype(np.uint8)
    r = rng.choice(np.array([0.1, 3.0, -3.0], np.float32), B, p=[0.8, 0.1, 0.1])
    term = (r == -3.0).astype(np.uint8)
    isw = rng.random(B).astype(np.float32) if per else None
    abs_err = torch.zeros(B, device="cuda"); y = torch.zeros(B, device="cuda")
    net.loss_backward(variant, frames, torch.from_numpy(a).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(term).cuda(),
                      torch.from_numpy(isw).cuda() if per else None, 0.99, loss_sum, None, abs_err, y)
    x = frames.cpu().numpy()
    loss, g_ref, ae_ref, y_ref, _ = qo.loss_and_grads(qnet.VARIANTS[variant], p, t, x[:, 0:4], x[:, 1:5], a, r, term, isw, 0.99,
                                                      loss_sum, None, 512, dueling)
    assert abs(net.loss.item() - loss) <= 1e-4 * abs(loss)
    np.testing.assert_allclose(y.cpu().numpy(), y_ref, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(abs_err.cpu().numpy(), ae_ref, rtol=1e-4, atol=1e-5)
    g = net.grads.cpu().numpy().astype(np.float64)
    L = qo.layout(512, dueling)
    for name, v in L.items():
        if name == "total":
            continue
        o, shp = v
        sz = int(np.prod(shp))
        num = np.linalg.norm(g[o:o + sz] - g_ref[o:o + sz]); den = np.linalg.norm(g_ref[o:o + sz])
        assert den > 0, name
        assert num <= 2e-4 * den, (name, num, den)


def test_sharded_gradients_sum_to_global_gradient(mods):
    """two shards of B/2 with global_batch = B; the sum of their gradients is the gradient of the global mean loss"""
    game, qnet = mods
    B = 32
    frames = _env_frames(game, B, 5)
    net = qnet.QNetwork(precision="fp32", max_batch=B)
    rng = np.random.default_rng(1)
    a = torch.from_numpy(rng.integers(0, 2, B).astype(np.uint8)).cuda()
    r = torch.from_numpy(rng.choice(np.array([0.1, 3.0, -3.0], np.float32), B)).cuda()
    term = (r == -3.0).to(torch.uint8)
    net.loss_backward("nature", frames, a, r, term)
    g_full = net.grads.clone(); loss_full = net.loss.item()
    h = B // 2
    net.loss_backward("nature", frames[:h].contiguous(), a[:h], r[:h], term[:h], global_batch=B)
    g0 = net.grads.clone(); l0 = net.loss.item()
    net.loss_backward("nature", frames[h:].contiguous(), a[h:], r[h:], term[h:], global_batch=B)
    g1 = net.grads.clone(); l1 = net.loss.item()
    assert abs((l0 + l1) - loss_full) <= 1e-5 * abs(loss_full)
    err = (g0 + g1 - g_full).norm().item() / g_full.norm().item()
    assert err < 1e-5, err


def test_adam_matches_tf1_rule(mods):
    game, qnet = mods
    net = qnet.QNetwork(precision="fp32", max_batch=8, seed=3)
    p = net.params.cpu().numpy().copy()
    ref = qo.AdamTF1(len(p))
    rng = np.random.default_rng(2)
    for step in range(3):
        g = (rng.standard_normal(len(p)) * 10 ** rng.uniform(-6, 1, len(p))).astype(np.float32)
        net.grads.copy_(torch.from_numpy(g))
        net.adam_step()
        p = ref.step(p, g)
    np.testing.assert_allclose(net.params.cpu().numpy(), p, rtol=0, atol=3e-7)
    np.testing.assert_allclose(net.adam_m.cpu().numpy(), ref.m, rtol=1e-5, atol=2e-7)     # FMA contraction on the device
    np.testing.assert_allclose(net.adam_v.cpu().numpy(), ref.v, rtol=1e-5, atol=1e-9)
    assert net.beta1_power == ref.b1p and net.beta2_power == ref.b2p
    net.sync_target()
    assert torch.equal(net.target, net.params)


def test_epsilon_greedy_matches_cpython_random(mods):
    game, qnet = mods
    N = 300
    gs = game.GameState(num_envs=N, seed=4)
    gs.step_random(9, 0.5, 1)
    net = qnet.QNetwork(precision="fp32", max_batch=512)
    pos = torch.zeros(N, dtype=torch.int32, device="cuda")
    pos_ref = np.zeros(N, np.int64)
    acts = torch.zeros(N, dtype=torch.uint8, device="cuda"); q = torch.zeros((N, 2), device="cuda")
    for eps in (0.0, 0.03, 0.5, 1.0, -6.66460567250383e-13):
        net.act(qnet.FrameBatch.from_ring(gs.ring, gs.slot), eps, 777, 50, pos, acts, q)
        want = qo.egreedy_actions(q.cpu().numpy(), eps, 777, 50, pos_ref)
        np.testing.assert_array_equal(acts.cpu().numpy(), want)
        np.testing.assert_array_equal(pos.cpu().numpy(), pos_ref)
