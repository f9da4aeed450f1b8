"""GPU parity of the replay memory (csrc/fb_replay.cu): sample indices bit-exact against CPython's
random.sample / the reference SumTree semantics for a fixed word stream; gathered minibatches byte-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import replay_oracle as ro  # noqa: E402


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dqnflappybird_b200 import game, replay
    return game, replay


@pytest.mark.parametrize("N,L,C,t,batch", [
    (1, 50004, 50000, 100, 32),       # pool branch (n <= 277)
    (1, 50004, 50000, 277, 32),       # pool branch boundary
    (1, 50004, 50000, 278, 32),       # set branch just above the boundary
    (1, 50004, 50000, 1001, 32),      # the reference's first training step (OBSERVE = 1000)
    (1, 50004, 50000, 70000, 32),     # full deque: n = 50,000
    (16, 68, 64, 1000, 256),          # n = 1024 <= 1045: pool branch at minibatch 256
    (16, 68, 64, 1000, 32),
    (64, 68, 64, 5000, 256),          # n = 4096, set branch, many duplicates / out-of-range words
    (3, 40, 36, 17, 5),               # tiny batch: setsize 21
    (16384, 64, 60, 5000, 256),       # BASELINE configs[2]: 16,384 envs x 60 transitions = 983,040, minibatch 256
])
def test_uniform_sample_indices_match_cpython(mods, N, L, C, t, batch):
    game, replay = mods
    ring = torch.zeros((N, L, 80, 80), dtype=torch.uint8, device="cuda")
    mem = replay.ReplayMemory(ring, C, seed=99, max_batch=256)
    mem.t = t
    orc = ro.UniformSampler(99)
    n = len(mem)
    assert n == N * min(t, C)
    for rep in range(12):
        mb = mem.sample(batch)
        want = orc.sample(n, batch)
        np.testing.assert_array_equal(mb.idx.cpu().numpy(), np.array(want, np.int32), err_msg=f"rep {rep}")
        assert mem.rng_positions()[0] == orc.pos
        e, k = ro.population_to_transition(np.array(want), t, N, C)
        np.testing.assert_array_equal(mb.env.cpu().numpy(), e)
        np.testing.assert_array_equal(mb.k.cpu().numpy(), k)


def test_sample_larger_than_population_raises(mods):
    game, replay = mods
    ring = torch.zeros((1, 40, 80, 80), dtype=torch.uint8, device="cuda")
    mem = replay.ReplayMemory(ring, 36, seed=1, max_batch=64)
    mem.t = 10
    with pytest.raises(ValueError, match="Sample larger than population"):
        mem.sample(32)


def _rollout_with_memory(game, replay, N, L, C, T, prioritized=False, mode=None):
    gs = game.GameState(num_envs=N, seed=5, history=L)
    mem = (replay.PrioritizedMemory(gs.ring, C, seed=7, max_batch=64, mode=mode) if prioritized
           else replay.ReplayMemory(gs.ring, C, seed=7, max_batch=64))
    hist_f, hist_a, hist_r, hist_t = [], [None], [None], [None]
    gs.frame_step(torch.zeros(N, dtype=torch.uint8, device="cuda"))          # f_0: the driver's initial no-op (FlappyBirdDQN.py:65-66)
    hist_f.append(gs.ring[:, gs.slot].cpu().numpy().copy())
    for k in range(1, T + 1):
        a_row, r_row, t_row = mem.rows(k)
        gs.step_random(1, 0.3, 21, a_row, r_row, t_row, None)
        assert gs.slot == k % L
        mem.appended(k)
        hist_f.append(gs.ring[:, gs.slot].cpu().numpy().copy())
        hist_a.append(a_row.cpu().numpy().copy()); hist_r.append(r_row.cpu().numpy().copy()); hist_t.append(t_row.cpu().numpy().copy())
    return gs, mem, hist_f, hist_a, hist_r, hist_t


def _check_minibatch(mb, hist_f, hist_a, hist_r, hist_t):
    env, ks = mb.env.cpu().numpy(), mb.k.cpu().numpy()
    frames = mb.frames.cpu().numpy()
    for b, (e, k) in enumerate(zip(env, ks)):
        for f in range(5):
            np.testing.assert_array_equal(frames[b, f], hist_f[max(k - 4 + f, 0)][e], err_msg=f"sample {b} env {e} k {k} frame {f}")
        assert mb.actions[b].item() == hist_a[k][e] and mb.rewards[b].item() == hist_r[k][e] and mb.terminals[b].item() == hist_t[k][e]


@pytest.mark.parametrize("T", [3, 30, 90])
def test_gather_returns_the_right_transitions(mods, T):
    """s = frames k-4..k-1, s' = k-3..k; early transitions replicate frame 0 (BrainDQN.py:239); ring wrap-around"""
    game, replay = mods
    N, L, C = 8, 40, 36
    gs, mem, hf, ha, hr, ht = _rollout_with_memory(game, replay, N, L, C, T)
    assert len(mem) == N * min(T, C)
    for _ in range(4):
        mb = mem.sample(min(16, len(mem)))
        ks = mb.k.cpu().numpy()
        assert ks.min() >= max(1, T - C + 1) and ks.max() <= T
        _check_minibatch(mb, hf, ha, hr, ht)


def _check_aux_trees(mem, cap):
    """the min-positive-leaf and max-leaf trees kept beside the SumTree (what replaced round 1's full leaf scans): leaves mirror
    the SumTree's, every inner node is the exact min / max of its children, so the roots are min_p / max_p over all leaves"""
    tree, mn, mx = mem.tree(), mem.aux_tree("min"), mem.aux_tree("max")
    leaves = tree[cap - 1:]
    inf = torch.tensor(float("inf"), dtype=torch.float64, device=tree.device)
    assert torch.equal(mx[cap - 1:], leaves) and torch.equal(mn[cap - 1:], torch.where(leaves > 0, leaves, inf))
    inner = torch.arange(cap - 1, device=tree.device)
    assert torch.equal(mn[inner], torch.minimum(mn[2 * inner + 1], mn[2 * inner + 2]))
    assert torch.equal(mx[inner], torch.maximum(mx[2 * inner + 1], mx[2 * inner + 2]))
    assert mx[0].item() == leaves.max().item()
    if bool((leaves > 0).any()):
        assert mn[0].item() == leaves[leaves > 0].min().item()


# (4096, 8, "rebuild"): 4,096 stores per step take the multi-CTA store kernel (one CTA per group of depth-11 sub-trees); the
# capacity 32,768 is a power of two, (3000, 11) is not (leaves on two levels)
@pytest.mark.parametrize("N,C,mode", [(1, 50, "reference"), (1, 1000, "reference"), (4, 64, "rebuild"), (4, 64, "reference"),
                                      (4096, 8, "rebuild"), (3000, 11, "rebuild")])
def test_sumtree_matches_reference_semantics(mods, N, C, mode):
    """store / sample / batch_update sequence: tree array bit-identical to the oracle (itself pinned to the
    reference classes), sampled tree indices identical, IS weights to 1e-12."""
    game, replay = mods
    L = C + 4
    ring = torch.zeros((N, L, 80, 80), dtype=torch.uint8, device="cuda")
    mem = replay.PrioritizedMemory(ring, C, seed=31, max_batch=64, mode=mode)
    orc = ro.Memory(N, C, seed=31, mode=mode)
    rng = np.random.default_rng(3)
    B = 32 if N * C >= 64 else 8
    for k in range(1, (3 * C + 7) if N <= 1000 else (2 * C + 6)):
        mem.appended(k); orc.store_step(k)
        if k >= 12 and k % 2 == 0:
            mb = mem.sample(B)
            idx, data, w = orc.sample(B)
            np.testing.assert_array_equal(mb.tree_idx.cpu().numpy(), idx, err_msg=f"k={k}")
            np.testing.assert_array_equal(mb.idx.cpu().numpy(), data)
            np.testing.assert_allclose(mb.is_weights.cpu().numpy(), w, rtol=1e-12)
            assert mem.beta == orc.beta
            e, kk = ro.data_index_to_transition(data, k, C)
            np.testing.assert_array_equal(mb.env.cpu().numpy(), e); np.testing.assert_array_equal(mb.k.cpu().numpy(), kk)
            ps = ro.Memory.priorities(rng.random(B) * rng.choice([0.02, 0.7, 3.0])).astype(np.float64)
            mem.batch_update(mb.tree_idx, priorities=torch.from_numpy(ps).cuda())
            orc.batch_update(idx, ps)
        if k % 25 == 0 or (N > 1000 and k % 5 == 0):
            np.testing.assert_array_equal(mem.tree().cpu().numpy(), orc.sum_tree.tree, err_msg=f"tree k={k}")
            _check_aux_trees(mem, N * C)
    np.testing.assert_array_equal(mem.tree().cpu().numpy(), orc.sum_tree.tree)
    _check_aux_trees(mem, N * C)
    assert mem.rng_positions()[1] == orc.pos


def test_priority_transform_on_device(mods):
    """Memory.batch_update's (|err|+0.01 clipped to 1)^0.6 in float32: within 4 float32 ulp of numpy"""
    game, replay = mods
    ring = torch.zeros((1, 68, 80, 80), dtype=torch.uint8, device="cuda")
    mem = replay.PrioritizedMemory(ring, 64, seed=1, max_batch=64)
    for k in range(1, 65):
        mem.appended(k)
    err = np.abs(np.random.default_rng(0).standard_normal(64)).astype(np.float32)
    idx = torch.arange(63, 127, dtype=torch.int32, device="cuda")
    mem.batch_update(idx, abs_errors=torch.from_numpy(err).cuda())
    leaves = mem.tree().cpu().numpy()[63:]
    want = ro.Memory.priorities(err).astype(np.float64)
    np.testing.assert_allclose(leaves, want, rtol=4 * 1.2e-7)
    assert abs(mem.total_p - leaves.sum()) < 1e-9


def test_prioritized_rollout_gather(mods):
    game, replay = mods
    gs, mem, hf, ha, hr, ht = _rollout_with_memory(game, replay, 4, 24, 20, 50, prioritized=True)
    mb = mem.sample(16)
    assert mb.tree_idx.min().item() >= 4 * 20 - 1
    _check_minibatch(mb, hf, ha, hr, ht)


def test_full_size_sumtree_properties(mods):
    """BASELINE configs[2] size -- 16,384 envs x 60 transitions = 983,040 leaves -- through size-independent properties:
    after stores that fill and wrap the memory and prioritized updates, every inner node is exactly left + right (rebuild
    mode), the root is the sum of the leaves, sampled leaves lie in their strata, and the IS weights follow
    (p / min_p)^-beta (BrainPrioritizedReplyDQN.py:127-144) to 1e-12."""
    game, replay = mods
    N, C, B = 16384, 60, 256
    ring = torch.zeros((N, C + 4, 80, 80), dtype=torch.uint8, device="cuda")
    mem = replay.PrioritizedMemory(ring, C, seed=5, max_batch=B)
    cap = N * C
    g = torch.Generator(device="cuda").manual_seed(2)
    for k in range(1, C + 25):
        mem.appended(k)
        if k % 3 == 0:
            mb = mem.sample(B)
            tree = mem.tree()
            total, leaves = tree[0].item(), tree[cap - 1:]
            li = mb.tree_idx.long()
            assert int(li.min()) >= cap - 1 and int(li.max()) <= 2 * cap - 2
            # stratified: the prefix sum just before leaf i is <= (i + 1) * total / B and the one after it >= i * total / B
            # leaves from left to right: the deepest level first (tree indices 2^D - 1 .. 2 cap - 2), then the rest of the
            # level above it (cap - 1 .. 2^D - 2) -- the capacity is not a power of two (BrainPrioritizedReplyDQN.py:39-47)
            D = (2 * cap - 1).bit_length() - 1
            order = torch.cat([torch.arange(2 ** D - 1, 2 * cap - 1, device="cuda"), torch.arange(cap - 1, 2 ** D - 1, device="cuda")])
            csum = torch.cumsum(tree[order], 0)
            pos = torch.empty(2 * cap - 1, dtype=torch.long, device="cuda")
            pos[order] = torch.arange(cap, device="cuda")
            d = li - (cap - 1)
            hi = csum[pos[li]]
            lo = hi - tree[li]
            seg = total / B
            i = torch.arange(B, device="cuda", dtype=torch.float64)
            assert bool((lo <= (i + 1) * seg * (1 + 1e-9)).all()) and bool((hi >= i * seg * (1 - 1e-9)).all())
            min_p = leaves[leaves > 0].min()
            want_w = (leaves[d] / min_p) ** (-mem.beta)
            torch.testing.assert_close(mb.is_weights, want_w, rtol=1e-12, atol=0)
            assert float(mb.is_weights.max()) <= 1.0 + 1e-12
            err = torch.rand(B, device="cuda", generator=g) * 2.0
            mem.batch_update(mb.tree_idx, abs_errors=err)
    _check_aux_trees(mem, cap)
    tree = mem.tree()
    inner = torch.arange(cap - 1, device="cuda")
    assert torch.equal(tree[inner], tree[2 * inner + 1] + tree[2 * inner + 2])       # exact: parents are recomputed, never drifted
    torch.testing.assert_close(tree[0], tree[cap - 1:].sum(), rtol=1e-9, atol=0)
    assert int((tree[cap - 1:] > 0).sum()) == cap                                     # the memory wrapped: every leaf was stored


def test_configs3_shard_size_prioritized_memory(mods):
    """BASELINE configs[3]: 65,536 envs over 8 GPUs = 8,192 envs per rank.  One rank's prioritized memory at that size (8,192 x
    28 = 229,376 leaves, not a power of two): stores through the multi-CTA kernel, samples through the warp-cooperative descent,
    priority updates; the SumTree stays exactly parents = left + right, the min / max trees stay exact, sampled leaves are live
    and the IS weights follow (p / min_p)^-beta with min_p over ALL leaves"""
    game, replay = mods
    N, C, B = 8192, 28, 32
    ring = torch.zeros((N, C + 4, 80, 80), dtype=torch.uint8, device="cuda")
    mem = replay.PrioritizedMemory(ring, C, seed=9, max_batch=B)
    cap = N * C
    g = torch.Generator(device="cuda").manual_seed(4)
    for k in range(1, C + 9):
        mem.appended(k)
        if k % 4 == 0:
            mb = mem.sample(B)
            tree = mem.tree()
            leaves = tree[cap - 1:]
            d = mb.tree_idx.long() - (cap - 1)
            assert bool((leaves[d] > 0).all())
            want_w = (leaves[d] / leaves[leaves > 0].min()) ** (-mem.beta)
            torch.testing.assert_close(mb.is_weights, want_w, rtol=1e-12, atol=0)
            mem.batch_update(mb.tree_idx, abs_errors=torch.rand(B, device="cuda", generator=g) * 2.0)
    _check_aux_trees(mem, cap)
    tree = mem.tree()
    inner = torch.arange(cap - 1, device="cuda")
    assert torch.equal(tree[inner], tree[2 * inner + 1] + tree[2 * inner + 2])


def test_prioritized_sample_with_a_global_minimum(mods):
    """several ranks, one memory (fb_per_min_root / fb_per_set_global_min): ISWeights = (p_i / min_p)^-beta with min_p the
    minimum over ALL shards, handed in as a device scalar (the ranks MIN-reduce it; here: set by hand)"""
    game, replay = mods
    N, C = 16, 8
    ring = torch.zeros((N, C + 4, 80, 80), dtype=torch.uint8, device="cuda")
    mem = replay.PrioritizedMemory(ring, C, seed=3, max_batch=32)
    for k in range(1, C + 3):
        mem.appended(k)
    cap = N * C
    idx = torch.arange(cap - 1, cap - 1 + 32, dtype=torch.int32, device="cuda")
    pr = torch.linspace(0.05, 2.0, 32, dtype=torch.float64, device="cuda")
    mem.batch_update(idx, priorities=pr)
    L = replay._lib.lib()
    root = torch.empty(1, dtype=torch.float64, device="cuda")
    replay._lib.check(L.fb_per_min_root(mem._h, root.data_ptr(), torch.cuda.current_stream().cuda_stream), "fb_per_min_root")
    leaves = mem.tree()[cap - 1:]
    assert root.item() == leaves[leaves > 0].min().item() == 0.05
    pos = mem.rng_positions()
    mb = mem.sample(32)
    w_local = mb.is_weights.clone(); picked = mb.tree_idx.clone()
    gmin = torch.tensor([0.0125], dtype=torch.float64, device="cuda")                 # another shard holds a smaller leaf
    replay._lib.check(L.fb_per_set_global_min(mem._h, gmin.data_ptr()), "fb_per_set_global_min")
    mem.set_rng_positions(pos); mem.beta -= mem.beta_increment_per_sampling           # the same draw again
    mb = mem.sample(32)
    assert torch.equal(mb.tree_idx, picked)
    p = mem.tree()[picked.long()]
    assert torch.allclose(mb.is_weights, (p / 0.0125) ** (-mem.beta), rtol=1e-12, atol=0)
    assert torch.allclose(mb.is_weights, w_local * 4.0 ** (-mem.beta), rtol=1e-12, atol=0)
    replay._lib.check(L.fb_per_set_global_min(mem._h, None), "fb_per_set_global_min")
    mem.set_rng_positions(pos); mem.beta -= mem.beta_increment_per_sampling
    assert torch.equal(mem.sample(32).is_weights, w_local)
