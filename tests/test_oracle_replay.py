"""Replay oracle pinned against the reference's own SumTree/Memory (executed verbatim by make_golden.py)
and against CPython's random.sample."""
import os
import random

import numpy as np

from oracle import replay_oracle as ro


def test_sumtree_oracle_reproduces_reference_classes(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_per.npz"))
    for case in range(int(g["n_cases"])):
        cap, n_store, batch = int(g[f"c{case}_cap"]), int(g[f"c{case}_nstore"]), int(g[f"c{case}_batch"])
        mem = ro.Memory(1, cap, seed=0, mode="reference")
        k = 0
        for _ in range(n_store):
            k += 1
            mem.store_step(k)
        V, IDX, W, ERR, TOT, DATA = (g[f"c{case}_{n}"] for n in ("v", "idx", "w", "abs_err", "total", "data"))
        for r in range(len(V)):
            T = mem.sum_tree
            assert T.total_p == TOT[r], (case, r)
            mem.beta = np.min([1., mem.beta + mem.beta_increment_per_sampling])
            leaves = T.tree[-T.capacity:]
            min_prob = leaves[leaves > 0].min() / T.total_p
            for i in range(batch):
                idx, p, data = T.get_leaf(V[r][i])                 # the uniforms the reference drew
                assert idx == IDX[r][i], (case, r, i)
                assert np.power(p / T.total_p / min_prob, -mem.beta) == W[r][i]
                # data payloads in the golden run are the store counters: position p holds the latest k with (k-1)%cap == p
            ps = np.power(np.minimum(ERR[r] + mem.epsilon, mem.abs_err_upper), mem.alpha)   # float64 path of the golden run
            mem.batch_update(IDX[r], ps)
            for _ in range(3):
                k += 1
                mem.store_step(k)
        np.testing.assert_array_equal(mem.sum_tree.tree, g[f"c{case}_final_tree"])
        assert mem.beta == float(g[f"c{case}_final_beta"])


def test_rebuild_mode_agrees_with_reference_mode_to_rounding():
    rng = np.random.default_rng(0)
    a, b = ro.Memory(3, 50, 1, "reference"), ro.Memory(3, 50, 1, "rebuild")
    for k in range(1, 120):
        a.store_step(k); b.store_step(k)
        if k > 10:
            ia, _, wa = a.sample(8); ib, _, wb = b.sample(8)
            np.testing.assert_array_equal(ia, ib)
            ps = ro.Memory.priorities(rng.random(8) * 2).astype(np.float64)
            a.batch_update(ia, ps); b.batch_update(ib, ps)
    np.testing.assert_allclose(a.sum_tree.tree, b.sum_tree.tree, rtol=1e-12)
    # rebuild mode is exactly consistent: every inner node is the sum of its children
    t = b.sum_tree.tree
    inner = np.arange(len(t) // 2)
    np.testing.assert_array_equal(t[inner], t[2 * inner + 1] + t[2 * inner + 2])


def test_word_stream_random_is_cpython_random():
    """WordStreamRandom fed with MT19937's own words must reproduce random.Random exactly"""
    from oracle.qnet_oracle import WordStreamRandom
    src = random.Random(12345)
    ref = random.Random(12345)
    R = WordStreamRandom(lambda: src.getrandbits(32))
    for n, k in ((50000, 32), (300, 32), (277, 32), (100, 32), (1200, 256), (1000, 256), (7, 3)):
        assert R.sample(range(n), k) == ref.sample(range(n), k)
    for _ in range(50):
        assert R.random() == ref.random()
        assert R.randrange(2) == ref.randrange(2)
    assert ro.cpython_setsize(32) == 277 and ro.cpython_setsize(256) == 1045 and ro.cpython_setsize(512) == 4117


def test_product_setsize_equals_cpythons_expression():
    """random.sample's set / list switch (21 + 4**ceil(log(3k, 4)) for k > 5, CPython's float expression): the product's
    host helper, the oracle's and the interpreter's own must agree for every minibatch size the library accepts"""
    from math import ceil, log
    from dqnflappybird_b200 import replay
    for k in range(1, 513):
        want = 21 + (4 ** ceil(log(k * 3, 4)) if k > 5 else 0)
        assert replay.cpython_setsize(k) == ro.cpython_setsize(k) == want, k
