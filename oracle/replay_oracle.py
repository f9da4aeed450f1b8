"""Replay-memory oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Uniform replay: the reference calls ``random.sample(self.replayMemory, BATCH_SIZE)`` on a deque
(BrainDQN.py:197).  The oracle IS CPython's own ``random.sample`` run on a ``random.Random`` subclass
whose words come from the same Philox stream the device consumes (qnet_oracle.WordStreamRandom).

Prioritized replay: numpy restatement of SumTree / Memory (BrainPrioritizedReplyDQN.py:32-151), line for
line, with the uniforms supplied by the caller's word stream instead of the global np.random.  It is
pinned against the reference's own classes executed verbatim (tests/golden/ref_per.npz, produced by
tests/golden/make_golden.py by AST-extracting the two ClassDefs).  ``rebuild`` mode restates the
batched variant (set leaves, recompute touched ancestors as left+right).
"""
from __future__ import annotations

import math

import numpy as np

from . import flappy_oracle as fo
from .qnet_oracle import WordStreamRandom


def cpython_setsize(k: int) -> int:
    """Lib/random.py sample(): setsize = 21; if k > 5: setsize += 4 ** _ceil(_log(k * 3, 4))"""
    setsize = 21
    if k > 5:
        setsize += 4 ** math.ceil(math.log(k * 3, 4))
    return setsize


class UniformSampler:
    def __init__(self, seed: int):
        self.seed, self.pos = seed, 0
        self.R = WordStreamRandom(self._next)

    def _next(self):
        w = fo.stream_word(self.seed, 3, 0, self.pos)
        self.pos += 1
        return w

    def sample(self, n: int, k: int):
        return self.R.sample(range(n), k)


def population_to_transition(j, t, N, C):
    """population index -> (env, k): j = e*cnt + (k-k_lo)"""
    k_lo = max(1, t - C + 1)
    cnt = t - k_lo + 1
    return j // cnt, k_lo + j % cnt


def data_index_to_transition(d, t, C):
    e, p = d // C, d % C
    return e, t - ((t - 1 - p) % C)


class SumTree:
    """BrainPrioritizedReplyDQN.py:32-104 (data array omitted: the data index is the payload)"""

    def __init__(self, capacity):
        self.capacity = capacity
        self.tree = np.zeros(2 * capacity - 1)
        self.size = 0
        self.data_pointer = 0

    def add(self, p):                                  # :50-60
        tree_idx = self.data_pointer + (self.capacity - 1)
        self.update(tree_idx, p)
        self.data_pointer += 1
        if self.data_pointer >= self.capacity:
            self.data_pointer = 0
        if self.size < self.capacity:
            self.size += 1

    def update(self, tree_idx, p):                     # :62-68
        change = p - self.tree[tree_idx]
        self.tree[tree_idx] = p
        while tree_idx != 0:
            tree_idx = (tree_idx - 1) // 2
            self.tree[tree_idx] += change

    def get_leaf(self, v):                             # :73-100
        parent_idx = 0
        while True:
            cl_idx = 2 * parent_idx + 1
            cr_idx = cl_idx + 1
            if cl_idx >= len(self.tree):
                leaf_idx = parent_idx
                break
            if v <= self.tree[cl_idx]:
                parent_idx = cl_idx
            else:
                v -= self.tree[cl_idx]
                parent_idx = cr_idx
        return leaf_idx, self.tree[leaf_idx], leaf_idx - self.capacity + 1

    @property
    def total_p(self):
        return self.tree[0]


def _depth(idx):
    return (idx + 1).bit_length() - 1


class Memory:
    """BrainPrioritizedReplyDQN.py:107-151 over N envs x C transitions (N = 1: the reference itself)."""
    epsilon = 0.01
    alpha = 0.6
    beta_increment_per_sampling = 0.001
    abs_err_upper = 1.

    def __init__(self, n_envs, cap_per_env, seed, mode="reference"):
        self.N, self.C = n_envs, cap_per_env
        self.sum_tree = SumTree(n_envs * cap_per_env)
        self.beta = 0.4
        self.seed, self.pos = seed, 0
        self.mode = mode

    def _word(self):
        w = fo.stream_word(self.seed, 4, 0, self.pos)
        self.pos += 1
        return w

    def _set_many(self, tree_idx, ps):
        T = self.sum_tree
        if self.mode == "reference":
            for ti, p in zip(tree_idx, ps):
                T.update(int(ti), float(p))
            return
        for ti, p in zip(tree_idx, ps):               # rebuild: later duplicates overwrite earlier ones
            T.tree[int(ti)] = float(p)
        max_d = _depth(len(T.tree) - 1)
        for d in range(max_d - 1, -1, -1):
            for ti in tree_idx:
                dl = _depth(int(ti))
                if dl > d:
                    node = ((int(ti) + 1) >> (dl - d)) - 1
                    T.tree[node] = T.tree[2 * node + 1] + T.tree[2 * node + 2]

    def store_step(self, k):
        """Memory.store (:121-125) of transition k for every env, env order"""
        T = self.sum_tree
        max_p = np.max(T.tree[-T.capacity:])
        if max_p == 0:
            max_p = self.abs_err_upper
        leaves = [(T.capacity - 1) + e * self.C + (k - 1) % self.C for e in range(self.N)]
        self._set_many(leaves, [max_p] * self.N)

    def sample(self, n):                               # :127-144
        T = self.sum_tree
        b_idx, b_data, ISWeights = np.empty((n,), np.int32), np.empty((n,), np.int32), np.empty((n,))
        pri_seg = T.total_p / n
        self.beta = np.min([1., self.beta + self.beta_increment_per_sampling])
        leaves = T.tree[-T.capacity:]
        min_prob = leaves[leaves > 0].min() / T.total_p      # == get_min_prob (:70-71): filled leaves are > 0
        for i in range(n):
            a, b = pri_seg * i, pri_seg * (i + 1)
            w0, w1 = self._word() >> 5, self._word() >> 6     # np.random.uniform: a + (b-a) * rk_double
            u = (w0 * 67108864.0 + w1) / 9007199254740992.0
            v = a + (b - a) * u
            idx, p, data = T.get_leaf(v)
            prob = p / T.total_p
            ISWeights[i] = np.power(prob / min_prob, -self.beta)
            b_idx[i], b_data[i] = idx, data
        return b_idx, b_data, ISWeights

    @staticmethod
    def priorities(abs_errors):
        """the transform of batch_update (:146-149), float32 like the arrays TensorFlow returns"""
        e = np.asarray(abs_errors, np.float32) + np.float32(0.01)
        return np.power(np.minimum(e, np.float32(1.)), np.float32(0.6))

    def batch_update(self, tree_idx, ps):              # :150-151, with the priorities already transformed
        self._set_many(tree_idx, ps)
