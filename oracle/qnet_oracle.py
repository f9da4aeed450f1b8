"""Q-network oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

float64 torch-CPU restatement of the TensorFlow-1.12 graph the reference builds.

What IS pinned (round 2): the graph itself.  Every checkpoint under train_history/ carries the serialized GraphDef that
_createQNetwork built (forward ops with their attributes, the loss, the complete tf.gradients sub-graph, the ApplyAdam ops
with their constants).  oracle/tf_graph.py parses it (fixtures tests/golden/ref_graph_*.json) and executes it op by op in
NumPy float64; tests/test_oracle_qnet_graph.py holds this file to it: op sequence, strides, paddings, variable shapes in
creation order, Adam constants -- and Q(s), Q_target(s'), the cost, every gradient tensor and two optimizer steps to 1e-9.

What stays PARITY UNPINNED: TensorFlow 1.12.0's own floating-point kernels (cuDNN / Eigen summation order; third-party, not
vendored, not installable here; version from train_history/*/bird-*.meta).  The reference ships neither tests, weights nor
recorded outputs for them, so the per-op arithmetic is restated from TF 1.12's published semantics and anchored on the
reference's call sites:

  BrainDQN.py:119-163        graph: NHWC, HWIO weights, SAME padding (symmetric 2 / 1 / 1 here),
                             one 2x2 max-pool, flatten (h,w,c), fc 1600->512->2, sum-of-squares loss
  BrainDQNNature.py:107-119  target net + reduce_mean loss; :149-183 Nature target
  BrainDoubleDQN.py:51-61    Double target
  BrainDuelingDQN_CC.py:68-77  dueling head Q = V + (A - mean_a A), b_fc2_v [1,1], b_fc2_a [1,A]
  BrainPrioritizedReplyDQN.py:245-253  abs_errors, IS-weighted mean loss
  BrainDQN.py:99-116         getAction epsilon-greedy, float64 epsilon schedule
  tf.train.AdamOptimizer(1e-6) -> TF-1 ApplyAdam functor (training_ops.h), beta powers kept in fp32

Tolerances for the CUDA paths are stated in tests/test_qnet_gpu.py.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn.functional as F

K1, K2, K3, FLAT = 256, 512, 576, 1600


def layout(hidden=512, dueling=False):
    names = [("w1", (8, 8, 4, 32)), ("b1", (32,)), ("w2", (4, 4, 32, 64)), ("b2", (64,)), ("w3", (3, 3, 64, 64)),
             ("b3", (64,)), ("wf1", (FLAT, hidden)), ("bf1", (hidden,))]
    if not dueling:
        names += [("wf2", (hidden, 2)), ("bf2", (2,))]
    else:
        names += [("wv", (hidden, 1)), ("bv", (1,)), ("wa", (hidden, 2)), ("ba", (2,))]
    out, o = {}, 0
    for n, shp in names:
        sz = int(np.prod(shp))
        out[n] = (o, shp)
        o += sz
    out["total"] = o
    return out


def init_params(hidden=512, dueling=False, seed=0) -> np.ndarray:
    """tf.truncated_normal(stddev=0.01) weights (resampled beyond 2 sigma), tf.constant(0.01) biases."""
    g = torch.Generator().manual_seed(seed)
    L = layout(hidden, dueling)
    flat = np.zeros(L["total"], np.float32)
    for n, v in L.items():
        if n == "total":
            continue
        o, shp = v
        sz = int(np.prod(shp))
        if n.startswith("b"):
            flat[o:o + sz] = 0.01
        else:
            w = torch.randn(sz, generator=g, dtype=torch.float64)
            bad = w.abs() > 2
            while bad.any():
                w[bad] = torch.randn(int(bad.sum()), generator=g, dtype=torch.float64)
                bad = w.abs() > 2
            flat[o:o + sz] = (w * 0.01).numpy().astype(np.float32)
    return flat


def _unpack(flat: torch.Tensor, hidden, dueling):
    L = layout(hidden, dueling)
    return {n: flat[v[0]:v[0] + int(np.prod(v[1]))].reshape(v[1]) for n, v in L.items() if n != "total"}


def _bf16(t):
    return t.to(torch.float32).to(torch.bfloat16).to(torch.float64)


def _fp16(t):
    return t.to(torch.float32).to(torch.float16).to(torch.float64)


def _tf32(t):
    """round-to-nearest-even to TF32 (10 explicit significand bits, fp32's exponent range): what a kind::tf32 operand keeps"""
    b = t.to(torch.float32).contiguous().view(torch.int32)
    b = (b + 0x0FFF + ((b >> 13) & 1)) & ~0x1FFF
    return b.view(torch.float32).to(torch.float64)


def _rounders(fmt, grad_scale=1.0):
    """(round a value to the operand format keeping the gradient, identity whose GRADIENT is rounded at `grad_scale`)"""
    if not fmt:
        return (lambda t: t), (lambda t: t)
    rnd = {"bf16": _bf16, "fp16": _fp16, "tf32": _tf32}[fmt]

    class RoundFwd(torch.autograd.Function):
        """value rounded to the 16-bit format, gradient passed through (an operand copy of an fp32 master tensor)"""

        @staticmethod
        def forward(ctx, t):
            return rnd(t)

        @staticmethod
        def backward(ctx, g):
            return g

    class RoundBwd(torch.autograd.Function):
        """identity whose GRADIENT is rounded (a gradient tensor stored in the 16-bit format between kernels, scaled by the
        power of two the fp16 path applies so that it stays in the format's normal range)"""

        @staticmethod
        def forward(ctx, t):
            return t.clone()

        @staticmethod
        def backward(ctx, g):
            return rnd(g * grad_scale) / grad_scale

    return RoundFwd.apply, RoundBwd.apply


def fp16_grad_scale(batch, loss_sum):
    """the power of two csrc/fb_qnet_tc.cu (train_step_launch) scales the gradient tensors by under FB_PRECISION_FP16"""
    s = 8.0
    if not loss_sum:
        while s < 8.0 * batch:
            s *= 2.0
    return s


def forward(flat: torch.Tensor, x_u8, hidden=512, dueling=False, return_all=False, emulate_bf16=False, emulate=None, grad_scale=1.0):
    """x_u8: [B,4,80,80] (channel = frame, oldest first; H = obs axis 0, W = obs axis 1) -> Q [B,2] float64.

    emulate_bf16: the same graph with the roundings of the tensor-core path (csrc/fb_qnet_tc.cu) put where that
    path has them -- conv / fc1 weights and the activations z1, a2, a3 rounded to bf16, and (in backward) the
    gradients dz1, dp1, dz2, dz3, dh1 rounded to bf16; sums stay exact.  Separates implementation errors
    (must be ~1e-3) from the precision of the format (a few percent)."""
    p = _unpack(flat, hidden, dueling)
    rw, rb = _rounders("bf16" if emulate_bf16 else emulate, grad_scale)     # emulate: None | "bf16" | "fp16"
    x = torch.as_tensor(np.asarray(x_u8)).to(torch.float64)                      # values 0.0 / 255.0, no normalisation
    z1 = rw(F.relu(rb(F.conv2d(x, rw(p["w1"]).permute(3, 2, 0, 1), p["b1"], stride=4, padding=2))))     # SAME: pad 2/2
    p1 = rb(F.max_pool2d(z1, 2, 2))
    a2 = rw(F.relu(rb(F.conv2d(p1, rw(p["w2"]).permute(3, 2, 0, 1), p["b2"], stride=2, padding=1))))    # SAME: pad 1/1
    a3 = rw(F.relu(rb(F.conv2d(a2, rw(p["w3"]).permute(3, 2, 0, 1), p["b3"], stride=1, padding=1))))
    flat3 = a3.permute(0, 2, 3, 1).reshape(-1, FLAT)                                           # tf.reshape of NHWC
    h1 = F.relu(rb(flat3 @ rw(p["wf1"]) + p["bf1"]))
    if not dueling:
        q = h1 @ p["wf2"] + p["bf2"]
    else:
        v = h1 @ p["wv"] + p["bv"]
        a = h1 @ p["wa"] + p["ba"]
        q = v + (a - a.mean(dim=1, keepdim=True))
    if return_all:
        return q, dict(z1=z1, p1=p1, a2=a2, a3=a3, h1=h1)
    return q


def loss_and_grads(variant, params32, target32, s, s2, actions, rewards, terminals, isw=None, gamma=0.99, loss_sum=False,
                   global_batch=None, hidden=512, dueling=False, emulate_bf16=False, isw_broadcast=False, emulate=None):
    """variant 0 vanilla / 1 nature / 2 double.  Returns loss, grads (float64 flat), abs_err, y (fp32 as fed), q(s).

    ``isw_broadcast``: the PER cost exactly as the reference's graph evaluates it -- ``ISWeights`` is a [B,1] placeholder and
    ``tf.square(q_target - q_eval)`` a [B] vector (BrainPrioritizedReplyDQN.py:243-251), so TensorFlow broadcasts the product
    to [B,B] and ``reduce_mean`` returns mean(w) * mean(err^2).  Default: the intended mean(w_i * err_i^2)."""
    P = torch.tensor(params32.astype(np.float64), requires_grad=True)
    T = torch.tensor((target32 if target32 is not None else params32).astype(np.float64))
    B = len(actions)
    gb = global_batch or B
    fmt = "bf16" if emulate_bf16 else emulate
    gs = fp16_grad_scale(gb, loss_sum) if fmt == "fp16" else 1.0
    with torch.no_grad():
        if variant == 0:
            x = forward(P.detach(), s2, hidden, dueling, emulate=fmt, grad_scale=gs).max(dim=1).values
        elif variant == 1:
            x = forward(T, s2, hidden, dueling, emulate=fmt, grad_scale=gs).max(dim=1).values
        else:
            qt = forward(T, s2, hidden, dueling, emulate=fmt, grad_scale=gs)
            am = forward(P.detach(), s2, hidden, dueling, emulate=fmt, grad_scale=gs).argmax(dim=1)
            x = qt[torch.arange(B), am]
        # the reference feeds fp32 Q-values into a Python float64 loop and feeds y back as fp32
        x32 = x.numpy().astype(np.float32).astype(np.float64)
        r = np.array([0.1 if abs(float(v) - 0.1) < 1e-6 else float(v) for v in rewards], np.float64)
        y = np.where(np.asarray(terminals).astype(bool), r, r + gamma * x32).astype(np.float32)
    q = forward(P, s, hidden, dueling, emulate=fmt, grad_scale=gs)
    onehot = F.one_hot(torch.as_tensor(np.asarray(actions).astype(np.int64)), 2).to(torch.float64)
    q_eval = (q * onehot).sum(dim=1)
    err = torch.as_tensor(y.astype(np.float64)) - q_eval
    w = torch.ones(B, dtype=torch.float64) if isw is None else torch.as_tensor(np.asarray(isw, np.float64))
    if isw is not None and isw_broadcast:
        prod = w.reshape(B, 1) * (err ** 2).reshape(1, B)          # [B,1] * [B] -> [B,B], what tf.multiply does
        loss = prod.sum() if loss_sum else prod.sum() / (B * gb)   # reduce_mean over B*B elements (gb = B on one GPU)
    else:
        loss = (w * err ** 2).sum() if loss_sum else (w * err ** 2).sum() / gb
    loss.backward()
    return float(loss.detach()), P.grad.numpy().copy(), err.detach().abs().numpy(), y, q.detach().numpy()


class AdamTF1:
    """TF-1.12 ApplyAdam in fp32: beta powers are fp32 variables multiplied every step."""

    def __init__(self, n, lr=1e-6, beta1=0.9, beta2=0.999, eps=1e-8):
        self.m = np.zeros(n, np.float32); self.v = np.zeros(n, np.float32)
        self.lr, self.b1, self.b2, self.eps = np.float32(lr), np.float32(beta1), np.float32(beta2), np.float32(eps)
        self.b1p, self.b2p = np.float32(beta1), np.float32(beta2)

    def alpha(self):
        return np.float32(self.lr * np.sqrt(np.float32(1) - self.b2p) / (np.float32(1) - self.b1p))

    def step(self, params32, grads32):
        g = grads32.astype(np.float32)
        a = self.alpha()
        self.m += (g - self.m) * (np.float32(1) - self.b1)
        self.v += (g * g - self.v) * (np.float32(1) - self.b2)
        params32 -= (self.m * a) / (np.sqrt(self.v) + self.eps)
        self.b1p = np.float32(self.b1p * self.b1); self.b2p = np.float32(self.b2p * self.b2)
        return params32


class WordStreamRandom(random.Random):
    """CPython's random.Random driven by an explicit 32-bit word stream (the same Philox words the device
    consumes), so that random(), randrange() and sample() follow CPython's own algorithms exactly."""

    def __init__(self, next_word):
        self._next_word = next_word
        super().__init__(0)

    def seed(self, *a, **k):
        return None

    def getrandbits(self, k):
        if k <= 0:
            raise ValueError
        out, shift = 0, 0
        while k > 0:                       # CPython fills 32-bit words, least significant first
            take = min(32, k)
            w = self._next_word() >> (32 - take)
            out |= w << shift
            shift += 32; k -= take
        return out

    def random(self):
        a = self._next_word() >> 5
        b = self._next_word() >> 6
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)


def egreedy_actions(q, epsilon, seed, first_env_id, rng_pos):
    """BrainDQN.py:102-108 per env; env e draws from Philox stream (seed, purpose 2, first_env_id+e)."""
    from . import flappy_oracle as fo
    acts = np.zeros(len(q), np.uint8)
    for e in range(len(q)):
        pos = [int(rng_pos[e])]

        def nxt(e=e, pos=pos):
            w = fo.stream_word(seed, 2, first_env_id + e, pos[0]); pos[0] += 1
            return w
        R = WordStreamRandom(nxt)
        if R.random() <= epsilon:
            acts[e] = R.randrange(2)
        else:
            acts[e] = int(np.argmax(q[e]))
        rng_pos[e] = pos[0]
    return acts
