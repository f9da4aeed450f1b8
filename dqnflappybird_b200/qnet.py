"""Host-side owner of the Q-network tensors; all math runs in libflappy_b200.so (csrc/fb_qnet*.cu).

Reference: the TensorFlow graph every Brain builds (BrainDQN.py:119-163; target copy
BrainDQNNature.py:75-111; dueling head BrainDuelingDQN_CC.py:68-77) and
``tf.train.AdamOptimizer(1e-6)`` (BrainDQN.py:163).  PyTorch owns the flat fp32 vectors
(parameters, target parameters, gradients, Adam slots); nothing here computes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

VARIANTS = {"vanilla": 0, "nature": 1, "double": 2}
# byte offsets of the four frames of s and of s' inside one minibatch sample u8[5][80][80]
_OFF_S = (C.c_int32 * 4)(0, 6400, 12800, 19200)
_OFF_N = (C.c_int32 * 4)(6400, 12800, 19200, 25600)
PRECISIONS = {"fp32": 0, "bf16": 1, "fp16": 2}


def truncated_normal_(t: torch.Tensor, std: float, generator=None):
    """tf.truncated_normal: N(0, std) resampled beyond two standard deviations (BrainDQN.py:122)."""
    torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=generator)
    return t


class FrameBatch:
    """A view of u8 frames as network inputs: sample b, channel c = base + b*stride + chan_off[c]."""

    def __init__(self, tensor: torch.Tensor, sample_stride: int, chan_off, batch: int, base_offset: int = 0):
        self.tensor = tensor                      # keeps the storage alive
        self.ptr = tensor.data_ptr() + base_offset
        self.sample_stride = int(sample_stride)
        self.chan_off = (C.c_int32 * 4)(*[int(o) for o in chan_off])
        self.batch = int(batch)

    @staticmethod
    def from_ring(ring: torch.Tensor, newest_slot: int) -> "FrameBatch":
        """Acting input: the last four frames of every env in a ring u8[N][L][80][80], newest last."""
        N, L = ring.shape[0], ring.shape[1]
        off = [((newest_slot - 3 + k) % L) * 6400 for k in range(4)]
        return FrameBatch(ring, L * 6400, off, N)

    @staticmethod
    def from_stack(frames: torch.Tensor, first: int = 0) -> "FrameBatch":
        """frames u8[B][F][80][80] with F >= first+4: channels = frames first..first+3."""
        B, Fr = frames.shape[0], frames.shape[1]
        assert frames.is_contiguous() and frames.dtype == torch.uint8 and Fr >= first + 4
        return FrameBatch(frames, Fr * 6400, [(first + k) * 6400 for k in range(4)], B)


class QNetwork:
    def __init__(self, device="cuda:0", hidden: int = 512, dueling: bool = False, max_batch: int = 256, seed: int = 0,
                 lr: float = 1e-6, beta1: float = 0.9, beta2: float = 0.999, adam_eps: float = 1e-8,
                 copy_target_at_init: bool = False, precision: str = "fp16"):
        if not torch.cuda.is_available():
            raise _lib.FlappyError("QNetwork needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self._L = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._L.fb_qnet_create(hidden, int(dueling), max_batch, C.byref(h)), "fb_qnet_create")
        self._h = h
        self.hidden, self.dueling, self.max_batch = hidden, bool(dueling), max_batch
        # "bf16" / "fp16": tcgen05 tensor cores (16-bit operands, fp32 accumulation; fp16 has TF32's significand); "fp32": strict CUDA-core FMA
        self.precision = precision
        _lib.check(self._L.fb_qnet_set_precision(self._h, PRECISIONS[precision]), "fb_qnet_set_precision")
        self._seen_versions = None
        self.compute_path = {"bf16": "TMA + tcgen05 implicit GEMM (bf16 operands, fp32 accumulate in TMEM)",
                             "fp16": "TMA + tcgen05 implicit GEMM (fp16 operands = TF32's 11-bit significand, fp32 accumulate in TMEM, "
                                     "power-of-two gradient scaling)",
                             "fp32": "fp32 CUDA-core implicit GEMM"}[precision]
        lay = (C.c_int32 * 16)()
        _lib.check(self._L.fb_qnet_layout(self._h, lay), "fb_qnet_layout")
        names = ["w1", "b1", "w2", "b2", "w3", "b3", "wf1", "bf1", "wf2", "bf2", "wv", "bv", "wa", "ba"]
        self.offsets = {n: int(lay[i]) for i, n in enumerate(names) if lay[i] >= 0}
        self.n_params = int(lay[14])
        z = lambda: torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        self.params, self.target, self.grads, self.adam_m, self.adam_v = z(), z(), z(), z(), z()
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        # TF variable initialisers: truncated_normal(0.01) weights, constant 0.01 biases (BrainDQN.py:122-123);
        # the target net is initialised INDEPENDENTLY (BrainDQNNature.py:75-102) unless asked otherwise
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.params.copy_(self._init_flat(g))
        self.target.copy_(self.params if copy_target_at_init else self._init_flat(g))
        self.lr, self.beta1, self.beta2, self.adam_eps = np.float32(lr), np.float32(beta1), np.float32(beta2), np.float32(adam_eps)
        self.beta1_power, self.beta2_power = np.float32(beta1), np.float32(beta2)      # TF keeps these as fp32 variables
        self.adam_steps = 0
        self.exchange = None                    # dist.PeerGradExchange once enable_peer_exchange() was called
        self.exchange_in_step = False

    def set_per_broadcast(self, on: bool):
        """PER loss exactly as the reference's graph evaluates it (BrainPrioritizedReplyDQN.py:243-251): the [B,1] ISWeights
        placeholder times the [B] squared error broadcasts to [B,B], i.e. every sample is weighted by mean(ISWeights)."""
        _lib.check(self._L.fb_qnet_set_per_broadcast(self._h, int(bool(on))), "fb_qnet_set_per_broadcast")

    def enable_peer_exchange(self, exchange=None):
        """Multi-GPU: keep the gradient vector in an NVLink-mapped exchange buffer and let adam_step() sum every rank's
        gradients from peer memory inside the Adam kernel (csrc/fb_dist.cu) instead of a separate all-reduce."""
        from .dist import PeerGradExchange
        self.exchange = exchange if exchange is not None else PeerGradExchange(self.n_params, self.device)
        self.grads = self.exchange.grads
        # tensor-core path: the exchange rides INSIDE the training step's graph (fb_dist.cu buckets) instead of after it
        self.exchange_in_step = self.precision != "fp32" and self.exchange.world > 1 and not getattr(self.exchange, "no_wait", False)
        _lib.check(self._L.fb_qnet_attach_exchange(self._h, self.exchange._h if self.exchange_in_step else None), "fb_qnet_attach_exchange")
        return self.exchange

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.fb_qnet_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _sizes(self):
        order = sorted(self.offsets.items(), key=lambda kv: kv[1])
        ends = [o for _, o in order[1:]] + [self.n_params]
        return [(n, o, e - o) for (n, o), e in zip(order, ends)]

    def _init_flat(self, gen) -> torch.Tensor:
        flat = torch.empty(self.n_params, dtype=torch.float32)
        for n, o, sz in self._sizes():
            if n.startswith("b"):
                flat[o:o + sz] = 0.01
            else:
                truncated_normal_(flat[o:o + sz], 0.01, gen)
        return flat

    def view(self, name: str, which: str = "params") -> torch.Tensor:
        """Named slice of a flat vector (which in params/target/grads/adam_m/adam_v), TF shapes."""
        shapes = {"w1": (8, 8, 4, 32), "b1": (32,), "w2": (4, 4, 32, 64), "b2": (64,), "w3": (3, 3, 64, 64), "b3": (64,),
                  "wf1": (1600, self.hidden), "bf1": (self.hidden,), "wf2": (self.hidden, 2), "bf2": (2,),
                  "wv": (self.hidden, 1), "bv": (1,), "wa": (self.hidden, 2), "ba": (2,)}
        o = self.offsets[name]
        shp = shapes[name]
        return getattr(self, which)[o:o + int(np.prod(shp))].view(shp)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _sync_versions(self):
        """The library refreshes its bf16 operand copies after its own Adam / target-sync; in-place writes made
        through torch (tests, load_state_dict, an all-reduce of the parameters) are noticed here."""
        v = (self.params._version, self.target._version)
        if v != self._seen_versions:
            _lib.check(self._L.fb_qnet_invalidate(self._h), "fb_qnet_invalidate")
            self._seen_versions = v

    # ---------------------------------------------------------------- forward / act
    def forward(self, fb: FrameBatch, target: bool = False, out: torch.Tensor | None = None) -> torch.Tensor:
        """QValue.eval(feed_dict={stateInput: ...}) (BrainDQN.py:100): f32[B][2]."""
        q = out if out is not None else torch.empty((fb.batch, 2), dtype=torch.float32, device=self.device)
        p = self.target if target else self.params
        self._sync_versions()
        _lib.check(self._L.fb_qnet_forward(self._h, p.data_ptr(), fb.ptr, fb.sample_stride, fb.chan_off, fb.batch,
                                           q.data_ptr(), self._stream()), "fb_qnet_forward")
        return q

    def act(self, fb: FrameBatch, epsilon: float, seed: int, first_env_id: int, rng_pos: torch.Tensor,
            actions_out: torch.Tensor, q_out: torch.Tensor):
        """getAction (BrainDQN.py:99-108) for every env of the batch."""
        self._sync_versions()
        _lib.check(self._L.fb_qnet_act(self._h, self.params.data_ptr(), fb.ptr, fb.sample_stride, fb.chan_off, fb.batch,
                                       float(epsilon), seed, first_env_id, rng_pos.data_ptr(), q_out.data_ptr(),
                                       actions_out.data_ptr(), self._stream()), "fb_qnet_act")
        return actions_out

    # ---------------------------------------------------------------- training
    def loss_backward(self, variant: str, frames: torch.Tensor, actions: torch.Tensor, rewards: torch.Tensor,
                      terminals: torch.Tensor, is_weights: torch.Tensor | None = None, gamma: float = 0.99,
                      loss_sum: bool = False, global_batch: int | None = None, abs_err: torch.Tensor | None = None,
                      q_target: torch.Tensor | None = None, sampling=None) -> torch.Tensor:
        """frames u8[B][5][80][80]: s = frames 0..3, s' = frames 1..4 of each sample.  Fills self.grads, self.loss.
        ``sampling``: as in train_step -- the minibatch is drawn by the first two kernels of the same graph."""
        B = frames.shape[0]
        assert frames.shape[1:] == (5, 80, 80) and frames.dtype == torch.uint8 and frames.is_contiguous()
        off_s, off_n = _OFF_S, _OFF_N
        ptr = lambda t: t.data_ptr() if t is not None else None
        self._sync_versions()
        if sampling is not None:
            assert frames.data_ptr() == sampling.frames_out_dev and B == sampling.batch and (is_weights is None) == (not sampling.prioritized)
            _lib.check(self._L.fb_qnet_train_step_sampled(
                self._h, C.addressof(sampling), VARIANTS[variant], self.params.data_ptr(), self.target.data_ptr(), off_s, off_n,
                global_batch or B, float(gamma), int(loss_sum), self.grads.data_ptr(), self.loss.data_ptr(), ptr(abs_err), ptr(q_target),
                None, None, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, self._stream()), "fb_qnet_train_step_sampled")
            return self.loss
        _lib.check(self._L.fb_qnet_loss_backward(
            self._h, VARIANTS[variant], self.params.data_ptr(), self.target.data_ptr(), frames.data_ptr(), 5 * 6400,
            off_s, off_n, actions.data_ptr(), rewards.data_ptr(), terminals.data_ptr(), ptr(is_weights), B,
            global_batch or B, float(gamma), int(loss_sum), self.grads.data_ptr(), self.loss.data_ptr(), ptr(abs_err),
            ptr(q_target), self._stream()), "fb_qnet_loss_backward")
        return self.loss

    def train_step(self, variant: str, frames: torch.Tensor, actions: torch.Tensor, rewards: torch.Tensor, terminals: torch.Tensor,
                   is_weights: torch.Tensor | None = None, gamma: float = 0.99, loss_sum: bool = False, global_batch: int | None = None,
                   abs_err: torch.Tensor | None = None, q_target: torch.Tensor | None = None, grad_scale: float = 1.0,
                   sampling=None) -> torch.Tensor:
        """session.run(trainStep) (BrainDQN.py:204-207): loss_backward + adam_step as one library call, bit-identical to the
        two; on the tensor-core path one CUDA graph launch with Adam inside the step's last kernel.  With a peer exchange
        (several GPUs) the sum over ranks lives in the Adam kernel, so the two calls are made separately.

        ``sampling`` (``ReplayMemory.step_sampling(batch)[0]``): the minibatch is drawn and gathered by the step's first two
        kernels -- ``random.sample`` and the list comprehensions of BrainDQN.py:197-201 ride in the same graph; ``frames`` /
        ``actions`` / ``rewards`` / ``terminals`` must then be the replay's own minibatch buffers."""
        in_step = self.exchange is None or self.exchange_in_step
        if sampling is not None and in_step:
            assert frames.data_ptr() == sampling.frames_out_dev and frames.shape[0] == sampling.batch
            assert (is_weights is None) == (not sampling.prioritized)
            off_s, off_n = _OFF_S, _OFF_N
            ptr = lambda t: t.data_ptr() if t is not None else None
            self._sync_versions()
            _lib.check(self._L.fb_qnet_train_step_sampled(
                self._h, C.addressof(sampling), VARIANTS[variant], self.params.data_ptr(), self.target.data_ptr(), off_s, off_n,
                global_batch or sampling.batch, float(gamma), int(loss_sum), self.grads.data_ptr(), self.loss.data_ptr(), ptr(abs_err),
                ptr(q_target), self.adam_m.data_ptr(), self.adam_v.data_ptr(), float(self.lr), float(self.beta1), float(self.beta2),
                float(self.adam_eps), float(grad_scale), float(self.beta1_power), float(self.beta2_power), self._stream()),
                "fb_qnet_train_step_sampled")
            self._after_fused_step()
            return self.loss
        if sampling is not None:                 # not fusable here: draw the minibatch with its own two launches
            assert not sampling.prioritized, "with a peer exchange call PrioritizedMemory.sample / batch_update yourself"
            _lib.check(self._L.fb_replay_sample_uniform(sampling.replay, sampling.t, sampling.batch, sampling.setsize, sampling.seed,
                                                        sampling.idx_out_dev, self._stream()), "fb_replay_sample_uniform")
            _lib.check(self._L.fb_replay_gather(sampling.replay, sampling.ring_dev, sampling.act_dev, sampling.rew_dev, sampling.term_dev,
                                                sampling.t, 0, sampling.idx_out_dev, sampling.batch, sampling.frames_out_dev,
                                                sampling.act_out_dev, sampling.rew_out_dev, sampling.term_out_dev, sampling.env_out_dev,
                                                sampling.k_out_dev, self._stream()), "fb_replay_gather")
        if self.exchange is not None and not self.exchange_in_step:
            self.loss_backward(variant, frames, actions, rewards, terminals, is_weights, gamma, loss_sum, global_batch, abs_err, q_target)
            self.adam_step(grad_scale)
            return self.loss
        B = frames.shape[0]
        assert frames.shape[1:] == (5, 80, 80) and frames.dtype == torch.uint8 and frames.is_contiguous()
        off_s, off_n = _OFF_S, _OFF_N
        ptr = lambda t: t.data_ptr() if t is not None else None
        self._sync_versions()
        _lib.check(self._L.fb_qnet_train_step(
            self._h, VARIANTS[variant], self.params.data_ptr(), self.target.data_ptr(), frames.data_ptr(), 5 * 6400,
            off_s, off_n, actions.data_ptr(), rewards.data_ptr(), terminals.data_ptr(), ptr(is_weights), B,
            global_batch or B, float(gamma), int(loss_sum), self.grads.data_ptr(), self.loss.data_ptr(), ptr(abs_err),
            ptr(q_target), self.adam_m.data_ptr(), self.adam_v.data_ptr(), float(self.lr), float(self.beta1), float(self.beta2),
            float(self.adam_eps), float(grad_scale), float(self.beta1_power), float(self.beta2_power), self._stream()),
            "fb_qnet_train_step")
        self._after_fused_step()
        return self.loss

    def _after_fused_step(self):
        self._advance_powers()
        if self.exchange_in_step:                # the step's graph carried the exchange: switch to the other exchange buffer
            _lib.check(self._L.fb_dist_advance(self.exchange._h), "fb_dist_advance")
            self.grads = self.exchange.grads

    def _advance_powers(self):
        """beta1_power *= beta1, beta2_power *= beta2 in fp32, like the update ops TF-1's Adam runs after every step"""
        self.beta1_power = np.float32(self.beta1_power * self.beta1)
        self.beta2_power = np.float32(self.beta2_power * self.beta2)
        self.adam_steps += 1

    def adam_step(self, grad_scale: float = 1.0):
        """One tf.train.AdamOptimizer step on self.params with self.grads (TF-1 ApplyAdam)."""
        one = np.float32(1)
        alpha = np.float32(self.lr * np.sqrt(one - self.beta2_power) / (one - self.beta1_power))
        if self.exchange is not None:           # sum over ranks + Adam in one kernel; then switch to the other buffer
            self.exchange.adam(self, alpha, grad_scale, wait=self.exchange.world > 1 and not getattr(self.exchange, "no_wait", False))
            self.grads = self.exchange.grads
        else:
            _lib.check(self._L.fb_qnet_adam(self._h, self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                                            self.adam_v.data_ptr(), float(alpha), float(self.beta1), float(self.beta2),
                                            float(self.adam_eps), float(grad_scale), self._stream()), "fb_qnet_adam")
        self._advance_powers()

    def sync_target(self):
        """sess.run(target_replace_op) (BrainDQNNature.py:107-111,151-152)."""
        _lib.check(self._L.fb_qnet_sync_target(self._h, self.target.data_ptr(), self.params.data_ptr(), self._stream()),
                   "fb_qnet_sync_target")

    # ---------------------------------------------------------------- checkpoint (row N2, weights + Adam slots + powers)
    def state_dict(self):
        return {"params": self.params.clone(), "target": self.target.clone(), "adam_m": self.adam_m.clone(),
                "adam_v": self.adam_v.clone(), "beta1_power": float(self.beta1_power), "beta2_power": float(self.beta2_power),
                "adam_steps": self.adam_steps, "hidden": self.hidden, "dueling": self.dueling}

    def load_state_dict(self, sd):
        assert sd["hidden"] == self.hidden and sd["dueling"] == self.dueling
        for k in ("params", "target", "adam_m", "adam_v"):
            getattr(self, k).copy_(sd[k])
        self.beta1_power, self.beta2_power = np.float32(sd["beta1_power"]), np.float32(sd["beta2_power"])
        self.adam_steps = int(sd["adam_steps"])
