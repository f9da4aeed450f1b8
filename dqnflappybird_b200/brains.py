"""The reference's ``Brain*`` agents, batched over N envs and resident on the device.

Same surface as the reference (BrainDQN.py:32-239): ``Brain(actionNum, gameName)``, ``setInitState``,
``getAction``, ``setPerception(nextObserv, action, reward, terminal, curScore)``, ``_trainQNetwork`` /
``trainQNetwork``, and the attributes callers and logs read (``timeStep``, ``onlineTimeStep``, ``epsilon``,
``gameTimes``, ``replayMemory``, ``currentState``).  The module-level hyper-parameters of the reference
(BrainDQN.py:19-28) are constructor keywords with the reference's values as defaults.

    class                      reference file                      target            head     loss
    BrainDQN                   BrainDQN.py                         online net        plain    sum  (:162)
    BrainDQNNature             BrainDQNNature.py                   target net        plain    mean (:119)
    BrainDoubleDQN             BrainDoubleDQN.py:51-61             double            plain    mean
    BrainDuelingDQN            BrainDuelingDQN_CC.py:68-77         target net        dueling  mean
    BrainPrioritizedReplyDQN   BrainPrioritizedReplyDQN.py         target net        plain    mean of w*err^2, SumTree

Reference quirks are switchable: ``reference_quirks=True`` reproduces what the shipped code
really executes (Q1: ``--model ddqn`` runs the Nature update; Q2: ``--model duelingdqn`` builds the plain
head; Q3: the PER brain never syncs its target net); the default runs the intended algorithm.

All math runs in libflappy_b200.so; with torch.distributed initialised the replicated learner all-reduces
its flat gradient vector (NCCL) before the Adam step, every rank training on its own env/replay shard.
"""
from __future__ import annotations

import numpy as np
import torch

from .qnet import FrameBatch, QNetwork
from .replay import PrioritizedMemory, ReplayMemory

# Hyper Parameters (BrainDQN.py:19-28)
FRAME_PER_ACTION = 1
BATCH_SIZE = 32
OBSERVE = 1000.
EXPLORE = 1000000.
GAMMA = 0.99
FINAL_EPSILON = 0
INITIAL_EPSILON = 0.03
REPLAY_MEMORY = 50000
REPLACE_TARGET_ITER = 500


class BrainDQN:
    variant = "vanilla"
    dueling = False
    loss_sum = True                 # BrainDQN.py:162 reduce_sum; every subclass uses reduce_mean
    prioritized = False
    uses_target = False

    def __init__(self, actionNum: int = 2, gameName: str = "bird", num_envs: int = 1, device="cuda:0", ring: torch.Tensor | None = None,
                 batch_size: int = BATCH_SIZE, observe: float = OBSERVE, explore: float = EXPLORE, gamma: float = GAMMA,
                 final_epsilon: float = FINAL_EPSILON, initial_epsilon: float = INITIAL_EPSILON,
                 replay_memory_per_env: int | None = None, replace_target_iter: int | None = REPLACE_TARGET_ITER,
                 hidden: int = 512, lr: float = 1e-6, seed: int = 0, first_env_id: int = 0, updates_per_step: int = 1,
                 reference_quirks: bool = False, copy_target_at_init: bool = False, record: bool = False, max_act_batch: int = 1024,
                 precision: str = "bf16", peer_exchange: bool | None = None):
        if actionNum != 2:
            raise ValueError("the Flappy Bird hot path has two actions (FlappyBirdDQN.py:38)")
        self.actionNum, self.gameName = actionNum, gameName
        self.device = torch.device(device)
        self.num_envs = N = int(num_envs)
        self.batch_size, self.observe, self.explore, self.gamma = batch_size, observe, explore, gamma
        self.final_epsilon, self.initial_epsilon = final_epsilon, initial_epsilon
        self.replace_target_iter = replace_target_iter
        self.updates_per_step = updates_per_step
        self.reference_quirks = reference_quirks
        self.record = record
        self.seed, self.first_env_id = seed, first_env_id
        # distributed: one process per GPU, replicated learner
        self.world = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
        self.local_batch = max(1, batch_size // self.world)
        # init replay memory (BrainDQN.py:35): REPLAY_MEMORY transitions per env, frames shared with the env's ring
        C = replay_memory_per_env if replay_memory_per_env is not None else (REPLAY_MEMORY if N == 1 else 60)
        L = C + 4
        if ring is None:
            ring = torch.zeros((N, L, 80, 80), dtype=torch.uint8, device=self.device)
        assert ring.shape == (N, L, 80, 80), f"ring must be u8[{N}][{L}][80][80] (capacity {C} + 4 frames)"
        self.ring = ring
        mem_cls = PrioritizedMemory if self.prioritized else ReplayMemory
        self.replayMemory = mem_cls(ring, C, seed=seed + 1, max_batch=max(self.local_batch, 8))
        # init other parameters (BrainDQN.py:37-41)
        self.onlineTimeStep = 0
        self.gameTimes = 0
        self.timeStep = 0
        self.epsilon = initial_epsilon
        # init Q network (BrainDQN.py:60)
        dueling = self.dueling and not (reference_quirks and type(self).__name__ == "BrainDuelingDQN")
        self.net = QNetwork(self.device, hidden=hidden, dueling=dueling, max_batch=max(self.local_batch, min(N, max_act_batch)),
                            seed=seed, lr=lr, copy_target_at_init=copy_target_at_init, precision=precision)
        # one process per GPU under NCCL: sum the gradients from NVLink peer memory inside the Adam kernel (fb_dist.cu)
        if peer_exchange is None:
            peer_exchange = self.world > 1 and torch.distributed.get_backend() == "nccl"
        if peer_exchange and self.world > 1:
            self.net.enable_peer_exchange()
        self._k = 0                               # time index of the newest frame in the ring
        self._rng_pos = torch.zeros(N, dtype=torch.int32, device=self.device)
        self._actions = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self._q = torch.zeros((N, 2), dtype=torch.float32, device=self.device)
        self._abs_err = torch.zeros(max(self.local_batch, 8), dtype=torch.float32, device=self.device)
        self._q_target = torch.zeros(max(self.local_batch, 8), dtype=torch.float32, device=self.device)
        self._isw32 = torch.zeros(max(self.local_batch, 8), dtype=torch.float32, device=self.device)
        self._game_times = torch.zeros((), dtype=torch.int64, device=self.device)
        self.lost_hist, self.q_target_list = [], []

    # ------------------------------------------------------------------ state
    @property
    def slot(self) -> int:
        return self._k % self.ring.shape[1]

    @property
    def currentState(self) -> torch.Tensor:
        """u8[N,80,80,4] copy of the last four frames, newest last (BrainDQN.py:68,239)"""
        L = self.ring.shape[1]
        idx = [max(self._k - 3 + c, 0) % L for c in range(4)]
        return self.ring[:, idx].permute(0, 2, 3, 1).contiguous()

    def _to_ring(self, observ, slot):
        dst = self.ring[:, slot]
        if torch.is_tensor(observ) and observ.is_cuda and observ.data_ptr() == dst.data_ptr():
            return                                 # the env drew straight into our ring
        o = torch.as_tensor(np.asarray(observ) if not torch.is_tensor(observ) else observ)
        dst.copy_(o.reshape(self.num_envs, 80, 80).to(self.device, torch.uint8))

    def setInitState(self, observ):
        """currentState = np.stack((observ,)*4, axis=2) (BrainDQN.py:238-239): frame 0; earlier times clamp to it"""
        self._k = 0
        self._to_ring(observ, 0)
        self.replayMemory.t = 0

    def _act_view(self) -> FrameBatch:
        L = self.ring.shape[1]
        off = [(max(self._k - 3 + c, 0) % L) * 6400 for c in range(4)]
        return FrameBatch(self.ring, L * 6400, off, self.num_envs)

    # ------------------------------------------------------------------ acting
    def getAction(self):
        """BrainDQN.py:99-116.  Returns u8[N] action indices on the device (0 = [1,0] no-op, 1 = [0,1] flap); for a
        single env the reference's one-hot float array."""
        self.net.act(self._act_view(), self.epsilon, self.seed + 2, self.first_env_id, self._rng_pos, self._actions, self._q)
        # change epsilon (BrainDQN.py:112-114), float64 like the reference
        if self.epsilon > self.final_epsilon and self.onlineTimeStep > self.observe:
            self.epsilon -= (self.initial_epsilon - self.final_epsilon) / self.explore
        if self.num_envs == 1:
            action = np.zeros(self.actionNum)
            action[int(self._actions[0].item())] = 1
            return action
        return self._actions

    # ------------------------------------------------------------------ perception
    def setPerception(self, nextObserv, action, reward, terminal, curScore=None):
        """BrainDQN.py:66-96 for all envs: append the transition, train once past OBSERVE, advance the counters."""
        k = self._k + 1
        L = self.ring.shape[1]
        self._to_ring(nextObserv, k % L)
        a_row, r_row, t_row = self.replayMemory.rows(k)
        dev = self.device
        a = torch.as_tensor(np.asarray(action) if not torch.is_tensor(action) else action)
        if a.dim() >= 1 and a.shape[-1] == 2 and (a.dim() == 2 or self.num_envs == 1):
            a = a.reshape(-1, 2)[:, 1]               # one-hot -> index
        a_row.copy_(a.reshape(-1).to(dev).to(torch.uint8))
        r_row.copy_(torch.as_tensor(reward, dtype=torch.float32).reshape(-1).to(dev))
        t_row.copy_(torch.as_tensor(terminal).reshape(-1).to(dev).to(torch.uint8))
        self._k = k
        self.replayMemory.appended(k)               # deque.append / Memory.store
        if self.onlineTimeStep > self.observe:
            for _ in range(self.updates_per_step):
                self._trainQNetwork()
        self._game_times += t_row.sum()              # gameTimes += 1 per terminal (BrainDQN.py:88-90), kept on the device
        self.timeStep += 1
        self.onlineTimeStep += 1

    @property
    def gameTimesTotal(self) -> int:
        return int(self._game_times.item())

    # ------------------------------------------------------------------ training
    def _maybe_sync_target(self):
        if self.uses_target and self.replace_target_iter and self.timeStep % self.replace_target_iter == 0:
            self.net.sync_target()                   # BrainDQNNature.py:151-152

    def _update(self, variant: str):
        mem = self.replayMemory
        mb = mem.sample(self.local_batch)
        isw = None
        if mb.is_weights is not None:
            isw = self._isw32[:self.local_batch]
            isw.copy_(mb.is_weights)                 # the placeholder is tf.float32 (BrainPrioritizedReplyDQN.py:243)
        self.net.loss_backward(variant, mb.frames, mb.actions, mb.rewards, mb.terminals, isw, self.gamma, self.loss_sum,
                               self.local_batch * self.world, self._abs_err[:self.local_batch], self._q_target[:self.local_batch])
        if self.world > 1 and self.net.exchange is None:
            torch.distributed.all_reduce(self.net.grads)          # sum of per-shard gradients of the global loss
        self.net.adam_step()                                      # with a peer exchange the sum happens inside the Adam kernel
        if mb.tree_idx is not None:
            mem.batch_update(mb.tree_idx, abs_errors=self._abs_err[:self.local_batch])   # :316
        if self.record:
            self.lost_hist.append(float(self.net.loss.item()))
            self.q_target_list.append(self._q_target[:self.local_batch].cpu().tolist())

    def _trainQNetwork(self):
        """BrainDQN.py:195-223"""
        self._update("vanilla")

    # ------------------------------------------------------------------ checkpoint (row N2)
    def state_dict(self):
        return {"net": self.net.state_dict(), "gameTimes": self.gameTimesTotal, "timeStep": self.timeStep, "epsilon": self.epsilon}

    def load_state_dict(self, sd):
        """_load_saved_parameters (BrainDQN.py:176-192): timeStep and epsilon are restored, onlineTimeStep is not (Q12)"""
        self.net.load_state_dict(sd["net"])
        self._game_times.fill_(int(sd["gameTimes"]))
        self.timeStep, self.epsilon = int(sd["timeStep"]), float(sd["epsilon"])


class BrainDQNNature(BrainDQN):
    variant = "nature"
    loss_sum = False
    uses_target = True

    def _trainQNetwork(self):
        """BrainDQNNature.py:149-183"""
        self._maybe_sync_target()
        self._update("nature")


class BrainDoubleDQN(BrainDQNNature):
    variant = "double"

    def trainQNetwork(self):
        """BrainDoubleDQN.py:37-69 -- the Double target the file defines but the shipped loop never calls (Q1)"""
        self._maybe_sync_target()
        self._update("double")

    def _trainQNetwork(self):
        if self.reference_quirks:
            return BrainDQNNature._trainQNetwork(self)
        return self.trainQNetwork()


class BrainDuelingDQN(BrainDQNNature):
    dueling = True

    def trainQNetwork(self):
        """BrainDuelingDQN_CC.py:171-202 (Nature target on the dueling net)"""
        return BrainDQNNature._trainQNetwork(self)


class BrainPrioritizedReplyDQN(BrainDQNNature):
    prioritized = True

    def _trainQNetwork(self):
        """BrainPrioritizedReplyDQN.py:277-315; the reference never runs target_replace_op here (Q3)"""
        if not self.reference_quirks:
            self._maybe_sync_target()
        self._update("nature")


MODELS = {"dqn": BrainDQN, "ddqn": BrainDoubleDQN, "dqnnature": BrainDQNNature, "duelingdqn": BrainDuelingDQN,
          "prioritydqn": BrainPrioritizedReplyDQN}       # FlappyBirdDQN.py:41-50
