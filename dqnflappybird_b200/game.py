"""Batched, device-resident mirror of the reference environment.

Reference: ``game/wrapped_flappy_bird.py`` (``GameState`` :58-183) plus the
``preprocess`` of ``FlappyBirdDQN.py:31-34``.  One ``GameState(num_envs=N)`` is N
independent reference processes; ``frame_step`` does step + render + preprocess
for all of them in one kernel launch (csrc/fb_env.cu) and returns the 80x80
observations as a view of the device frame ring.

PyTorch only owns the tensors and the stream; all work goes through the C ABI
(include/flappy_b200.h).  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

STATE_FIELDS = ("playery", "playerVelY", "playerIndex", "loopIter", "cyclePhase", "basex", "score",
                "nPipes", "pipe0_x", "pipe1_x", "pipe2_x", "pipe0_gap", "pipe1_gap", "pipe2_gap", "rng_draws", "_")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class GameState:
    """``game.GameState`` for ``num_envs`` environments at once.

    Parameters mirror the reference where it has any (it has none: ``GameState()``,
    wrapped_flappy_bird.py:59); the rest configure the batch:

    num_envs     number of independent envs (each is one reference process)
    device       CUDA device
    seed         seed of the per-env Philox gap streams (``random.randint`` of :212)
    replay_gaps  optional u8[num_envs][G] logged gap indices (0..7): replay mode
    assets_dir   a reference-style ``assets`` directory; default: the packed sprites
                 shipped with the package
    history      L, slots per env in the frame ring ``u8[N][L][80][80]`` (>= 4)
    first_env_id global id of env 0 (sharding across GPUs keeps streams distinct)
    ring         optionally an existing ring tensor to draw into (e.g. a replay ring)
    """

    def __init__(self, num_envs: int = 1, device="cuda:0", seed: int = 0, replay_gaps=None,
                 assets_dir: str | None = None, history: int = 4, first_env_id: int = 0, ring: torch.Tensor | None = None):
        if not torch.cuda.is_available():
            raise _lib.FlappyError("GameState needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.FlappyError("GameState runs on CUDA devices only")
        self.num_envs = int(num_envs)
        self.history = int(history)
        if self.history < 1:
            raise ValueError("history must be >= 1")
        self._L = _lib.lib()
        torch.cuda.set_device(self.device)
        _lib.ensure_assets(assets_dir, self.device.index or 0)
        import ctypes as C
        h = C.c_void_p()
        _lib.check(self._L.fb_env_create(self.num_envs, seed, first_env_id, C.byref(h)), "fb_env_create")
        self._h = h
        N = self.num_envs
        if ring is None:
            ring = torch.zeros((N, self.history, 80, 80), dtype=torch.uint8, device=self.device)
        else:
            assert ring.shape == (N, self.history, 80, 80) and ring.dtype == torch.uint8 and ring.is_contiguous()
        self.ring = ring
        self.slot = self.history - 1          # slot of the most recent frame; first step writes slot 0
        self.steps = 0
        self.reward = torch.zeros(N, dtype=torch.float32, device=self.device)
        self.terminal = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self.score = torch.zeros(N, dtype=torch.int32, device=self.device)
        self._actions = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self._gaps = None
        if replay_gaps is not None:
            self.set_gap_replay(replay_gaps)

    # ------------------------------------------------------------------ lifecycle
    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.fb_env_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_gap_replay(self, gaps):
        """Replay mode: consume the logged gap sequence instead of the Philox stream, then reset."""
        g = torch.as_tensor(np.asarray(gaps, dtype=np.uint8) if not torch.is_tensor(gaps) else gaps)
        g = g.to(self.device, torch.uint8).contiguous()
        if g.dim() != 2 or g.shape[0] != self.num_envs:
            raise ValueError("replay_gaps must be u8[num_envs][G]")
        self._gaps = g
        _lib.check(self._L.fb_env_set_gap_replay(self._h, g.data_ptr(), g.shape[1]), "fb_env_set_gap_replay")
        self.reset()

    def reset(self):
        _lib.check(self._L.fb_env_reset(self._h, _stream_ptr(self.device)), "fb_env_reset")
        self.slot = self.history - 1
        self.steps = 0

    # ------------------------------------------------------------------ stepping
    def _to_action_index(self, input_actions) -> torch.Tensor:
        N = self.num_envs
        a = input_actions
        if torch.is_tensor(a) and a.is_cuda:
            if a.dim() == 2 and a.shape == (N, 2):
                # one-hot on the device: rows that do not sum to 1 become 2 -> device error flag
                ok = a.sum(dim=1) == 1
                a = torch.where(ok, (a[:, 1] != 0).to(torch.uint8), torch.full((), 2, dtype=torch.uint8, device=a.device))
            elif a.dim() == 1 and a.shape[0] == N:
                a = a.to(torch.uint8)
            else:
                raise ValueError("actions must be [N] indices or [N,2] one-hot")
            return a.contiguous()
        a = np.asarray(a)
        if a.ndim == 1 and a.shape[0] == 2 and N == 1:
            a = a[None, :]
        if a.ndim == 2 and a.shape == (N, 2):
            if (a.sum(axis=1) != 1).any():
                raise ValueError("Multiple input actions!")            # wrapped_flappy_bird.py:99-100
            a = (a[:, 1] == 1).astype(np.uint8)
        elif a.ndim == 1 and a.shape[0] == N:
            if ((a != 0) & (a != 1)).any():
                raise ValueError("Multiple input actions!")
            a = a.astype(np.uint8)
        else:
            raise ValueError("actions must be [N] indices, [N,2] one-hot, or one one-hot pair when num_envs == 1")
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device, non_blocking=True)

    def frame_step(self, input_actions, render_full: bool = False, out=None):
        """``GameState.frame_step`` (wrapped_flappy_bird.py:87-183) for every env.

        Returns ``(obs, reward, terminal, score)`` -- the reference's 4-tuple (:183) with the
        preprocessed 80x80 observation in place of ``image_data``:

        * batched call: ``obs`` u8[N,80,80] is a VIEW of the device ring (valid until the slot is
          reused ``history`` steps later), ``reward`` f32[N], ``terminal`` bool[N], ``score`` i32[N],
          all device tensors;
        * ``num_envs == 1`` with a single one-hot pair: numpy ``u8[80,80,1]`` (what
          ``preprocess`` returns, FlappyBirdDQN.py:34), ``float``, ``bool``, ``int`` -- or the raw
          ``u8[288,512,3]`` ``image_data`` when ``render_full=True``.

        ``out=(reward f32[N], terminal u8[N])`` makes the step kernel write those two straight into the caller's device
        tensors -- e.g. the replay rows ``Brain.next_rows()`` hands out, so that no transition field is ever copied;
        ``terminal`` is then returned as that u8 tensor.
        """
        single = (not torch.is_tensor(input_actions)) and np.asarray(input_actions).shape == (2,) and self.num_envs == 1
        a = self._to_action_index(input_actions)
        slot = (self.slot + 1) % self.history
        reward, terminal = (self.reward, self.terminal) if out is None else out
        if out is not None:
            N = self.num_envs
            assert reward.shape == (N,) and reward.dtype == torch.float32 and reward.is_cuda and reward.is_contiguous()
            assert terminal.shape == (N,) and terminal.dtype == torch.uint8 and terminal.is_cuda and terminal.is_contiguous()
        _lib.check(self._L.fb_env_step(self._h, 1, a.data_ptr(), self.ring.data_ptr(), self.history, slot,
                                       reward.data_ptr(), terminal.data_ptr(), self.score.data_ptr(),
                                       _stream_ptr(self.device)), "fb_env_step")
        self.slot = slot
        self.steps += 1
        obs = self.ring[:, slot]
        if single:
            if render_full:
                img = self.render_full(0, 1)[0].cpu().numpy()
            else:
                img = obs[0].cpu().numpy().reshape(80, 80, 1)
            return img, float(reward[0].item()), bool(terminal[0].item()), int(self.score[0].item())
        return obs, reward, (terminal.bool() if out is None else terminal), self.score

    def step_random(self, n_steps: int = 1, p_flap: float = 0.5, action_seed: int = 1234, actions_out: torch.Tensor | None = None,
                    reward: torch.Tensor | None = None, terminal: torch.Tensor | None = None, score: torch.Tensor | None = None,
                    draw: bool = True):
        """n_steps of frame_step with Bernoulli(p_flap) actions drawn on the device (config "random actions")."""
        thr = min(int(round(p_flap * 4294967296.0)), 4294967295)
        slot = (self.slot + 1) % self.history
        ptr = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self._L.fb_env_step_random(self._h, n_steps, action_seed, self.steps, thr, ptr(actions_out),
                                              self.ring.data_ptr() if draw else None, self.history, slot,
                                              ptr(reward), ptr(terminal), ptr(score), _stream_ptr(self.device)),
                   "fb_env_step_random")
        self.slot = (self.slot + n_steps) % self.history
        self.steps += n_steps

    def frame_step_host(self, actions_host: torch.Tensor, reward_host: torch.Tensor, terminal_host: torch.Tensor,
                        score_host: torch.Tensor):
        """The reference-facing call with (pinned) HOST buffers: H2D actions, step, D2H reward/terminal/score.
        The observation stays in the device ring; returns its view."""
        slot = (self.slot + 1) % self.history
        _lib.check(self._L.fb_env_step_host(self._h, actions_host.data_ptr(), self.ring.data_ptr(), self.history, slot,
                                            reward_host.data_ptr(), terminal_host.data_ptr(), score_host.data_ptr(),
                                            _stream_ptr(self.device)), "fb_env_step_host")
        self.slot = slot
        self.steps += 1
        return self.ring[:, slot]

    def frame_step_host_submit(self, actions_host: torch.Tensor, reward_host: torch.Tensor, terminal_host: torch.Tensor,
                               score_host: torch.Tensor):
        """frame_step_host without the final wait: up to two steps may be in flight, so the copies of one step overlap
        the kernel of the next.  The host buffers of a step are valid after the matching frame_step_host_wait()."""
        slot = (self.slot + 1) % self.history
        _lib.check(self._L.fb_env_step_host_submit(self._h, actions_host.data_ptr(), self.ring.data_ptr(), self.history, slot,
                                                   reward_host.data_ptr(), terminal_host.data_ptr(), score_host.data_ptr(),
                                                   _stream_ptr(self.device)), "fb_env_step_host_submit")
        self.slot = slot
        self.steps += 1
        return self.ring[:, slot]

    def frame_step_host_wait(self):
        """Block until the oldest submitted host step has delivered reward / terminal / score."""
        _lib.check(self._L.fb_env_step_host_wait(self._h), "fb_env_step_host_wait")

    def draw(self) -> torch.Tensor:
        """Draw the current state into the next ring slot without stepping; returns the obs view."""
        slot = (self.slot + 1) % self.history
        _lib.check(self._L.fb_env_draw(self._h, self.ring.data_ptr(), self.history, slot, _stream_ptr(self.device)), "fb_env_draw")
        self.slot = slot
        return self.ring[:, slot]

    def check_errors(self):
        """Raise ValueError('Multiple input actions!') if any device-side action was not one-hot."""
        _lib.check(self._L.fb_env_check(self._h, _stream_ptr(self.device)), "fb_env_check")

    # ------------------------------------------------------------------ views
    def stacked_state(self) -> torch.Tensor:
        """u8[N,80,80,4] copy of the last four frames, newest last (BrainDQN.py:68,239)."""
        L = self.history
        idx = [(self.slot - 3 + k) % L for k in range(4)]
        return self.ring[:, idx].permute(0, 2, 3, 1).contiguous()

    def export_state(self) -> torch.Tensor:
        out = torch.empty((self.num_envs, 16), dtype=torch.int32, device=self.device)
        _lib.check(self._L.fb_env_export_state(self._h, out.data_ptr(), _stream_ptr(self.device)), "fb_env_export_state")
        return out

    def import_state(self, state: torch.Tensor):
        s = state.to(self.device, torch.int32).contiguous()
        _lib.check(self._L.fb_env_import_state(self._h, s.data_ptr(), _stream_ptr(self.device)), "fb_env_import_state")
        self.check_errors()

    def obs_exact(self) -> torch.Tensor:
        """Observation of the current state by per-pixel arithmetic only (cross-check of the table path)."""
        out = torch.empty((self.num_envs, 80, 80), dtype=torch.uint8, device=self.device)
        _lib.check(self._L.fb_env_obs_exact(self._h, out.data_ptr(), _stream_ptr(self.device)), "fb_env_obs_exact")
        return out

    def render_full(self, first: int = 0, n: int | None = None) -> torch.Tensor:
        """``image_data`` of frame_step: u8[n,288,512,3] (surfarray.array3d layout, :177)."""
        n = self.num_envs - first if n is None else n
        out = torch.empty((n, 288, 512, 3), dtype=torch.uint8, device=self.device)
        _lib.check(self._L.fb_render_full(self._h, first, n, out.data_ptr(), _stream_ptr(self.device)), "fb_render_full")
        return out
