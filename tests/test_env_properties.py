"""Property tests of the env path (layer 3): random action strings x random gap strings, the library against the
oracle.  `hypothesis` draws the strings (and shrinks a failure to a minimal one).

CPU part (not gpu): fb_debug_host_step / fb_debug_host_obs run the HOST build of exactly the functions the step kernel
runs (fb_env_logic.cuh is __host__ __device__), against the C oracle.  GPU part: the kernel itself through fb_env_step."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import flappy_oracle as fo

gap_strings = st.lists(st.integers(0, 7), min_size=1, max_size=37)
# action strings built from runs (long no-op / flap runs reach the ground, the ceiling and the pipes; i.i.d. bits rarely get past the first pipe)
runs = st.lists(st.tuples(st.integers(0, 1), st.integers(1, 40)), min_size=1, max_size=40)


def _expand(run_list, cap=700):
    a = np.concatenate([np.full(n, v, np.uint8) for v, n in run_list])[:cap]
    a[0] = 0                                   # the driver's first step is a no-op (FlappyBirdDQN.py:65-66)
    return a


@pytest.fixture(scope="module")
def L():
    from dqnflappybird_b200 import _lib
    from dqnflappybird_b200.assets import load_blob
    lib = _lib.lib()
    blob = load_blob()
    assert lib.fb_debug_assets_load_host(blob, len(blob)) == 0, lib.fb_last_error()
    return lib


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(gaps=gap_strings, run_list=runs)
def test_host_logic_equals_oracle_on_random_strings(L, gaps, run_list):
    gaps = np.array(gaps, np.uint8)
    acts = _expand(run_list)
    orc = fo.OracleEnvs(1, gaps=gaps[None, :])
    state = np.zeros(16, np.int32)
    assert L.fb_debug_host_reset(state.ctypes.data, gaps.ctypes.data, len(gaps), 0, 0) == 0
    r, t, s = C.c_float(), C.c_uint8(), C.c_int32()
    for k, a in enumerate(acts):
        assert L.fb_debug_host_step(state.ctypes.data, int(a), gaps.ctypes.data, len(gaps), 0, 0, C.byref(r), C.byref(t), C.byref(s)) == 0
        _, o_r, o_t, o_s = orc.step(np.array([a], np.uint8), want_obs=False)
        assert (r.value, t.value, s.value) == (o_r[0], o_t[0], o_s[0]), k
        np.testing.assert_array_equal(state, orc.export_state()[0], err_msg=f"step {k}")
    want = orc.obs(0)
    for mode in (0, 1):                        # table path and per-pixel path of the observation
        out = np.empty((80, 80), np.uint8)
        assert L.fb_debug_host_obs(state.ctypes.data, mode, out.ctypes.data) == 0
        np.testing.assert_array_equal(out, want)


@pytest.mark.gpu
@settings(max_examples=40, deadline=None)
@given(data=st.data())
def test_kernel_equals_oracle_on_random_strings(data):
    """the step kernel: 8 envs per example, each with its own gap string (padded by repetition to one length) and action string"""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dqnflappybird_b200 import _lib, game
    N = 8
    glen = data.draw(st.integers(1, 23))
    gaps = np.array([data.draw(st.lists(st.integers(0, 7), min_size=glen, max_size=glen)) for _ in range(N)], np.uint8)
    acts = [_expand(data.draw(runs), cap=300) for _ in range(N)]
    T = max(len(a) for a in acts)
    actions = np.zeros((T, N), np.uint8)
    for e, a in enumerate(acts):
        actions[:len(a), e] = a
    orc = fo.OracleEnvs(N, gaps=gaps)
    o_rew = np.zeros((T, N), np.float32); o_term = np.zeros((T, N), np.uint8); o_score = np.zeros((T, N), np.int32)
    o_obs = np.zeros((T, N, 80, 80), np.uint8)
    for t in range(T):
        o_obs[t], o_rew[t], o_term[t], o_score[t] = orc.step(actions[t], want_obs=True)
    gs = game.GameState(num_envs=N, replay_gaps=gaps, history=T)
    a = torch.from_numpy(actions).cuda()
    rew = torch.empty((T, N), dtype=torch.float32, device="cuda"); term = torch.empty((T, N), dtype=torch.uint8, device="cuda")
    score = torch.empty((T, N), dtype=torch.int32, device="cuda")
    _lib.check(gs._L.fb_env_step(gs._h, T, a.data_ptr(), gs.ring.data_ptr(), T, 0, rew.data_ptr(), term.data_ptr(), score.data_ptr(),
                                 game._stream_ptr(gs.device)), "fb_env_step")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(rew.cpu().numpy(), o_rew)
    np.testing.assert_array_equal(term.cpu().numpy(), o_term)
    np.testing.assert_array_equal(score.cpu().numpy(), o_score)
    np.testing.assert_array_equal(gs.export_state().cpu().numpy(), orc.export_state())
    np.testing.assert_array_equal(gs.ring.permute(1, 0, 2, 3).cpu().numpy(), o_obs)
