"""The oracle (oracle/flappy_oracle.c) pinned against the REAL reference.

Fixtures come from the reference's own game/ modules run verbatim
(tests/golden/make_golden.py) and from the run logs the reference ships.
"""
import os

import numpy as np
import pytest

from oracle import flappy_oracle as fo


def _traj(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_env_trajectories.npz"))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(v) for v in fo.philox4x32_10(ctr, key)] == want


def test_oracle_matches_reference_trajectories(golden_dir):
    g = _traj(golden_dir)
    for ti in range(int(g["n_traj"])):
        gaps = g[f"t{ti}_gaps"]
        acts = g[f"t{ti}_actions"]
        env = fo.OracleEnvs(1, gaps=gaps[None, :])
        frames = dict(zip(g[f"t{ti}_frame_idx"].tolist(), g[f"t{ti}_frames"]))
        for t in range(len(acts)):
            obs, r, term, sc = env.step(acts[t:t + 1])
            assert r[0] == g[f"t{ti}_reward"][t], (ti, t)
            assert term[0] == g[f"t{ti}_terminal"][t], (ti, t)
            assert sc[0] == g[f"t{ti}_score"][t], (ti, t)
            st = env.export_state()[0]
            ref = g[f"t{ti}_state"][t].copy()
            ref[4] = st[4]                       # cyclePhase is not observable in the reference object
            np.testing.assert_array_equal(st, ref, err_msg=f"traj {ti} step {t}")
            np.testing.assert_array_equal(np.packbits(obs[0] > 0), g[f"t{ti}_obsbits"][t], err_msg=f"obs traj {ti} step {t}")
            if t in frames:
                np.testing.assert_array_equal(env.render_full(0), frames[t], err_msg=f"frame traj {ti} step {t}")


def test_oracle_matches_long_reference_trajectory(golden_dir):
    """42,000 steps of the reference itself with a competent controller (tests/golden/make_golden.py envlong): 1,063 scoring
    events, 56 crashes into lower pipes / upper pipes / the ground, hundreds of spawns and pops -- every step's reward, terminal,
    score and state, and the 9,314 kept observations"""
    g = np.load(os.path.join(golden_dir, "ref_env_long.npz"))
    acts = g["actions"]
    assert int((g["reward"] == 3).sum()) >= 1000 and int(g["terminal"].sum()) >= 50
    env = fo.OracleEnvs(1, gaps=g["gaps"][None, :])
    want = np.zeros(len(acts), np.uint8); want[g["obs_idx"]] = 1
    r, term, sc, st, obs = env.run(0, acts, want)
    np.testing.assert_array_equal(r, g["reward"]); np.testing.assert_array_equal(term, g["terminal"]); np.testing.assert_array_equal(sc, g["score"])
    ref = g["state"].copy(); ref[:, 4] = st[:, 4]        # cyclePhase is not observable in the reference object
    np.testing.assert_array_equal(st, ref)
    np.testing.assert_array_equal(np.packbits(obs.reshape(len(obs), -1) > 0, axis=1), g["obsbits"])
    # crash kinds: the bird's y before the fatal step separates ground hits (y >= 360) from lower-pipe hits (below the gap:
    # y + 24 > gapY + 100 >= 200) and upper-pipe hits (y < gapY <= 170)
    yb = g["crash_y_before"]
    assert (yb >= 355).sum() >= 3 and ((yb >= 180) & (yb < 355)).sum() >= 5 and (yb < 176).sum() >= 10, sorted(yb.tolist())


def test_known_episode_lengths():
    # all-no-op dies on step 19, all-flap on step 50, for any gaps
    for gap in range(8):
        gaps = np.full((1, 4), gap, np.uint8)
        for action, want in ((0, 19), (1, 50)):
            env = fo.OracleEnvs(1, gaps=gaps)
            for t in range(1, 200):
                _, r, term, _ = env.step(np.array([action], np.uint8), want_obs=False)
                if term[0]:
                    assert r[0] == -3
                    break
            # the first frame_step of a process is the same as any other, so count from 1
            assert t == want, (gap, action, t)


@pytest.mark.parametrize("name", ["dqn", "ddqn", "dqnnature", "duelingdqn", "prioritydqn"])
def test_oracle_reproduces_reference_logs(golden_dir, name):
    """Every logged (ACTION -> REWARD[, SCORE]) of the reference's real runs.

    The driver takes one unlogged no-op step first (FlappyBirdDQN.py:65-66).
    All crashes in the logs are gap independent (ground / top of upper pipe),
    so any gap script reproduces them; we check that claim with two scripts.
    """
    g = np.load(os.path.join(golden_dir, "ref_logs.npz"))
    acts, rews, scores = g[name + "_action"], g[name + "_reward"], g[name + "_score"]
    for gap_script in (np.zeros((1, 3), np.uint8), np.array([[7, 3, 5, 1, 0, 6]], np.uint8)):
        env = fo.OracleEnvs(1, gaps=gap_script)
        env.step(np.array([0], np.uint8), want_obs=False)
        ends = []
        for t in range(len(acts)):
            _, r, term, sc = env.step(acts[t:t + 1], want_obs=False)
            assert abs(float(r[0]) - float(rews[t])) < 1e-6, (name, t)
            if scores[t] >= 0:
                assert sc[0] == scores[t], (name, t)
            if term[0]:
                ends.append(t)
        np.testing.assert_array_equal(np.array(ends), g[name + "_episode_end"])


def test_epsilon_schedule_known_answers(golden_dir):
    """BrainDQN.py:112-114 in float64: 0.03 - k*3e-8 leaves -6.66460567250383e-13 (dqn.log:4)."""
    eps = 0.03
    k = 0
    while eps > 0:
        eps -= (0.03 - 0) / 1000000.
        k += 1
    assert eps == -6.66460567250383e-13
    g = np.load(os.path.join(golden_dir, "ref_logs.npz"))
    assert g["dqn_epsilon"][0] == eps
    # prioritydqn.log: epsilon first decrements at onlineTimeStep 1001 (OBSERVE = 1000)
    pe = g["prioritydqn_epsilon"]
    ts = g["prioritydqn_timestep"]
    e, expect = 0.03, []
    for t in ts:
        expect.append(e)           # setPerception prints epsilon after getAction already decremented it
    first_change = int(np.argmax(pe != 0.03))
    assert ts[first_change] == 1001
    e = 0.03
    for i in range(first_change, len(pe)):
        e -= 0.03 / 1000000.
        assert pe[i] == e, i
