// Host-side test hooks: run the SAME physics / table / exact-pixel code the kernels run
// (fb_env_logic.cuh is __host__ __device__) on the CPU, so the CPU-only test-suite can check
// the derived tables and the step logic against the oracle without a GPU.  Not used by the
// product path.
#include <stddef.h>
#include <string.h>

#include "fb_env_logic.cuh"

extern "C" int fb_debug_host_step(int32_t *state16, int action, const uint8_t *gaps, int gaps_len, uint64_t seed,
                                  uint64_t env_id, float *reward, uint8_t *terminal, int32_t *score) {
    const ExactTables *ex = fb_host_exact_tables();
    if (!ex) { fb_set_error("fb_debug_host_step: load the assets first"); return FB_ERR_ASSETS; }
    FB_REQUIRE(state16 && reward && terminal && score, "fb_debug_host_step: NULL argument");
    if (action != 0 && action != 1) { fb_set_error("Multiple input actions!"); return FB_ERR_ACTION; }
    EnvState s;
    memset(&s, 0, sizeof(s));
    if (!ints_to_state(state16, s)) { fb_set_error("fb_debug_host_step: state outside the reachable range"); return FB_ERR_STATE; }
    GapSource g{gaps, gaps_len, seed, env_id};
    env_step(s, action, g, ex, *reward, *terminal, *score);
    state_to_ints(s, state16);
    return FB_OK;
}

extern "C" int fb_debug_host_reset(int32_t *state16, const uint8_t *gaps, int gaps_len, uint64_t seed, uint64_t env_id) {
    FB_REQUIRE(state16 != nullptr, "fb_debug_host_reset: NULL argument");
    EnvState s;
    memset(&s, 0, sizeof(s));
    GapSource g{gaps, gaps_len, seed, env_id};
    env_reset(s, g);
    state_to_ints(s, state16);
    return FB_OK;
}

// mode 0: table path (obs_row_mask + per-pixel fix of the bird window when flagged), as render_env does;
// mode 1: every pixel by exact arithmetic, base strip included, as env_obs_exact_kernel does.
extern "C" int fb_debug_host_obs(const int32_t *state16, int mode, uint8_t *out) {
    const ObsTables *T = fb_host_obs_tables();
    const ExactTables *ex = fb_host_exact_tables();
    if (!T || !ex) { fb_set_error("fb_debug_host_obs: load the assets first"); return FB_ERR_ASSETS; }
    EnvState s;
    memset(&s, 0, sizeof(s));
    if (!ints_to_state(state16, s)) { fb_set_error("fb_debug_host_obs: state outside the reachable range"); return FB_ERR_STATE; }
    DrawList d = make_draw_list(s);
    if (mode == 1) {
        for (int i = 0; i < kObs; i++)
            for (int j = 0; j < kObs; j++) out[i * kObs + j] = exact_obs_bit(ex, d, s.basex, true, i, j) ? 255 : 0;
        return FB_OK;
    }
    const int j0 = T->birdJ0[d.y];
    for (int i = 0; i < kObs; i++) {
        unsigned long long m = obs_row_mask(*T, d, i);
        int r = i - 16;
        if ((d.np_mixed & 16) && r >= 0 && r < kBirdRows) {
            unsigned byte = 0;
            for (int b = 0; b < 8; b++)
                if (j0 + b < kBaseJ && exact_obs_bit(ex, d, 0, false, i, j0 + b)) byte |= 1u << b;
            m = obs_row_fix(m, byte, j0);
        }
        for (int j = 0; j < kObs; j++) out[i * kObs + j] = (j >= 64 || ((m >> j) & 1)) ? 255 : 0;
    }
    return FB_OK;
}

extern "C" int fb_debug_host_mixed(const int32_t *state16) {
    EnvState s;
    memset(&s, 0, sizeof(s));
    if (!ints_to_state(state16, s)) return -1;
    return (make_draw_list(s).np_mixed & 16) ? 1 : 0;
}

// layout of fb_step_sampling as this library was compiled: sizeof, then the byte offset of every field in declaration
// order -- the CPU test-suite checks the ctypes mirror (dqnflappybird_b200/_lib.py StepSampling) against it
extern "C" int fb_debug_step_sampling_layout(int32_t *out, int capacity) {
    const int32_t v[] = {(int32_t)sizeof(fb_step_sampling),
                         (int32_t)offsetof(fb_step_sampling, replay), (int32_t)offsetof(fb_step_sampling, ring_dev),
                         (int32_t)offsetof(fb_step_sampling, act_dev), (int32_t)offsetof(fb_step_sampling, rew_dev),
                         (int32_t)offsetof(fb_step_sampling, term_dev), (int32_t)offsetof(fb_step_sampling, t),
                         (int32_t)offsetof(fb_step_sampling, batch), (int32_t)offsetof(fb_step_sampling, setsize),
                         (int32_t)offsetof(fb_step_sampling, seed), (int32_t)offsetof(fb_step_sampling, idx_out_dev),
                         (int32_t)offsetof(fb_step_sampling, frames_out_dev), (int32_t)offsetof(fb_step_sampling, act_out_dev),
                         (int32_t)offsetof(fb_step_sampling, rew_out_dev), (int32_t)offsetof(fb_step_sampling, term_out_dev),
                         (int32_t)offsetof(fb_step_sampling, env_out_dev), (int32_t)offsetof(fb_step_sampling, k_out_dev),
                         (int32_t)offsetof(fb_step_sampling, prioritized), (int32_t)offsetof(fb_step_sampling, per_mode),
                         (int32_t)offsetof(fb_step_sampling, beta), (int32_t)offsetof(fb_step_sampling, tree_idx_out_dev),
                         (int32_t)offsetof(fb_step_sampling, is_weights_out_dev), (int32_t)offsetof(fb_step_sampling, prio_out_dev),
                         (int32_t)offsetof(fb_step_sampling, is_weights_f32_out_dev)};
    const int n = (int)(sizeof(v) / sizeof(v[0]));
    FB_REQUIRE(out != nullptr && capacity >= n, "fb_debug_step_sampling_layout: buffer too small");
    for (int i = 0; i < n; i++) out[i] = v[i];
    return n;
}
