// Batched Flappy Bird environment: frame_step + render + preprocess in one kernel.
//
// Replaces, for N independent envs at once,
//   game/wrapped_flappy_bird.py:87-162   frame_step physics, spawn, score, checkCrash, reset
//   game/wrapped_flappy_bird.py:165-177  the pygame blits and surfarray.array3d
//   FlappyBirdDQN.py:31-34               cv2.resize(80x80) / cvtColor / threshold
//
// Kernel shape.  A CTA owns EPC envs for the whole launch (state stays in registers across the
// n_steps of one launch).  Per step: (1) one thread per env runs the integer physics and posts a
// 16-byte draw list to shared memory; (2) one warp per env builds the 80 observation rows as
// 64-bit masks (pipe rows and bird windows come from tables derived from the sprites with cv2's
// exact fixed-point arithmetic; the rare rows where bird and pipe pixels share a 2x2 tap
// footprint are evaluated per pixel) and streams the 6400-byte frame into the ring with 16-byte
// stores, 512 contiguous bytes per warp instruction.  No full-resolution frame exists anywhere.
#include <new>

#include "fb_env_logic.cuh"

struct fb_env {
    int n;
    uint64_t seed, first_id;
    EnvState *state;
    const uint8_t *gaps;
    int gaps_len;
    int *err_flag;
    // device staging for the host-buffer calls, two slots so that the copies of one step overlap the kernel of the next
    uint8_t *stage_act[2];
    float *stage_rew[2];
    uint8_t *stage_term[2];
    int32_t *stage_score[2];
    cudaStream_t s_in, s_out;               // H2D / D2H copy streams
    cudaEvent_t ev_in[2], ev_step[2], ev_out[2];
    unsigned long long submitted, waited;   // host-step tickets
    int num_sms;
};

struct StepArgs {
    EnvState *state;
    int n;
    uint64_t seed, first_id;
    const uint8_t *gaps;
    int gaps_len;
    const uint8_t *actions;         // u8[n_steps][n] or nullptr (device-generated)
    uint64_t act_seed;
    uint32_t first_step, flap_threshold;
    uint8_t *actions_out;
    uint8_t *ring;
    int ring_len, ring_slot;
    float *reward;
    uint8_t *terminal;
    int32_t *score;
    int n_steps;
    int draw_only;                  // 1: draw the current state once, no physics
    int *err_flag;
    const ObsTables *obs_tab;
    const ExactTables *ex;
};

__device__ __forceinline__ uint32_t expand4(uint32_t nib) {      // 4 bits -> 4 bytes of 0x00 / 0xFF
    return ((nib * 0x00204081u) & 0x01010101u) * 255u;
}

// One warp draws one env's 80x80 observation.
__device__ __forceinline__ void render_env(const ObsTables &T, const ExactTables *ex, const DrawList d,
                                           unsigned long long *rowmask, uint8_t *out, int lane) {
    const int j0 = T.birdJ0[d.y];
#pragma unroll
    for (int it = 0; it < 3; it++) {
        int i = lane + 32 * it;
        if (i < kObs) rowmask[i] = obs_row_mask(T, d, i);
    }
    __syncwarp();
    if (d.np_mixed & 16) {                           // rare: evaluate the bird window per pixel
#pragma unroll 1
        for (int q = 0; q < 3; q++) {
            int idx = q * 32 + lane, r = idx >> 3, j = j0 + (idx & 7);
            int bit = 0;
            if (r < kBirdRows && j < kBaseJ) bit = exact_obs_bit(ex, d, 0, false, 16 + r, j);
            unsigned bal = __ballot_sync(0xFFFFFFFFu, bit);
            int rr = 4 * q + lane;
            if (lane < 4 && rr < kBirdRows) {
                rowmask[16 + rr] = obs_row_fix(rowmask[16 + rr], (bal >> (8 * lane)) & 0xFFu, j0);
            }
        }
        __syncwarp();
    }
    uint4 *dst = reinterpret_cast<uint4 *>(out);
#pragma unroll
    for (int it = 0; it < 13; it++) {
        int c = lane + 32 * it;                      // 16-byte chunk: row c / 5, columns 16 * (c % 5) ..
        if (c < 400) {
            int i = (c * 205) >> 10, part = c - 5 * i;
            uint32_t bits = part == 4 ? 0xFFFFu : (uint32_t)(rowmask[i] >> (16 * part)) & 0xFFFFu;
            uint4 v;
            v.x = expand4(bits & 15u); v.y = expand4((bits >> 4) & 15u);
            v.z = expand4((bits >> 8) & 15u); v.w = expand4(bits >> 12);
            __stcs(dst + c, v);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------- the step kernel

template <int THREADS, int EPC>
__global__ void __launch_bounds__(THREADS) env_step_kernel(const StepArgs a) {
    constexpr int WARPS = THREADS / 32;
    __shared__ __align__(16) ObsTables T;
    __shared__ DrawList dl[EPC];
    __shared__ unsigned long long rowmask[WARPS][kObs];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (a.ring) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.obs_tab);
        uint4 *dstT = reinterpret_cast<uint4 *>(&T);
        for (int k = tid; k < (int)(sizeof(ObsTables) / 16); k += THREADS) dstT[k] = __ldg(src + k);
    }
    __syncthreads();

    const int n_groups = (a.n + EPC - 1) / EPC;
    for (int group = blockIdx.x; group < n_groups; group += gridDim.x) {
        const int env0 = group * EPC;
        const int e_mine = env0 + tid;
        const bool have = tid < EPC && e_mine < a.n;
        EnvState s;
        GapSource gs;
        if (have) {
            const uint4 *sp = reinterpret_cast<const uint4 *>(a.state + e_mine);
            uint4 lo = sp[0], hi = sp[1];
            *reinterpret_cast<uint4 *>(&s) = lo;
            *(reinterpret_cast<uint4 *>(&s) + 1) = hi;
            gs.script = a.gaps ? a.gaps + (size_t)e_mine * a.gaps_len : nullptr;
            gs.script_len = a.gaps_len;
            gs.seed = a.seed; gs.env_id = a.first_id + (uint64_t)e_mine;
        }
        const int n_here = min(EPC, a.n - env0);
        for (int step = 0; step < a.n_steps; step++) {
            if (have && a.draw_only) dl[tid] = make_draw_list(s);
            if (have && !a.draw_only) {
                int act;
                const size_t o = (size_t)step * a.n + e_mine;
                if (a.actions) {
                    act = a.actions[o];
                    if (act > 1) { atomicOr(a.err_flag, 1); act = 0; }
                } else {
                    act = stream_word(a.act_seed, 1u, gs.env_id, a.first_step + (uint32_t)step) < a.flap_threshold ? 1 : 0;
                    if (a.actions_out) a.actions_out[o] = (uint8_t)act;
                }
                float rew; uint8_t term; int32_t sc;
                env_step(s, act, gs, a.ex, rew, term, sc);
                if (a.reward) a.reward[o] = rew;
                if (a.terminal) a.terminal[o] = term;
                if (a.score) a.score[o] = sc;
                if (a.ring) dl[tid] = make_draw_list(s);
            }
            if (a.ring) {
                const int slot = (a.ring_slot + step) % a.ring_len;
                if (EPC == THREADS) {
                    // each warp draws the 32 envs its own lanes have just stepped: only warp-level synchronisation, so the
                    // warps of a CTA drift apart and one warp's physics overlaps another's stores
                    __syncwarp();
                    for (int e = warp * 32; e < min(n_here, warp * 32 + 32); e++) {
                        uint8_t *out = a.ring + ((size_t)(env0 + e) * a.ring_len + slot) * (size_t)FB_FRAME_BYTES;
                        render_env(T, a.ex, dl[e], rowmask[warp], out, lane);
                    }
                    __syncwarp();
                } else {
                    __syncthreads();
                    for (int e = warp; e < n_here; e += WARPS) {
                        uint8_t *out = a.ring + ((size_t)(env0 + e) * a.ring_len + slot) * (size_t)FB_FRAME_BYTES;
                        render_env(T, a.ex, dl[e], rowmask[warp], out, lane);
                    }
                    __syncthreads();
                }
            }
        }
        if (have) {
            uint4 *sp = reinterpret_cast<uint4 *>(a.state + e_mine);
            sp[0] = *reinterpret_cast<uint4 *>(&s);
            sp[1] = *(reinterpret_cast<uint4 *>(&s) + 1);
        }
    }
}

// ------------------------------------------------------------------------------- small kernels

__global__ void env_reset_kernel(EnvState *state, int n, uint64_t seed, uint64_t first_id, const uint8_t *gaps, int gaps_len) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    EnvState s;
    memset(&s, 0, sizeof(s));
    GapSource gs{gaps ? gaps + (size_t)e * gaps_len : nullptr, gaps_len, seed, first_id + (uint64_t)e};
    env_reset(s, gs);
    state[e] = s;
}

__global__ void env_export_kernel(const EnvState *state, int n, int32_t *out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    EnvState s = state[e];
    state_to_ints(s, out + (size_t)e * FB_STATE_INTS);
}

__global__ void env_import_kernel(EnvState *state, int n, const int32_t *in, int *err_flag) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    EnvState s;
    if (!ints_to_state(in + (size_t)e * FB_STATE_INTS, s)) { atomicOr(err_flag, 2); return; }
    state[e] = s;
}

// per-pixel evaluation of the whole observation, base included: the in-library cross-check
__global__ void env_obs_exact_kernel(const EnvState *state, int n, const ExactTables *ex, uint8_t *obs) {
    int e = blockIdx.x;
    if (e >= n) return;
    __shared__ DrawList d;
    __shared__ int basex;
    if (threadIdx.x == 0) { EnvState s = state[e]; d = make_draw_list(s); basex = s.basex; }
    __syncthreads();
    for (int p = threadIdx.x; p < kObs * kObs; p += blockDim.x) {
        int i = p / kObs, j = p - i * kObs;
        obs[(size_t)e * FB_FRAME_BYTES + p] = exact_obs_bit(ex, d, basex, true, i, j) ? 255 : 0;
    }
}

// image_data = surfarray.array3d(display) (wrapped_flappy_bird.py:177): u8[x][y][3]
__global__ void env_render_full_kernel(const EnvState *state, int first, int n, const ExactTables *ex, uint8_t *rgb) {
    int e = blockIdx.y;
    if (e >= n) return;
    __shared__ DrawList d;
    __shared__ int basex;
    if (threadIdx.x == 0) { EnvState s = state[first + e]; d = make_draw_list(s); basex = s.basex; }
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kScreenW * kScreenH) return;
    int X = p / kScreenH, Y = p - X * kScreenH;
    uint32_t c = scene_rgb(ex, d, basex, true, X, Y);
    uint8_t *o = rgb + ((size_t)e * kScreenW * kScreenH + p) * 3;
    o[0] = c & 255; o[1] = (c >> 8) & 255; o[2] = (c >> 16) & 255;
}

// ------------------------------------------------------------------------------- C ABI

extern "C" int fb_env_create(int n_envs, uint64_t seed, uint64_t first_env_id, fb_env **out) {
    FB_REQUIRE(out != nullptr, "fb_env_create: out is NULL");
    FB_REQUIRE(n_envs > 0, "fb_env_create: n_envs must be positive");
    if (!fb_tables().loaded) { fb_set_error("fb_env_create: call fb_assets_load first"); return FB_ERR_ASSETS; }
    fb_env *e = new (std::nothrow) fb_env();
    FB_REQUIRE(e != nullptr, "fb_env_create: out of host memory");
    e->n = n_envs; e->seed = seed; e->first_id = first_env_id; e->gaps = nullptr; e->gaps_len = 0;
    int dev = 0;
    FB_CUDA_OK(cudaGetDevice(&dev));
    FB_CUDA_OK(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, dev));
    FB_CUDA_OK(cudaMalloc(&e->state, sizeof(EnvState) * (size_t)n_envs));
    FB_CUDA_OK(cudaMalloc(&e->err_flag, sizeof(int)));
    FB_CUDA_OK(cudaMemset(e->err_flag, 0, sizeof(int)));
    for (int k = 0; k < 2; k++) {
        FB_CUDA_OK(cudaMalloc(&e->stage_act[k], (size_t)n_envs));
        FB_CUDA_OK(cudaMalloc(&e->stage_rew[k], sizeof(float) * (size_t)n_envs));
        FB_CUDA_OK(cudaMalloc(&e->stage_term[k], (size_t)n_envs));
        FB_CUDA_OK(cudaMalloc(&e->stage_score[k], sizeof(int32_t) * (size_t)n_envs));
        FB_CUDA_OK(cudaEventCreateWithFlags(&e->ev_in[k], cudaEventDisableTiming));
        FB_CUDA_OK(cudaEventCreateWithFlags(&e->ev_step[k], cudaEventDisableTiming));
        FB_CUDA_OK(cudaEventCreateWithFlags(&e->ev_out[k], cudaEventDisableTiming));
    }
    FB_CUDA_OK(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    FB_CUDA_OK(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
    e->submitted = e->waited = 0;
    *out = e;
    return fb_env_reset(e, nullptr);
}

extern "C" int fb_env_destroy(fb_env *e) {
    if (!e) return FB_OK;
    cudaFree(e->state); cudaFree(e->err_flag);
    for (int k = 0; k < 2; k++) {
        cudaFree(e->stage_act[k]); cudaFree(e->stage_rew[k]); cudaFree(e->stage_term[k]); cudaFree(e->stage_score[k]);
        if (e->ev_in[k]) cudaEventDestroy(e->ev_in[k]);
        if (e->ev_step[k]) cudaEventDestroy(e->ev_step[k]);
        if (e->ev_out[k]) cudaEventDestroy(e->ev_out[k]);
    }
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_out) cudaStreamDestroy(e->s_out);
    delete e;
    return FB_OK;
}

extern "C" int fb_env_num_envs(const fb_env *e) { return e ? e->n : 0; }

extern "C" int fb_env_set_gap_replay(fb_env *e, const uint8_t *gaps_dev, int per_env_len) {
    FB_REQUIRE(e != nullptr, "fb_env_set_gap_replay: env is NULL");
    FB_REQUIRE(gaps_dev == nullptr || per_env_len > 0, "fb_env_set_gap_replay: per_env_len must be positive");
    e->gaps = gaps_dev; e->gaps_len = gaps_dev ? per_env_len : 0;
    return FB_OK;
}

extern "C" int fb_env_reset(fb_env *e, void *stream) {
    FB_REQUIRE(e != nullptr, "fb_env_reset: env is NULL");
    env_reset_kernel<<<(e->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->state, e->n, e->seed, e->first_id, e->gaps, e->gaps_len);
    FB_CUDA_OK(cudaGetLastError());
    return FB_OK;
}

static int launch_step(fb_env *e, StepArgs &a, cudaStream_t st) {
    a.state = e->state; a.n = e->n; a.seed = e->seed; a.first_id = e->first_id;
    a.gaps = e->gaps; a.gaps_len = e->gaps_len; a.err_flag = e->err_flag;
    a.obs_tab = fb_tables().obs_dev; a.ex = fb_tables().exact_dev;
    // Few envs: 32 per CTA so that every SM gets work; many envs: 128 per CTA (4 full physics warps).
    if (e->n < 128 * e->num_sms) {
        int groups = (e->n + 31) / 32;
        env_step_kernel<128, 32><<<groups, 128, 0, st>>>(a);
    } else {
        int groups = (e->n + 127) / 128;
        env_step_kernel<128, 128><<<groups, 128, 0, st>>>(a);
    }
    FB_CUDA_OK(cudaGetLastError());
    return FB_OK;
}

static int check_ring(const fb_env *e, const uint8_t *ring, int ring_len, int ring_slot, int n_steps) {
    FB_REQUIRE(e != nullptr, "fb_env_step: env is NULL");
    FB_REQUIRE(n_steps > 0, "fb_env_step: n_steps must be positive");
    if (ring) {
        FB_REQUIRE(ring_len > 0 && ring_slot >= 0 && ring_slot < ring_len, "fb_env_step: ring_slot outside the ring");
        FB_REQUIRE((reinterpret_cast<uintptr_t>(ring) & 15) == 0, "fb_env_step: obs ring must be 16-byte aligned");
    }
    return FB_OK;
}

extern "C" int fb_env_step(fb_env *e, int n_steps, const uint8_t *actions_dev, uint8_t *obs_ring_dev, int ring_len,
                           int ring_slot, float *reward_dev, uint8_t *terminal_dev, int32_t *score_dev, void *stream) {
    int rc = check_ring(e, obs_ring_dev, ring_len, ring_slot, n_steps);
    if (rc) return rc;
    FB_REQUIRE(actions_dev != nullptr, "fb_env_step: actions_dev is NULL");
    StepArgs a{};
    a.actions = actions_dev; a.ring = obs_ring_dev; a.ring_len = ring_len; a.ring_slot = ring_slot;
    a.reward = reward_dev; a.terminal = terminal_dev; a.score = score_dev; a.n_steps = n_steps;
    return launch_step(e, a, (cudaStream_t)stream);
}

extern "C" int fb_env_draw(fb_env *e, uint8_t *obs_ring_dev, int ring_len, int ring_slot, void *stream) {
    int rc = check_ring(e, obs_ring_dev, ring_len, ring_slot, 1);
    if (rc) return rc;
    FB_REQUIRE(obs_ring_dev != nullptr, "fb_env_draw: obs_ring_dev is NULL");
    StepArgs a{};
    a.ring = obs_ring_dev; a.ring_len = ring_len; a.ring_slot = ring_slot; a.n_steps = 1; a.draw_only = 1;
    return launch_step(e, a, (cudaStream_t)stream);
}

extern "C" int fb_env_step_random(fb_env *e, int n_steps, uint64_t action_seed, uint32_t first_step, uint32_t flap_threshold,
                                  uint8_t *actions_out_dev, uint8_t *obs_ring_dev, int ring_len, int ring_slot,
                                  float *reward_dev, uint8_t *terminal_dev, int32_t *score_dev, void *stream) {
    int rc = check_ring(e, obs_ring_dev, ring_len, ring_slot, n_steps);
    if (rc) return rc;
    StepArgs a{};
    a.actions = nullptr; a.act_seed = action_seed; a.first_step = first_step; a.flap_threshold = flap_threshold;
    a.actions_out = actions_out_dev; a.ring = obs_ring_dev; a.ring_len = ring_len; a.ring_slot = ring_slot;
    a.reward = reward_dev; a.terminal = terminal_dev; a.score = score_dev; a.n_steps = n_steps;
    return launch_step(e, a, (cudaStream_t)stream);
}

// frame_step with HOST buffers, split in two so that a caller can keep two steps in flight: the H2D copy of the
// actions, the step kernel and the D2H copies of reward / terminal / score run on three streams chained by events,
// and the copies of step t overlap the kernel of step t+1.  The reference raises ValueError for a non one-hot action
// BEFORE touching the state (wrapped_flappy_bird.py:99-100): the actions are checked on the host first.
extern "C" int fb_env_step_host_submit(fb_env *e, const uint8_t *actions_host, uint8_t *obs_ring_dev, int ring_len, int ring_slot,
                                       float *reward_host, uint8_t *terminal_host, int32_t *score_host, void *stream) {
    int rc = check_ring(e, obs_ring_dev, ring_len, ring_slot, 1);
    if (rc) return rc;
    FB_REQUIRE(actions_host != nullptr, "fb_env_step_host: actions_host is NULL");
    FB_REQUIRE(e->submitted - e->waited < 2, "fb_env_step_host_submit: two steps are already in flight; call fb_env_step_host_wait");
    cudaStream_t st = (cudaStream_t)stream;
    {   // every byte must be 0 or 1: eight at a time
        const int n = e->n;
        int k = 0;
        unsigned long long bad = 0;
        for (; k + 8 <= n; k += 8) { unsigned long long w; memcpy(&w, actions_host + k, 8); bad |= w & 0xFEFEFEFEFEFEFEFEull; }
        for (; k < n; k++) bad |= (unsigned long long)(actions_host[k] & 0xFE);
        if (bad) { fb_set_error("Multiple input actions!"); return FB_ERR_ACTION; }
    }
    const int sl = (int)(e->submitted & 1);
    FB_CUDA_OK(cudaMemcpyAsync(e->stage_act[sl], actions_host, (size_t)e->n, cudaMemcpyHostToDevice, e->s_in));
    FB_CUDA_OK(cudaEventRecord(e->ev_in[sl], e->s_in));
    FB_CUDA_OK(cudaStreamWaitEvent(st, e->ev_in[sl], 0));
    StepArgs a{};
    a.actions = e->stage_act[sl]; a.ring = obs_ring_dev; a.ring_len = ring_len; a.ring_slot = ring_slot;
    a.reward = e->stage_rew[sl]; a.terminal = e->stage_term[sl]; a.score = e->stage_score[sl]; a.n_steps = 1;
    rc = launch_step(e, a, st);
    if (rc) return rc;
    FB_CUDA_OK(cudaEventRecord(e->ev_step[sl], st));
    FB_CUDA_OK(cudaStreamWaitEvent(e->s_out, e->ev_step[sl], 0));
    if (reward_host) FB_CUDA_OK(cudaMemcpyAsync(reward_host, e->stage_rew[sl], sizeof(float) * (size_t)e->n, cudaMemcpyDeviceToHost, e->s_out));
    if (terminal_host) FB_CUDA_OK(cudaMemcpyAsync(terminal_host, e->stage_term[sl], (size_t)e->n, cudaMemcpyDeviceToHost, e->s_out));
    if (score_host) FB_CUDA_OK(cudaMemcpyAsync(score_host, e->stage_score[sl], sizeof(int32_t) * (size_t)e->n, cudaMemcpyDeviceToHost, e->s_out));
    FB_CUDA_OK(cudaEventRecord(e->ev_out[sl], e->s_out));
    e->submitted++;
    return FB_OK;
}

// Blocks until the OLDEST submitted step has delivered its reward / terminal / score to the host buffers.
extern "C" int fb_env_step_host_wait(fb_env *e) {
    FB_REQUIRE(e != nullptr, "fb_env_step_host_wait: env is NULL");
    FB_REQUIRE(e->submitted > e->waited, "fb_env_step_host_wait: nothing in flight");
    FB_CUDA_OK(cudaEventSynchronize(e->ev_out[e->waited & 1]));
    e->waited++;
    return FB_OK;
}

extern "C" int fb_env_step_host(fb_env *e, const uint8_t *actions_host, uint8_t *obs_ring_dev, int ring_len, int ring_slot,
                                float *reward_host, uint8_t *terminal_host, int32_t *score_host, void *stream) {
    FB_REQUIRE(e != nullptr, "fb_env_step_host: env is NULL");
    while (e->submitted > e->waited) { int rc = fb_env_step_host_wait(e); if (rc) return rc; }
    int rc = fb_env_step_host_submit(e, actions_host, obs_ring_dev, ring_len, ring_slot, reward_host, terminal_host, score_host, stream);
    if (rc) return rc;
    return fb_env_step_host_wait(e);
}

extern "C" int fb_env_check(fb_env *e, void *stream) {
    FB_REQUIRE(e != nullptr, "fb_env_check: env is NULL");
    int flag = 0;
    FB_CUDA_OK(cudaMemcpyAsync(&flag, e->err_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    FB_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    if (flag) {
        FB_CUDA_OK(cudaMemsetAsync(e->err_flag, 0, sizeof(int), (cudaStream_t)stream));
        if (flag & 1) { fb_set_error("Multiple input actions!"); return FB_ERR_ACTION; }
        fb_set_error("fb_env_import_state: state outside the reachable range");
        return FB_ERR_STATE;
    }
    return FB_OK;
}

extern "C" int fb_env_export_state(fb_env *e, int32_t *out_dev, void *stream) {
    FB_REQUIRE(e != nullptr && out_dev != nullptr, "fb_env_export_state: NULL argument");
    env_export_kernel<<<(e->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->state, e->n, out_dev);
    FB_CUDA_OK(cudaGetLastError());
    return FB_OK;
}

extern "C" int fb_env_import_state(fb_env *e, const int32_t *in_dev, void *stream) {
    FB_REQUIRE(e != nullptr && in_dev != nullptr, "fb_env_import_state: NULL argument");
    env_import_kernel<<<(e->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->state, e->n, in_dev, e->err_flag);
    FB_CUDA_OK(cudaGetLastError());
    return FB_OK;
}

extern "C" int fb_env_obs_exact(fb_env *e, uint8_t *obs_dev, void *stream) {
    FB_REQUIRE(e != nullptr && obs_dev != nullptr, "fb_env_obs_exact: NULL argument");
    env_obs_exact_kernel<<<e->n, 256, 0, (cudaStream_t)stream>>>(e->state, e->n, fb_tables().exact_dev, obs_dev);
    FB_CUDA_OK(cudaGetLastError());
    return FB_OK;
}

extern "C" int fb_render_full(fb_env *e, int first, int n, uint8_t *rgb_dev, void *stream) {
    FB_REQUIRE(e != nullptr && rgb_dev != nullptr, "fb_render_full: NULL argument");
    FB_REQUIRE(first >= 0 && n > 0 && first + n <= e->n, "fb_render_full: env range outside the handle");
    dim3 grid((kScreenW * kScreenH + 255) / 256, n);
    env_render_full_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(e->state, first, n, fb_tables().exact_dev, rgb_dev);
    FB_CUDA_OK(cudaGetLastError());
    return FB_OK;
}
