// Batched Flappy Bird environment: frame_step + render + preprocess in one kernel.
//
// Replaces, for N independent envs at once,
//   game/wrapped_flappy_bird.py:87-162   frame_step physics, spawn, score, checkCrash, reset
//   game/wrapped_flappy_bird.py:165-177  the pygame blits and surfarray.array3d
//   FlappyBirdDQN.py:31-34               cv2.resize(80x80) / cvtColor / threshold
//
// Kernel shape.  Persistent and warp-centric (see env_step_kernel): every resident warp owns an equal, contiguous share of
// the envs; 32 envs at a time, each lane runs one env's integer physics (collision against the hitmask bit rows held in shared
// memory) and posts a 16-byte draw list; the warp then draws those envs one after another: 80 observation rows as 64-bit masks
// (pipe rows and bird windows come from tables derived from the sprites with cv2's exact fixed-point arithmetic; the rare rows
// where bird and pipe pixels share a 2x2 tap footprint are evaluated per pixel), expanded to bytes through a 256-entry table
// and streamed into the ring with 16-byte stores, 512 contiguous bytes per warp instruction.  No full-resolution frame exists.
#include <stdlib.h>

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
    int step_slots;                         // resident CTAs of env_step_kernel on this device (0 until the first launch)
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

// One warp draws one env's 80x80 observation.  Measured on the B200 (tools/write_bw_probe.py, fb_debug_write_probe): a kernel
// that only WRITES 131,072 frames of 6,400 bytes reaches 5.75 TB/s with 16 resident warps per SM and 6.3 TB/s with 32 --
// whether the bytes leave as 16-byte streaming stores or as cp.async.bulk copies from shared memory, whether the frames are
// contiguous or 25,600 bytes apart.  What a frame-writing kernel needs is stores IN FLIGHT (~200 KB per SM); staging frames in
// shared memory caps the resident warps at 16 beside the 37 KB of tables (measured: 0.79 of the copy peak, below round 1's
// register-store kernel).  So: stores straight from registers, 32 warps per SM (<= 64 registers), and fewer instructions per
// frame than round 1 (~730 issued per frame at 20 warps per SM):
//   * lane l computes the masks of rows l, l + 32, l + 64; what depends only on the row (first source column, phase in cv2's
//     5-row coefficient cycle) is computed once per kernel (RowConst);
//   * bits -> bytes through a 256-entry table (8 bits -> 8 bytes, 2 KB of shared memory) instead of multiply-and-mask;
//   * 16-byte chunk c = lane + 32 it of the frame (row c / 5, columns 16 (c % 5) ..): one warp instruction writes 512 contiguous bytes.
struct RowConst { int sx[3], tab[3]; };      // per lane: first source column and byte offset of the phase block inside pipeObs, rows l, l+32, l+64

// what the drawing warp needs, as whole words (no byte extraction in the inner loops): 32 bytes, read as two broadcast LDS.128
struct __align__(16) DrawWords {
    int px[3];          // pipe x; 9999 for an absent pipe (fails the window test)
    int flags;          // bit 0: bird and pipe pixels can share a 2x2 tap footprint -> per-pixel detour; bits 1-7: first obs column of the
                        // bird window (birdJ0[y]); bits 8-19: y; bits 20-21: pidx; bits 24-27: number of pipes
    int goff[3];        // gap index * 8: byte offset of the gap's mask inside a pipeObs phase block
    int bird;           // (pidx * 380 + y) * 12: byte offset of the bird window rows in birdObs
};
__device__ __forceinline__ DrawWords to_words(const ObsTables &T, const DrawList &d) {
    DrawWords w;
    const int np = d.np_mixed & 15;
#pragma unroll
    for (int k = 0; k < 3; k++) { w.px[k] = k < np ? (int)d.px[k] : 9999; w.goff[k] = (int)d.gap[k] * 8; }
    w.flags = ((d.np_mixed & 16) ? 1 : 0) | ((int)T.birdJ0[d.y] << 1) | ((int)d.y << 8) | ((int)d.pidx << 20) | (np << 24);
    w.bird = ((int)d.pidx * (kMaxY + 1) + (int)d.y) * (kBirdRows + 3);
    return w;
}

// rows of the frame as 16-bit pieces: chunk c = 5 i + part of the frame (16 bytes: row i, columns 16 part ..) has its bits at m16[c];
// part 4 (columns 64..79, the base strip) is 0xFFFF.  The store loop then needs no index arithmetic at all: lane l reads m16[l + 32 it].
constexpr int kChunks = kObs * 5;

__device__ __forceinline__ void render_env(const ObsTables &T, const uint2 *lut8, const ExactTables *ex, const DrawWords w,
                                           const RowConst &rc, unsigned short *m16, unsigned long long *fixrows, uint8_t *out, int lane) {
    const int j0 = (w.flags >> 1) & 127;
    const unsigned char *pipe_bytes = reinterpret_cast<const unsigned char *>(&T.pipeObs[0][0][0]);
    const unsigned char *bird_bytes = reinterpret_cast<const unsigned char *>(&T.birdObs[0][0][0]) + w.bird;
    unsigned long long m[3];
#pragma unroll
    for (int it = 0; it < 3; it++) {
        unsigned long long v = 1ull << 63;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const unsigned c = (unsigned)(rc.sx[it] - w.px[k] + 1);
            if (c < 54u) v |= *reinterpret_cast<const unsigned long long *>(pipe_bytes + c * 320u + rc.tab[it] + w.goff[k]);
        }
        m[it] = v;
    }
    if (lane >= 16 && lane < 16 + kBirdRows) m[0] |= (unsigned long long)bird_bytes[lane - 16] << j0;     // rows 16..24 are lanes 16..24, first pass
    if (w.flags & 1) {                               // rare (warp-uniform): evaluate the bird window per pixel
        DrawList d;                                  // (the byte-packed form the per-pixel scene code reads)
        d.y = (int16_t)((w.flags >> 8) & 4095); d.pidx = (uint8_t)((w.flags >> 20) & 3); d.np_mixed = (uint8_t)(((w.flags >> 24) & 15) | 16);
#pragma unroll
        for (int k = 0; k < 3; k++) { d.px[k] = (int16_t)(w.px[k] == 9999 ? 0 : w.px[k]); d.gap[k] = (uint8_t)(w.goff[k] >> 3); d.pad[k] = 0; }
        if (lane >= 16 && lane < 16 + kBirdRows) fixrows[lane - 16] = m[0];
        __syncwarp();
#pragma unroll 1
        for (int q = 0; q < 3; q++) {
            int idx = q * 32 + lane, r = idx >> 3, j = j0 + (idx & 7);
            int bit = 0;
            if (r < kBirdRows && j < kBaseJ) bit = exact_obs_bit(ex, d, 0, false, 16 + r, j);
            unsigned bal = __ballot_sync(0xFFFFFFFFu, bit);
            int rr = 4 * q + lane;
            if (lane < 4 && rr < kBirdRows) fixrows[rr] = obs_row_fix(fixrows[rr], (bal >> (8 * lane)) & 0xFFu, j0);
        }
        __syncwarp();
        if (lane >= 16 && lane < 16 + kBirdRows) m[0] = fixrows[lane - 16];
    }
#pragma unroll
    for (int it = 0; it < 3; it++) {
        const int i = lane + 32 * it;
        if (i < kObs) {
            unsigned short *row = m16 + 5 * i;
            row[0] = (unsigned short)m[it]; row[1] = (unsigned short)(m[it] >> 16);
            row[2] = (unsigned short)(m[it] >> 32); row[3] = (unsigned short)(m[it] >> 48);
        }
    }
    __syncwarp();
    const unsigned short *mine = m16 + lane;
    uint4 *dst = reinterpret_cast<uint4 *>(out) + lane;
#pragma unroll
    for (int it = 0; it < 13; it++) {
        if (it < 12 || lane < kChunks - 32 * 12) {
            const uint32_t bits = mine[32 * it];
            const uint2 a = lut8[bits & 255u], b = lut8[bits >> 8];
            __stcs(dst + 32 * it, make_uint4(a.x, a.y, b.x, b.y));
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------- the step kernel
// Persistent and warp-centric: the launch has one CTA slot's worth of warps per SM (occupancy query), warp w of W owns the
// contiguous envs [n w / W, n (w+1) / W) -- equal shares, no tail wave, no inter-warp synchronisation after the tables are
// loaded.  A warp walks its range 32 envs at a time: every lane steps one env (integer physics, collision against the
// hitmask bit rows in shared memory) and posts a 16-byte draw list; the warp then draws those envs one after another
// (render_env).  State stays in registers across the n_steps of one launch.
constexpr int kStepThreads = 256, kStepWarps = kStepThreads / 32;

__global__ void __launch_bounds__(kStepThreads, 4) env_step_kernel(const StepArgs a) {
    extern __shared__ __align__(16) unsigned char t_raw[];           // ObsTables (31.5 KB; with the rest 54.7 KB a CTA: four per SM)
    ObsTables &T = *reinterpret_cast<ObsTables *>(t_raw);
    __shared__ __align__(16) HitTables H;                           // flappy_bird_utils hitmasks as bit rows (checkCrash)
    __shared__ DrawWords dw[kStepWarps][32];
    __shared__ __align__(16) unsigned short m16[kStepWarps][kChunks + 8];
    __shared__ unsigned long long fixrows[kStepWarps][kBirdRows + 1];
    __shared__ uint2 lut8[256];                                     // 8 bits -> 8 bytes of 0x00 / 0xFF
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Programmatic dependent launch, both ways: the NEXT step's CTAs may take the slots this step's CTAs free one by one, and
    // load their tables (constants) while this step's last warps still draw; nothing a previous kernel wrote (state, actions)
    // is read before griddepcontrol.wait below.
    asm volatile("griddepcontrol.launch_dependents;");
    lut8[tid] = make_uint2(expand4(tid & 15u), expand4(tid >> 4));

    {
        const uint4 *srcH = reinterpret_cast<const uint4 *>(a.ex);                       // ExactTables begins with the HitTables members
        uint4 *dstH = reinterpret_cast<uint4 *>(&H);
        for (int k = tid; k < (int)(sizeof(HitTables) / 16); k += kStepThreads) dstH[k] = __ldg(srcH + k);
    }
    if (a.ring) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.obs_tab);
        uint4 *dstT = reinterpret_cast<uint4 *>(&T);
        for (int k = tid; k < (int)(sizeof(ObsTables) / 16); k += kStepThreads) dstT[k] = __ldg(src + k);
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    RowConst rc;
#pragma unroll
    for (int it = 0; it < 3; it++) {
        const int i = min(lane + 32 * it, kObs - 1);
        rc.sx[it] = T.sx[i]; rc.tab[it] = (i - 5 * ((i * 205) >> 10)) * 64;
    }
    for (int i = lane; i < kObs; i += 32) m16[warp][5 * i + 4] = 0xFFFFu;        // the base strip: written once
    __syncwarp();
    const long long W = (long long)gridDim.x * kStepWarps, w = (long long)blockIdx.x * kStepWarps + warp;
    const int lo = (int)((long long)a.n * w / W), hi = (int)((long long)a.n * (w + 1) / W);
    for (int base = lo; base < hi; base += 32) {
        const int e_mine = base + lane;
        const bool have = e_mine < hi;
        const int n_here = min(32, hi - base);
        EnvState s;
        GapSource gs;
        if (have) {
            const uint4 *sp = reinterpret_cast<const uint4 *>(a.state + e_mine);
            uint4 v0 = sp[0], v1 = sp[1];
            *reinterpret_cast<uint4 *>(&s) = v0;
            *(reinterpret_cast<uint4 *>(&s) + 1) = v1;
            gs.script = a.gaps ? a.gaps + (size_t)e_mine * a.gaps_len : nullptr;
            gs.script_len = a.gaps_len;
            gs.seed = a.seed; gs.env_id = a.first_id + (uint64_t)e_mine;
        }
        for (int step = 0; step < a.n_steps; step++) {
            if (have && a.draw_only) dw[warp][lane] = to_words(T, make_draw_list(s));
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
                env_step(s, act, gs, &H, rew, term, sc);
                if (a.reward) a.reward[o] = rew;
                if (a.terminal) a.terminal[o] = term;
                if (a.score) a.score[o] = sc;
                if (a.ring) dw[warp][lane] = to_words(T, make_draw_list(s));
            }
            if (a.ring) {
                const int slot = (a.ring_slot + step) % a.ring_len;
                __syncwarp();
                for (int e = 0; e < n_here; e++) {
                    uint8_t *out = a.ring + ((size_t)(base + e) * a.ring_len + slot) * (size_t)FB_FRAME_BYTES;
                    render_env(T, lut8, a.ex, dw[warp][e], rc, m16[warp], fixrows[warp], out, lane);
                }
                __syncwarp();
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
    e->step_slots = 0;
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

static int launch_step(fb_env *e, StepArgs &a, cudaStream_t st, bool early = true) {
    a.state = e->state; a.n = e->n; a.seed = e->seed; a.first_id = e->first_id;
    a.gaps = e->gaps; a.gaps_len = e->gaps_len; a.err_flag = e->err_flag;
    a.obs_tab = fb_tables().obs_dev; a.ex = fb_tables().exact_dev;
    // Few envs: 32 per CTA so that every SM gets work; many envs: 128 per CTA (4 full physics warps).
    if (e->step_slots == 0) {                      // CTAs that fit the device at once (4 per SM: 54.7 KB of shared memory, 64 registers)
        FB_CUDA_OK(cudaFuncSetAttribute(env_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ObsTables)));
        int per_sm = 0;
        FB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, env_step_kernel, kStepThreads, sizeof(ObsTables)));
        FB_REQUIRE(per_sm >= 1, "env_step_kernel does not fit an SM");
        e->step_slots = per_sm * e->num_sms;
    }
    // one env per warp while there are fewer envs than resident warps (every SM draws); equal contiguous shares beyond that
    static int oversub = 0;                        // CTAs per resident slot (FB_ENV_OVERSUB, default 1): > 1 lets the hardware scheduler balance SMs
    if (oversub == 0) { const char *v = getenv("FB_ENV_OVERSUB"); oversub = (v && v[0] >= '1' && v[0] <= '8') ? v[0] - '0' : 1; }
    long long warps = (long long)e->step_slots * kStepWarps * oversub;
    if (warps > e->n) warps = e->n;
    const int grid = (int)((warps + kStepWarps - 1) / kStepWarps);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kStepThreads); cfg.dynamicSmemBytes = sizeof(ObsTables); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    // early (programmatic) launch behind the previous kernel of the stream: 8.30e8 -> 8.55e8 frames/s for back-to-back steps.
    // Not for the host-buffer path, whose kernels wait for an H2D copy through an event (measured 1 % slower there).
    static int pdl = -1;                           // FB_ENV_PDL=0: plain launches everywhere
    if (pdl < 0) { const char *v = getenv("FB_ENV_PDL"); pdl = (v && v[0] == '0') ? 0 : 1; }
    at[0].val.programmaticStreamSerializationAllowed = pdl && early ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    FB_CUDA_OK(cudaLaunchKernelEx(&cfg, env_step_kernel, (const StepArgs)a));
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
// split: physics first (12 us at 131,072 envs: new state, reward / terminal / score), then the drawing as a second launch of the
// same kernel -- the D2H copies (1.2 MB, ~45 us over PCIe) run beside the 150 us of drawing instead of after it.  Used by the
// strictly synchronous fb_env_step_host; with two steps in flight (submit / wait) the copies already overlap the NEXT step's
// kernel and the one fused launch is cheaper.
static int host_submit(fb_env *e, const uint8_t *actions_host, uint8_t *obs_ring_dev, int ring_len, int ring_slot,
                       float *reward_host, uint8_t *terminal_host, int32_t *score_host, void *stream, bool split) {
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
    split = split && obs_ring_dev != nullptr;
    if (split) a.ring = nullptr;                                   // physics only
    rc = launch_step(e, a, st, false);
    if (rc) return rc;
    FB_CUDA_OK(cudaEventRecord(e->ev_step[sl], st));
    if (split) {
        StepArgs d{};
        d.ring = obs_ring_dev; d.ring_len = ring_len; d.ring_slot = ring_slot; d.n_steps = 1; d.draw_only = 1;
        rc = launch_step(e, d, st);
        if (rc) return rc;
        FB_CUDA_OK(cudaEventRecord(e->ev_in[sl], st));             // (ev_in[sl] has done its job above: reused for "frames drawn")
    }
    FB_CUDA_OK(cudaStreamWaitEvent(e->s_out, e->ev_step[sl], 0));
    if (reward_host) FB_CUDA_OK(cudaMemcpyAsync(reward_host, e->stage_rew[sl], sizeof(float) * (size_t)e->n, cudaMemcpyDeviceToHost, e->s_out));
    if (terminal_host) FB_CUDA_OK(cudaMemcpyAsync(terminal_host, e->stage_term[sl], (size_t)e->n, cudaMemcpyDeviceToHost, e->s_out));
    if (score_host) FB_CUDA_OK(cudaMemcpyAsync(score_host, e->stage_score[sl], sizeof(int32_t) * (size_t)e->n, cudaMemcpyDeviceToHost, e->s_out));
    if (split) FB_CUDA_OK(cudaStreamWaitEvent(e->s_out, e->ev_in[sl], 0));   // the step is over when the frames are in the ring, too
    FB_CUDA_OK(cudaEventRecord(e->ev_out[sl], e->s_out));
    e->submitted++;
    return FB_OK;
}

extern "C" int fb_env_step_host_submit(fb_env *e, const uint8_t *actions_host, uint8_t *obs_ring_dev, int ring_len, int ring_slot,
                                       float *reward_host, uint8_t *terminal_host, int32_t *score_host, void *stream) {
    return host_submit(e, actions_host, obs_ring_dev, ring_len, ring_slot, reward_host, terminal_host, score_host, stream, false);
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
    int rc = host_submit(e, actions_host, obs_ring_dev, ring_len, ring_slot, reward_host, terminal_host, score_host, stream, e->n >= 8192);
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
