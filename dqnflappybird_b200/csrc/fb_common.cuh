// Shared declarations for libflappy_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/flappy_b200.h"

void fb_set_error(const std::string &msg);

#define FB_CUDA_OK(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            fb_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
            return FB_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define FB_REQUIRE(cond, msg)                                                                 \
    do {                                                                                      \
        if (!(cond)) { fb_set_error(std::string(msg)); return FB_ERR_INVALID; }               \
    } while (0)

// ---- game constants (game/wrapped_flappy_bird.py:14-52, flappy_bird_utils.py) ------------
constexpr int kScreenW = 288, kScreenH = 512;
constexpr int kBirdW = 34, kBirdH = 24;      // PLAYER_WIDTH / PLAYER_HEIGHT (:46-47)
constexpr int kPipeW = 52, kPipeH = 320;     // PIPE_WIDTH / PIPE_HEIGHT (:48-49)
constexpr int kBaseW = 336, kBaseH = 112;
constexpr int kPlayerX = 57;                 // int(SCREENWIDTH * 0.2) (:61)
constexpr int kInitY = 244;                  // int((SCREENHEIGHT - PLAYER_HEIGHT) / 2) (:62)
constexpr int kBaseYDraw = 404;              // int(BASEY) as a blit coordinate (:171)
constexpr int kGroundY = 380;                // playery + velY >= 380  <=>  ground crash (:115, :251)
constexpr int kBaseShift = kBaseW - kScreenW;  // 48 (:64)
constexpr int kGapSize = 100;                // PIPEGAPSIZE (:43)
constexpr int kMaxY = 379;                   // largest playery that is ever drawn

constexpr int kObs = FB_OBS;
constexpr int kBaseJ = 63;                   // obs columns j >= 63 sample only the base strip
constexpr int kBirdRows = 9;                 // obs rows whose taps can touch the bird (i = 16..24)

// ---- packed per-env state: 32 bytes ------------------------------------------------------
struct __align__(16) EnvState {
    int16_t y;          // playery 0..379
    int8_t vel;         // playerVelY -9..10
    uint8_t pidx;       // playerIndex 0..2
    uint8_t loop;       // loopIter 0..29
    uint8_t phase;      // position in PLAYER_INDEX_GEN cycle([0,1,2,1]); survives resets (:52)
    int8_t basex;       // -47..0
    uint8_t npipes;     // 2..3
    int16_t px[3];      // pipe x (upper and lower share it)
    uint8_t gap[3];     // gap index 0..7: gapY = 100 + 10*gap, upper y = gapY-320, lower y = gapY+100
    uint8_t pad0;
    int16_t pad1;
    int32_t score;
    uint32_t draws;     // gap-RNG words (or script entries) consumed
    uint32_t pad2;
};
static_assert(sizeof(EnvState) == 32, "EnvState must be 32 bytes");

// ---- tables derived from the sprites (built on the host in fb_assets.cu) -------------------
struct ObsTables {                           // copied to shared memory by the step kernel
    unsigned long long pipeObs[54][5][8];    // [sx - pipe.x + 1][i % 5][gap] -> bits j = 0..62 of a pipe-only row
    unsigned char birdObs[3][kMaxY + 1][kBirdRows + 3];  // [pidx][y][i - 16] -> 8 bits j = birdJ0[y] + k (12 B stride)
    unsigned char birdJ0[kMaxY + 1 + 4];     // first obs column the bird at y can touch
    short sx[kObs];                          // first source column (game x) of obs row i
};

static_assert(sizeof(ObsTables) % 16 == 0, "ObsTables is copied with 16-byte loads");

struct HitTables {                           // HITMASKS of flappy_bird_utils.getHitmask as bit rows: 5,696 bytes, copied to shared memory
    unsigned long long birdRow[3][kBirdH];   // hitmask bit rows: bit x of row r  (getHitmask)
    unsigned long long pipeRowLo[kPipeH];    // lower-pipe hitmask bit rows
    unsigned long long pipeRowUp[kPipeH];    // upper (rotated 180) pipe
};
static_assert(sizeof(HitTables) % 16 == 0, "HitTables is copied with 16-byte loads");

struct ExactTables {                         // global memory; rare exact path + full render; begins with the HitTables members
    unsigned long long birdRow[3][kBirdH];   // hitmask bit rows: bit x of row r  (getHitmask)
    unsigned long long pipeRowLo[kPipeH];    // lower-pipe hitmask bit rows
    unsigned long long pipeRowUp[kPipeH];    // upper (rotated 180) pipe
    uint32_t birdPix[3][kBirdW][kBirdH];     // RGBA little endian: r | g<<8 | b<<16 | a<<24
    uint32_t pipeLo[kPipeW][kPipeH];
    uint32_t pipeUp[kPipeW][kPipeH];
    uint32_t basePix[kBaseW][kBaseH];
    int sx[kObs], a0[kObs], a1[kObs];        // cv2 coefficient tables, rows of array3d = game x
    int sy[kObs], b0[kObs], b1[kObs];        // cols of array3d = game y
};

struct FbTables {
    ObsTables *obs_dev;
    ExactTables *exact_dev;
    int device;
    bool loaded;
};
const FbTables &fb_tables();
const ObsTables *fb_host_obs_tables();      // host copies (debug / CPU tests)
const ExactTables *fb_host_exact_tables();

// ---- Philox4x32-10 (Salmon et al. SC'11) -------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// word n of stream (seed, purpose, env): lane n&3 of block n>>2, counter (block, purpose, env_lo, env_hi)
__host__ __device__ inline uint32_t stream_word(uint64_t seed, uint32_t purpose, uint64_t env, uint32_t n) {
    uint32_t o[4];
    philox4x32_10(n >> 2, purpose, (uint32_t)env, (uint32_t)(env >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return o[n & 3];
}

// ---- cv2 fixed-point bilinear + gray + threshold for one output pixel -----------------------
// p00 = P[sx][sy], p01 = P[sx][sy+1], p10 = P[sx+1][sy], p11 = P[sx+1][sy+1], packed r|g<<8|b<<16.
// FlappyBirdDQN.py:31-34; cv2 HResizeLinear / VResizeLinear (8U, 11-bit coefficients) and
// RGB2Gray 15-bit coefficients applied to RGB data as if it were BGR.
__host__ __device__ inline int obs_pixel_gt1(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, int a0, int a1,
                                             int b0, int b1) {
    int v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int s = 8 * c;
        int h0 = (int)((p00 >> s) & 255) * b0 + (int)((p01 >> s) & 255) * b1;
        int h1 = (int)((p10 >> s) & 255) * b0 + (int)((p11 >> s) & 255) * b1;
        v[c] = (((a0 * (h0 >> 4)) >> 16) + ((a1 * (h1 >> 4)) >> 16) + 2) >> 2;
    }
    int g = (3735 * v[0] + 19235 * v[1] + 9798 * v[2] + (1 << 14)) >> 15;
    return g > 1;
}
