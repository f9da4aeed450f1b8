// Host+device logic of the env path: physics, collision, draw list, exact per-pixel scene.
// Shared by the step kernel (fb_env.cu) and the host-side debug entry points (fb_debug.cu), so
// the CPU test-suite exercises the very code the kernel runs.
#pragma once
#include "fb_common.cuh"

#define FB_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define FB_LDG(p) __ldg(p)
#else
#define FB_LDG(p) (*(p))
#endif
FB_HD int fb_min(int a, int b) { return a < b ? a : b; }
FB_HD int fb_max(int a, int b) { return a > b ? a : b; }

struct __align__(16) DrawList {     // what the render warp needs, 16 bytes
    int16_t y;
    uint8_t pidx;
    uint8_t np_mixed;               // npipes | mixed << 4
    int16_t px[3];
    uint8_t gap[3];
    uint8_t pad[3];
};
static_assert(sizeof(DrawList) == 16, "DrawList must be 16 bytes");

// ------------------------------------------------------------------------------- physics

struct GapSource {
    const uint8_t *script;          // this env's row of the replay script, or nullptr
    int script_len;
    uint64_t seed, env_id;
};

// random.randint(0, 7) of getRandomPipe (wrapped_flappy_bird.py:212): CPython draws
// getrandbits(4) = top 4 bits of a 32-bit word and rejects values >= 8.
FB_HD int draw_gap(EnvState &s, const GapSource &g) {
    if (g.script) {
        int v = g.script[s.draws % (uint32_t)g.script_len] & 7;
        s.draws++;
        return v;
    }
    for (;;) {
        uint32_t w = stream_word(g.seed, 0u, g.env_id, s.draws);
        s.draws++;
        if ((w >> 28) < 8u) return (int)(w >> 28);
    }
}

// GameState.__init__ (wrapped_flappy_bird.py:59-85); the PLAYER_INDEX_GEN phase is a module
// global there and survives (:52).
FB_HD void env_reset(EnvState &s, const GapSource &g) {
    s.score = 0; s.pidx = 0; s.loop = 0;
    s.y = kInitY; s.vel = 0; s.basex = 0;
    int g1 = draw_gap(s, g), g2 = draw_gap(s, g);
    s.px[0] = kScreenW; s.px[1] = kScreenW + kScreenW / 2; s.px[2] = 0;
    s.gap[0] = (uint8_t)g1; s.gap[1] = (uint8_t)g2; s.gap[2] = 0;
    s.npipes = 2;
}

// checkCrash's pipe part (wrapped_flappy_bird.py:254-275) + pixelCollision (:278-300) on bit rows.
// TB: HitTables in shared memory (the step kernel) or ExactTables in global / host memory (debug entry points): plain loads
template <class TB>
FB_HD bool hits_pipe(const EnvState &s, const TB *ex) {
    bool hit = false;
    for (int k = 0; k < s.npipes; k++) {
        int x = s.px[k];
        if (x >= kPlayerX + kBirdW || x + kPipeW <= kPlayerX) continue;      // rects do not overlap in x
        int gapY = 100 + 10 * s.gap[k], y = s.y, sh = kPlayerX - x;          // pipe column = bird column + sh
        int y_end = fb_min(y + kBirdH, gapY);                                   // upper pipe rows [gapY-320, gapY)
        for (int Y = y; Y < y_end; Y++) {
            unsigned long long b = ex->birdRow[s.pidx][Y - y], p = ex->pipeRowUp[Y - (gapY - kPipeH)];
            p = sh >= 0 ? p >> sh : p << (-sh);
            hit |= (b & p) != 0;
        }
        int ly = gapY + kGapSize;                                            // lower pipe rows [gapY+100, ...)
        for (int Y = fb_max(y, ly); Y < y + kBirdH; Y++) {
            unsigned long long b = ex->birdRow[s.pidx][Y - y], p = ex->pipeRowLo[Y - ly];
            p = sh >= 0 ? p >> sh : p << (-sh);
            hit |= (b & p) != 0;
        }
    }
    return hit;
}

// one frame_step (wrapped_flappy_bird.py:95-162), integer restatement of SURVEY appendix A
template <class TB>
FB_HD void env_step(EnvState &s, int action, const GapSource &g, const TB *ex,
                                         float &reward, uint8_t &terminal, int32_t &score_out) {
    reward = 0.1f;
    int vel = s.vel;
    if (action == 1) vel = -9;                       // :105-108 (playery > -48 always holds)
    else if (vel < 10) vel += 1;                     // :110-111
    s.vel = (int8_t)vel;
    int ny = s.y + vel;                              // :115-117
    bool ground = ny >= kGroundY;                    // == checkCrash's BASEY test (:251)
    if (!ground) s.y = (int16_t)fb_max(ny, 0);
    int loop1 = s.loop + 1;                          // :120-123
    if (loop1 % 3 == 0) { s.pidx = (s.phase == 3) ? 1 : s.phase; s.phase = (s.phase + 1) & 3; }
    s.loop = (uint8_t)(loop1 == 30 ? 0 : loop1);
    s.basex = (int8_t)(-((100 - s.basex) % kBaseShift));
    for (int k = 0; k < 3; k++) if (k < s.npipes) s.px[k] -= 4;                    // :126-128
    if (s.px[0] > 0 && s.px[0] < 5) {                // :131-134 (npipes is 2 here)
        int gi = draw_gap(s, g);
        s.px[2] = kScreenW + 10; s.gap[2] = (uint8_t)gi; s.npipes = 3;
    }
    if (s.px[0] < -kPipeW) {                         // :137-139
        s.px[0] = s.px[1]; s.px[1] = s.px[2]; s.gap[0] = s.gap[1]; s.gap[1] = s.gap[2];
        s.px[2] = 0; s.gap[2] = 0; s.npipes -= 1;
    }
    for (int k = 0; k < 3; k++)                      // :142-148  pipe.x in {45..48}
        if (k < s.npipes && s.px[k] >= 45 && s.px[k] <= 48) { s.score += 1; reward = 3.0f; }
    bool crash = ground || hits_pipe(s, ex);         // :151-153
    score_out = s.score;                             // :155
    terminal = crash ? 1 : 0;
    if (crash) { env_reset(s, g); reward = -3.0f; }  // :157-162
}

FB_HD DrawList make_draw_list(const EnvState &s) {
    DrawList d;
    d.y = s.y; d.pidx = s.pidx;
    bool mixed = false;                              // can a bird pixel and a pipe pixel share a 2x2 tap footprint?
    for (int k = 0; k < 3; k++) {
        d.px[k] = s.px[k]; d.gap[k] = s.gap[k];
        if (k < s.npipes) {
            int x = s.px[k], gapY = 100 + 10 * s.gap[k];
            bool xo = x < kPlayerX + kBirdW + 1 && x + kPipeW > kPlayerX - 1;
            bool yo = gapY >= s.y || gapY + kGapSize < s.y + kBirdH + 1;
            mixed |= xo && yo;
        }
    }
    d.np_mixed = (uint8_t)(s.npipes | (mixed ? 16 : 0));
    d.pad[0] = d.pad[1] = d.pad[2] = 0;
    return d;
}

// ------------------------------------------------------------------------------- drawing

// Colour of screen pixel (X, Y) after the blits of wrapped_flappy_bird.py:165-175, packed r|g<<8|b<<16.
// Painter's order resolved top-down: player, base, pipes, black background.
FB_HD uint32_t scene_rgb(const ExactTables *ex, const DrawList &d, int basex, bool with_base, int X, int Y) {
    int bx = X - kPlayerX, by = Y - d.y;
    if ((unsigned)bx < (unsigned)kBirdW && (unsigned)by < (unsigned)kBirdH) {
        uint32_t p = FB_LDG(&ex->birdPix[d.pidx][bx][by]);
        if (p >> 24) return p & 0xFFFFFFu;
    }
    if (with_base && Y >= kBaseYDraw) return FB_LDG(&ex->basePix[X - basex][Y - kBaseYDraw]) & 0xFFFFFFu;
    int np = d.np_mixed & 15;
    for (int k = 0; k < np; k++) {
        int c = X - d.px[k];
        if ((unsigned)c >= (unsigned)kPipeW) continue;
        int gapY = 100 + 10 * d.gap[k];
        uint32_t p = 0;
        if (Y < gapY) { int r = Y - (gapY - kPipeH); if (r >= 0) p = FB_LDG(&ex->pipeUp[c][r]); }
        else if (Y >= gapY + kGapSize) { int r = Y - (gapY + kGapSize); if (r < kPipeH) p = FB_LDG(&ex->pipeLo[c][r]); }
        if (p >> 24) return p & 0xFFFFFFu;
    }
    return 0u;
}

FB_HD int exact_obs_bit(const ExactTables *ex, const DrawList &d, int basex, bool with_base, int i, int j) {
    int sx = FB_LDG(&ex->sx[i]), sy = FB_LDG(&ex->sy[j]);
    uint32_t p00 = scene_rgb(ex, d, basex, with_base, sx, sy), p01 = scene_rgb(ex, d, basex, with_base, sx, sy + 1);
    uint32_t p10 = scene_rgb(ex, d, basex, with_base, sx + 1, sy), p11 = scene_rgb(ex, d, basex, with_base, sx + 1, sy + 1);
    return obs_pixel_gt1(p00, p01, p10, p11, FB_LDG(&ex->a0[i]), FB_LDG(&ex->a1[i]), FB_LDG(&ex->b0[j]), FB_LDG(&ex->b1[j]));
}


// Row i of the observation as a 64-bit mask (bit j <=> obs[i][j] == 255 for j < 64; columns 63..79 are
// the base strip and always 255).  Table path: one lookup per pipe that the row's two source
// columns touch, one 8-column bird window for rows 16..24.
FB_HD unsigned long long obs_row_mask(const ObsTables &T, const DrawList &d, int i) {
    const int np = d.np_mixed & 15;
    int q5 = (i * 205) >> 10, ph5 = i - 5 * q5, sx = T.sx[i];
    unsigned long long m = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        unsigned c = (unsigned)(sx - d.px[k] + 1);
        if (k < np && c < 54u) m |= T.pipeObs[c][ph5][d.gap[k]];
    }
    unsigned r = (unsigned)(i - 16);
    if (r < (unsigned)kBirdRows) m |= (unsigned long long)T.birdObs[d.pidx][d.y][r] << T.birdJ0[d.y];
    return m | (1ull << 63);
}

// The 8-column bird window of row 16 + r re-evaluated per pixel (rows where bird and pipe pixels can
// share a tap footprint): returns the row mask with bits [j0, j0+8) replaced.
FB_HD unsigned long long obs_row_fix(unsigned long long m, unsigned byte, int j0) {
    m = (m & ~(0xFFull << j0)) | ((unsigned long long)byte << j0);
    return m | (1ull << 63);
}

// EnvState <-> the 16-int export order of fb_env_export_state
FB_HD void state_to_ints(const EnvState &s, int32_t *o) {
    o[0] = s.y; o[1] = s.vel; o[2] = s.pidx; o[3] = s.loop; o[4] = s.phase; o[5] = s.basex; o[6] = s.score; o[7] = s.npipes;
    for (int k = 0; k < 3; k++) { o[8 + k] = k < s.npipes ? s.px[k] : 0; o[11 + k] = k < s.npipes ? s.gap[k] : 0; }
    o[14] = (int32_t)s.draws; o[15] = 0;
}

FB_HD bool ints_to_state(const int32_t *o, EnvState &s) {
    s.y = (int16_t)o[0]; s.vel = (int8_t)o[1]; s.pidx = (uint8_t)o[2]; s.loop = (uint8_t)o[3]; s.phase = (uint8_t)o[4];
    s.basex = (int8_t)o[5]; s.npipes = (uint8_t)o[7];
    for (int k = 0; k < 3; k++) { s.px[k] = (int16_t)o[8 + k]; s.gap[k] = (uint8_t)o[11 + k]; }
    s.pad0 = 0; s.pad1 = 0; s.pad2 = 0;
    s.score = o[6]; s.draws = (uint32_t)o[14];
    bool ok = o[0] >= 0 && o[0] <= kMaxY && o[2] >= 0 && o[2] < 3 && o[3] >= 0 && o[3] < 30 && o[4] >= 0 && o[4] < 4 &&
              o[5] > -kBaseShift && o[5] <= 0 && o[7] >= 2 && o[7] <= 3 && o[1] >= -9 && o[1] <= 10;
    for (int k = 0; k < 3; k++) ok = ok && o[11 + k] >= 0 && o[11 + k] < 8 && o[8 + k] >= -60 && o[8 + k] <= 440;
    return ok;
}
