/*
 * flappy_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference hot path
 *   game/wrapped_flappy_bird.py  (GameState.__init__, frame_step, getRandomPipe,
 *                                 checkCrash, pixelCollision)
 *   game/flappy_bird_utils.py    (getHitmask)
 *   FlappyBirdDQN.py:31-34       (preprocess: cv2.resize / cvtColor / threshold)
 * of angela000/DQNFlappyBird.  Every function cites the lines it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libflappy_b200.so)
 * never does.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_*.py) against
 *   - trajectories + full 288x512x3 frames produced by the reference's own
 *     game/ modules run VERBATIM on a pygame shim (tests/golden/make_golden.py,
 *     fixtures committed under tests/golden/),
 *   - the reference's five run logs (ACTION/REWARD/SCORE sequences),
 *   - real cv2 4.13 for the preprocess stage.
 *
 * It deliberately does the work the reference does: it composes the full
 * 288x512x3 frame with painter's-order blits and then resizes it, so it can
 * double as the "port" CPU baseline.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SCREENWIDTH 288            /* wrapped_flappy_bird.py:16 */
#define SCREENHEIGHT 512           /* wrapped_flappy_bird.py:17 */
#define PIPEGAPSIZE 100            /* wrapped_flappy_bird.py:43 */
#define MAX_PIPES 4

/* ------------------------------------------------------------------ assets */

typedef struct {
    int w, h;
    uint8_t *rgba;                 /* [x][y][4] like Surface.get_at((x,y)) */
    uint8_t *mask;                 /* [x][y], getHitmask: bool(alpha) */
} fo_sprite;

static fo_sprite g_bird[3], g_pipe[2], g_base;   /* g_pipe[0] = rotated 180 (upper) */
static int g_loaded = 0;

static void sprite_alloc(fo_sprite *s, int w, int h) {
    s->w = w; s->h = h;
    s->rgba = (uint8_t *)malloc((size_t)w * h * 4);
    s->mask = (uint8_t *)malloc((size_t)w * h);
}

/* flappy_bird_utils.py:103-124  mask[x][y] = bool(image.get_at((x,y))[3]) */
static void sprite_hitmask(fo_sprite *s) {
    for (int x = 0; x < s->w; x++)
        for (int y = 0; y < s->h; y++)
            s->mask[x * s->h + y] = s->rgba[(x * s->h + y) * 4 + 3] != 0;
}

static uint32_t rd32(const uint8_t *p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24; }

/* blob layout: dqnflappybird_b200/assets.py */
int fo_assets_load(const uint8_t *blob, size_t n) {
    if (n < 32 || memcmp(blob, "FBPK", 4) != 0 || rd32(blob + 4) != 1) return -1;
    int bw = rd32(blob + 8), bh = rd32(blob + 12), pw = rd32(blob + 16), ph = rd32(blob + 20);
    int sw = rd32(blob + 24), sh = rd32(blob + 28);
    size_t need = 32 + (size_t)3 * bw * bh * 4 + (size_t)pw * ph * 4 + (size_t)sw * sh * 4;
    if (need != n) return -2;
    const uint8_t *p = blob + 32;
    for (int k = 0; k < 3; k++) {          /* flappy_bird_utils.py:59-63 */
        sprite_alloc(&g_bird[k], bw, bh);
        memcpy(g_bird[k].rgba, p, (size_t)bw * bh * 4); p += (size_t)bw * bh * 4;
        sprite_hitmask(&g_bird[k]);
    }
    sprite_alloc(&g_pipe[1], pw, ph);      /* flappy_bird_utils.py:66-70 */
    memcpy(g_pipe[1].rgba, p, (size_t)pw * ph * 4); p += (size_t)pw * ph * 4;
    sprite_alloc(&g_pipe[0], pw, ph);      /* transform.rotate(pipe, 180) */
    for (int x = 0; x < pw; x++)
        for (int y = 0; y < ph; y++)
            memcpy(&g_pipe[0].rgba[(x * ph + y) * 4], &g_pipe[1].rgba[((pw - 1 - x) * ph + (ph - 1 - y)) * 4], 4);
    sprite_hitmask(&g_pipe[0]); sprite_hitmask(&g_pipe[1]);
    sprite_alloc(&g_base, sw, sh);         /* flappy_bird_utils.py:54 */
    memcpy(g_base.rgba, p, (size_t)sw * sh * 4);
    sprite_hitmask(&g_base);
    g_loaded = 1;
    return 0;
}

/* --------------------------------------------------------------- gap source */

/* Philox4x32-10 (Salmon et al., SC'11; Random123 constants). */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void fo_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32_10(ctr, key, out); }

/* word n of stream (seed, purpose, env): block n>>2 of counter (block, purpose, env_lo, env_hi) */
uint32_t fo_stream_word(uint64_t seed, uint32_t purpose, uint64_t env, uint32_t n) {
    uint32_t ctr[4] = {n >> 2, purpose, (uint32_t)env, (uint32_t)(env >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, out[4];
    philox4x32_10(ctr, key, out);
    return out[n & 3];
}

typedef struct {
    /* wrapped_flappy_bird.py:59-85 */
    int score, playerIndex, loopIter;
    double playery;                /* Python int until the BASEY float branch of :115 */
    int playerVelY, playerFlapped;
    int basex;
    int nPipes;
    int pipeX[MAX_PIPES], upperY[MAX_PIPES], lowerY[MAX_PIPES];
    /* PLAYER_INDEX_GEN, a module global that survives resets (:52) */
    int cyclePhase;
    /* gap source: scripted (replay mode) or Philox stream */
    const uint8_t *gaps; int gapsLen;
    uint64_t seed, envId;
    uint32_t draws;                /* scripted: gaps consumed; philox: words consumed */
    int lastGap[MAX_PIPES];        /* bookkeeping for state export */
} fo_env;

/* random.randint(0, 7): CPython _randbelow(8) = getrandbits(4) with rejection,
 * getrandbits(4) = top 4 bits of one 32-bit word.  wrapped_flappy_bird.py:212 */
static int next_gap_index(fo_env *e) {
    if (e->gaps) { int g = e->gaps[e->draws % (uint32_t)e->gapsLen]; e->draws++; return g & 7; }
    for (;;) {
        uint32_t w = fo_stream_word(e->seed, 0, e->envId, e->draws++);
        uint32_t r = w >> 28;
        if (r < 8) return (int)r;
    }
}

/* wrapped_flappy_bird.py:208-221 getRandomPipe */
static void get_random_pipe(fo_env *e, int *pipeX, int *uy, int *ly, int *gidx) {
    static const int gapYs[8] = {20, 30, 40, 50, 60, 70, 80, 90};
    const double BASEY = SCREENHEIGHT * 0.79;
    int index = next_gap_index(e);
    int gapY = gapYs[index];
    gapY += (int)(BASEY * 0.2);
    *pipeX = SCREENWIDTH + 10;
    *uy = gapY - g_pipe[0].h;
    *ly = gapY + PIPEGAPSIZE;
    *gidx = index;
}

/* wrapped_flappy_bird.py:59-85 GameState.__init__ (cyclePhase untouched) */
static void env_init(fo_env *e) {
    e->score = e->playerIndex = e->loopIter = 0;
    e->playery = (int)((SCREENHEIGHT - g_bird[0].h) / 2);
    e->basex = 0;
    int px, uy, ly, gi;
    get_random_pipe(e, &px, &uy, &ly, &gi);
    e->pipeX[0] = SCREENWIDTH; e->upperY[0] = uy; e->lowerY[0] = ly; e->lastGap[0] = gi;
    get_random_pipe(e, &px, &uy, &ly, &gi);
    e->pipeX[1] = SCREENWIDTH + SCREENWIDTH / 2; e->upperY[1] = uy; e->lowerY[1] = ly; e->lastGap[1] = gi;
    e->nPipes = 2;
    e->playerVelY = 0;
    e->playerFlapped = 0;
}

fo_env *fo_env_create(uint64_t seed, uint64_t env_id, const uint8_t *gaps, int gaps_len) {
    if (!g_loaded) return NULL;
    fo_env *e = (fo_env *)calloc(1, sizeof(fo_env));
    e->seed = seed; e->envId = env_id;
    if (gaps && gaps_len > 0) {
        uint8_t *g = (uint8_t *)malloc(gaps_len);
        memcpy(g, gaps, gaps_len);
        e->gaps = g; e->gapsLen = gaps_len;
    }
    e->cyclePhase = 0;
    env_init(e);
    return e;
}

void fo_env_destroy(fo_env *e) { if (e) { free((void *)e->gaps); free(e); } }

typedef struct { int x, y, w, h; } fo_rect;

/* pygame.Rect.clip */
static fo_rect rect_clip(fo_rect a, fo_rect b) {
    fo_rect r = {a.x, a.y, 0, 0};
    int x0 = a.x > b.x ? a.x : b.x, y0 = a.y > b.y ? a.y : b.y;
    int x1 = (a.x + a.w < b.x + b.w) ? a.x + a.w : b.x + b.w;
    int y1 = (a.y + a.h < b.y + b.h) ? a.y + a.h : b.y + b.h;
    if (x1 > x0 && y1 > y0) { r.x = x0; r.y = y0; r.w = x1 - x0; r.h = y1 - y0; }
    return r;
}

/* wrapped_flappy_bird.py:278-300 pixelCollision */
static int pixel_collision(fo_rect r1, fo_rect r2, const fo_sprite *s1, const fo_sprite *s2) {
    fo_rect r = rect_clip(r1, r2);
    if (r.w == 0 || r.h == 0) return 0;
    int x1 = r.x - r1.x, y1 = r.y - r1.y, x2 = r.x - r2.x, y2 = r.y - r2.y;
    for (int x = 0; x < r.w; x++)
        for (int y = 0; y < r.h; y++)
            if (s1->mask[(x1 + x) * s1->h + y1 + y] && s2->mask[(x2 + x) * s2->h + y2 + y]) return 1;
    return 0;
}

/* wrapped_flappy_bird.py:244-275 checkCrash */
static int check_crash(const fo_env *e) {
    const double BASEY = SCREENHEIGHT * 0.79;
    int pi = e->playerIndex, pw = g_bird[0].w, ph = g_bird[0].h;
    if (e->playery + ph >= BASEY - 1) return 1;
    fo_rect pr = {(int)(SCREENWIDTH * 0.2), (int)e->playery, pw, ph};
    for (int k = 0; k < e->nPipes; k++) {
        fo_rect ur = {e->pipeX[k], e->upperY[k], g_pipe[0].w, g_pipe[0].h};
        fo_rect lr = {e->pipeX[k], e->lowerY[k], g_pipe[1].w, g_pipe[1].h};
        if (pixel_collision(pr, ur, &g_bird[pi], &g_pipe[0]) ||
            pixel_collision(pr, lr, &g_bird[pi], &g_pipe[1])) return 1;
    }
    return 0;
}

/* wrapped_flappy_bird.py:87-162 frame_step up to (not including) drawing.
 * action: 0 = input_actions [1,0] (no-op), 1 = [0,1] (flap).
 * returns 0, or -1 for an action that is not one-hot (:99-100 ValueError). */
int fo_env_step(fo_env *e, int action, float *reward, uint8_t *terminal, int32_t *score) {
    static const int cyc[4] = {0, 1, 2, 1};
    const double BASEY = SCREENHEIGHT * 0.79;
    const int PLAYER_HEIGHT = g_bird[0].h, PLAYER_WIDTH = g_bird[0].w, PIPE_WIDTH = g_pipe[0].w;
    if (action != 0 && action != 1) return -1;
    double rew = 0.1; int term = 0;
    if (action == 1) {                                           /* :105-108 */
        if (e->playery > -2 * PLAYER_HEIGHT) { e->playerVelY = -9; e->playerFlapped = 1; }
    }
    if (e->playerVelY < 10 && !e->playerFlapped) e->playerVelY += 1;   /* :110-111 */
    if (e->playerFlapped) e->playerFlapped = 0;                  /* :112-113 */
    {                                                            /* :115-117 */
        double room = BASEY - e->playery - PLAYER_HEIGHT;
        e->playery += (e->playerVelY < room) ? e->playerVelY : room;
        if (e->playery < 0) e->playery = 0;
    }
    if ((e->loopIter + 1) % 3 == 0) {                            /* :120-121 */
        e->playerIndex = cyc[e->cyclePhase]; e->cyclePhase = (e->cyclePhase + 1) & 3;
    }
    e->loopIter = (e->loopIter + 1) % 30;                        /* :122 */
    {                                                            /* :123, Python % is floor-mod */
        int baseShift = g_base.w - SCREENWIDTH;
        int m = (-e->basex + 100) % baseShift; if (m < 0) m += baseShift;
        e->basex = -m;
    }
    for (int k = 0; k < e->nPipes; k++) e->pipeX[k] += -4;       /* :126-128 */
    if (0 < e->pipeX[0] && e->pipeX[0] < 5) {                    /* :131-134 */
        int k = e->nPipes, gi;
        if (k >= MAX_PIPES) abort();
        get_random_pipe(e, &e->pipeX[k], &e->upperY[k], &e->lowerY[k], &gi);
        e->lastGap[k] = gi;
        e->nPipes++;
    }
    if (e->pipeX[0] < -PIPE_WIDTH) {                             /* :137-139 */
        for (int k = 1; k < e->nPipes; k++) {
            e->pipeX[k - 1] = e->pipeX[k]; e->upperY[k - 1] = e->upperY[k];
            e->lowerY[k - 1] = e->lowerY[k]; e->lastGap[k - 1] = e->lastGap[k];
        }
        e->nPipes--;
    }
    {                                                            /* :142-148 */
        double playerMidPos = (int)(SCREENWIDTH * 0.2) + PLAYER_WIDTH / 2.0;
        for (int k = 0; k < e->nPipes; k++) {
            double pipeMidPos = e->pipeX[k] + PIPE_WIDTH / 2.0;
            if (pipeMidPos <= playerMidPos && playerMidPos < pipeMidPos + 4) { e->score += 1; rew = 3; }
        }
    }
    int crash = check_crash(e);                                  /* :151-153 */
    int score_return = e->score;                                 /* :155 */
    if (crash) { term = 1; env_init(e); rew = -3; }              /* :157-162 */
    *reward = (float)rew; *terminal = (uint8_t)term; *score = score_return;
    return 0;
}

/* state export, field order shared with fb_env_export_state (include/flappy_b200.h):
 * [0]=playery [1]=velY [2]=playerIndex [3]=loopIter [4]=cyclePhase [5]=basex [6]=score
 * [7]=nPipes [8..10]=pipe x [11..13]=gap index (unused slots 0) [14]=rng draws [15]=0 */
void fo_env_export_state(const fo_env *e, int32_t *out) {
    memset(out, 0, 16 * sizeof(int32_t));
    out[0] = (int32_t)e->playery; out[1] = e->playerVelY; out[2] = e->playerIndex; out[3] = e->loopIter;
    out[4] = e->cyclePhase; out[5] = e->basex; out[6] = e->score; out[7] = e->nPipes;
    for (int k = 0; k < e->nPipes && k < 3; k++) { out[8 + k] = e->pipeX[k]; out[11 + k] = e->lastGap[k]; }
    out[14] = (int32_t)e->draws;
}

/* test helper: overwrite the state (same field order as fo_env_export_state) */
void fo_env_import_state(fo_env *e, const int32_t *in) {
    e->playery = in[0]; e->playerVelY = in[1]; e->playerIndex = in[2]; e->loopIter = in[3];
    e->cyclePhase = in[4]; e->basex = in[5]; e->score = in[6]; e->nPipes = in[7];
    e->playerFlapped = 0;
    for (int k = 0; k < e->nPipes && k < 3; k++) {
        int gapY = 100 + 10 * in[11 + k];
        e->pipeX[k] = in[8 + k]; e->lastGap[k] = in[11 + k];
        e->upperY[k] = gapY - g_pipe[0].h; e->lowerY[k] = gapY + PIPEGAPSIZE;
    }
    e->draws = (uint32_t)in[14];
}

/* ------------------------------------------------------------------ drawing */

/* Surface.blit of a binary-alpha sprite: copy where alpha != 0, clipped to the
 * 288x512 screen.  Destination coordinates truncate to int ([assumed], see
 * SURVEY 8c).  frame is surfarray.array3d layout u8[288][512][3]. */
static void blit(uint8_t *frame, const fo_sprite *s, int dx, int dy) {
    for (int x = 0; x < s->w; x++) {
        int X = dx + x; if (X < 0 || X >= SCREENWIDTH) continue;
        for (int y = 0; y < s->h; y++) {
            int Y = dy + y; if (Y < 0 || Y >= SCREENHEIGHT) continue;
            const uint8_t *p = &s->rgba[(x * s->h + y) * 4];
            if (p[3]) { uint8_t *q = &frame[(X * SCREENHEIGHT + Y) * 3]; q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; }
        }
    }
}

/* wrapped_flappy_bird.py:165-177: background, pipes, base, player -> array3d */
void fo_env_render(const fo_env *e, uint8_t *frame) {
    const double BASEY = SCREENHEIGHT * 0.79;
    memset(frame, 0, (size_t)SCREENWIDTH * SCREENHEIGHT * 3);    /* black background :165 */
    for (int k = 0; k < e->nPipes; k++) {                        /* :167-169 */
        blit(frame, &g_pipe[0], e->pipeX[k], e->upperY[k]);
        blit(frame, &g_pipe[1], e->pipeX[k], e->lowerY[k]);
    }
    blit(frame, &g_base, e->basex, (int)BASEY);                  /* :171 */
    blit(frame, &g_bird[e->playerIndex], (int)(SCREENWIDTH * 0.2), (int)e->playery);  /* :174-175 */
}

/* ------------------------------------------------------------- preprocess */

#define OBS 80
static int g_sx[OBS], g_a0[OBS], g_a1[OBS];   /* rows of the cv2 image  = game x (288) */
static int g_sy[OBS], g_b0[OBS], g_b1[OBS];   /* cols of the cv2 image  = game y (512) */
static int g_tabs = 0;

/* cv2 resize (imgproc/src/resize.cpp, INTER_LINEAR, 8U): scale = 1/(dst/src) in
 * double; fx = (float)((d+0.5)*scale-0.5); s = floor(fx); fx -= s; coefficients
 * saturate_cast<short>((1-fx)*2048), (fx*2048) (round half to even). */
static void coef_table(int src, int dst, int *s0, int *c0, int *c1) {
    double inv_scale = (double)dst / src, scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(fx);
        fx -= s;
        if (s < 0) { fx = 0; s = 0; }
        if (s >= src - 1) { fx = 0; s = src - 1; }
        s0[d] = s;
        c0[d] = (int)lrintf((1.f - fx) * 2048.f);
        c1[d] = (int)lrintf(fx * 2048.f);
    }
}

void fo_resize_tables(int32_t *out /* [6][80]: sx a0 a1 sy b0 b1 */) {
    if (!g_tabs) { coef_table(SCREENWIDTH, OBS, g_sx, g_a0, g_a1); coef_table(SCREENHEIGHT, OBS, g_sy, g_b0, g_b1); g_tabs = 1; }
    if (out) for (int i = 0; i < OBS; i++) {
        out[i] = g_sx[i]; out[80 + i] = g_a0[i]; out[160 + i] = g_a1[i];
        out[240 + i] = g_sy[i]; out[320 + i] = g_b0[i]; out[400 + i] = g_b1[i];
    }
}

/* FlappyBirdDQN.py:31-34: threshold(cvtColor(resize(frame,(80,80)), BGR2GRAY), 1, 255).
 * frame u8[288][512][3] (rows = game x), out u8[80][80] (out[i][j]: i ~ x, j ~ y). */
void fo_preprocess(const uint8_t *frame, uint8_t *out) {
    fo_resize_tables(NULL);
    for (int i = 0; i < OBS; i++) {
        const uint8_t *r0 = frame + (size_t)g_sx[i] * SCREENHEIGHT * 3;
        int i1 = g_sx[i] + 1 < SCREENWIDTH ? g_sx[i] + 1 : SCREENWIDTH - 1;
        const uint8_t *r1 = frame + (size_t)i1 * SCREENHEIGHT * 3;
        for (int j = 0; j < OBS; j++) {
            int y0 = g_sy[j], y1 = y0 + 1 < SCREENHEIGHT ? y0 + 1 : SCREENHEIGHT - 1;
            int v[3];
            for (int c = 0; c < 3; c++) {
                int h0 = r0[y0 * 3 + c] * g_b0[j] + r0[y1 * 3 + c] * g_b1[j];   /* HResizeLinear */
                int h1 = r1[y0 * 3 + c] * g_b0[j] + r1[y1 * 3 + c] * g_b1[j];
                v[c] = (((g_a0[i] * (h0 >> 4)) >> 16) + ((g_a1[i] * (h1 >> 4)) >> 16) + 2) >> 2;  /* VResizeLinear */
            }
            /* COLOR_BGR2GRAY on RGB data: channel 0 gets the B weight */
            int g = (3735 * v[0] + 19235 * v[1] + 9798 * v[2] + (1 << 14)) >> 15;
            out[i * OBS + j] = g > 1 ? 255 : 0;
        }
    }
}

/* frame_step + preprocess for one env */
void fo_env_obs(const fo_env *e, uint8_t *frame_scratch, uint8_t *out) {
    fo_env_render(e, frame_scratch);
    fo_preprocess(frame_scratch, out);
}

/* ------------------------------------------------------------- batch helpers
 * Used by the parity tests and by the CPU baseline (one fo_env per reference
 * process; threads stand in for processes). */
typedef struct {
    fo_env **envs; int lo, hi; const uint8_t *actions; float *reward; uint8_t *terminal;
    int32_t *score; uint8_t *obs; int rc;
} fo_job;

static void *batch_worker(void *arg) {
    fo_job *j = (fo_job *)arg;
    uint8_t *frame = (uint8_t *)malloc((size_t)SCREENWIDTH * SCREENHEIGHT * 3);
    for (int k = j->lo; k < j->hi; k++) {
        if (fo_env_step(j->envs[k], j->actions[k], &j->reward[k], &j->terminal[k], &j->score[k]) != 0) j->rc = -1;
        if (j->obs) fo_env_obs(j->envs[k], frame, j->obs + (size_t)k * OBS * OBS);
    }
    free(frame);
    return NULL;
}

/* T consecutive frame_steps of ONE env (long golden trajectories): per-step reward / terminal / score / exported state, and the
 * preprocessed observation after every step whose want_obs flag is set (packed in order into obs_out) */
int fo_env_run(fo_env *e, int T, const uint8_t *actions, const uint8_t *want_obs, float *reward, uint8_t *terminal, int32_t *score,
               int32_t *state16 /* [T][16] */, uint8_t *obs_out /* [n_wanted][80][80] */) {
    uint8_t *frame = (uint8_t *)malloc((size_t)SCREENWIDTH * SCREENHEIGHT * 3);
    size_t k = 0;
    int rc = 0;
    for (int t = 0; t < T; t++) {
        if (fo_env_step(e, actions[t], &reward[t], &terminal[t], &score[t]) != 0) { rc = -1; break; }
        fo_env_export_state(e, state16 + (size_t)t * 16);
        if (want_obs && want_obs[t]) fo_env_obs(e, frame, obs_out + (k++) * OBS * OBS);
    }
    free(frame);
    return rc;
}

int fo_batch_step(fo_env **envs, int n, const uint8_t *actions, float *reward, uint8_t *terminal,
                  int32_t *score, uint8_t *obs /* [n][80][80] or NULL */, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads > n) n_threads = n > 0 ? n : 1;
    pthread_t th[256]; fo_job jobs[256];
    int rc = 0;
    for (int t = 0; t < n_threads; t++) {
        fo_job j = {envs, (int)((long long)n * t / n_threads), (int)((long long)n * (t + 1) / n_threads),
                    actions, reward, terminal, score, obs, 0};
        jobs[t] = j;
        if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    batch_worker(&jobs[0]);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < n_threads; t++) if (jobs[t].rc) rc = -1;
    return rc;
}
