// Asset loading and table derivation (host side).
//
// Replaces flappy_bird_utils.load() + getHitmask (game/flappy_bird_utils.py:16-124) and the
// module-level constants of game/wrapped_flappy_bird.py:42-50.  Everything the kernels look
// up is derived here from the six sprites with the same fixed-point arithmetic cv2 uses
// (obs_pixel_gt1), so the tables are exact by construction; the GPU tests additionally check
// them exhaustively against the oracle.
#include <math.h>
#include <string.h>

#include "fb_common.cuh"

static thread_local std::string g_err;
void fb_set_error(const std::string &msg) { g_err = msg; }
extern "C" const char *fb_last_error(void) { return g_err.c_str(); }
extern "C" int fb_version(void) { return 1; }

static FbTables g_tables = {nullptr, nullptr, -1, false};
static ExactTables g_exact_host;
static ObsTables g_obs_host;
static bool g_host_ready = false;
const FbTables &fb_tables() { return g_tables; }
const ObsTables *fb_host_obs_tables() { return g_host_ready ? &g_obs_host : nullptr; }
const ExactTables *fb_host_exact_tables() { return g_host_ready ? &g_exact_host : nullptr; }

// cv2 resize coefficient table (imgproc resize.cpp, INTER_LINEAR 8U):
//   scale = 1 / (dst / src) in double; f = (float)((d + .5) * scale - .5); s = floor(f); f -= s;
//   c0 = saturate_cast<short>((1 - f) * 2048), c1 = saturate_cast<short>(f * 2048)
static void coef_table(int src, int dst, int *s0, int *c0, int *c1) {
    double inv_scale = (double)dst / src, scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src - 1) { f = 0.f; s = src - 1; }
        s0[d] = s;
        c0[d] = (int)lrintf((1.f - f) * 2048.f);
        c1[d] = (int)lrintf(f * 2048.f);
    }
}

static inline uint32_t rd32(const uint8_t *p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24; }
static inline uint32_t rgb_or_black(uint32_t rgba) { return (rgba >> 24) ? (rgba & 0xFFFFFFu) : 0u; }

extern "C" int fb_resize_tables(int32_t *out) {
    if (!g_host_ready) { fb_set_error("fb_assets_load has not been called"); return FB_ERR_ASSETS; }
    const ExactTables &E = g_exact_host;
    for (int i = 0; i < kObs; i++) {
        out[i] = E.sx[i]; out[80 + i] = E.a0[i]; out[160 + i] = E.a1[i];
        out[240 + i] = E.sy[i]; out[320 + i] = E.b0[i]; out[400 + i] = E.b1[i];
    }
    return FB_OK;
}

static int build_host_tables(const uint8_t *blob, size_t n) {
    g_host_ready = false;
    if (!blob || n < 32 || memcmp(blob, "FBPK", 4) != 0 || rd32(blob + 4) != 1) {
        fb_set_error("fb_assets_load: not an FBPK v1 blob");
        return FB_ERR_ASSETS;
    }
    int bw = rd32(blob + 8), bh = rd32(blob + 12), pw = rd32(blob + 16), ph = rd32(blob + 20);
    int sw = rd32(blob + 24), sh = rd32(blob + 28);
    if (bw != kBirdW || bh != kBirdH || pw != kPipeW || ph != kPipeH || sw != kBaseW || sh != kBaseH) {
        fb_set_error("fb_assets_load: sprite sizes differ from the reference's (34x24, 52x320, 336x112)");
        return FB_ERR_ASSETS;
    }
    size_t need = 32 + (size_t)4 * (3 * bw * bh + pw * ph + sw * sh);
    if (need != n) { fb_set_error("fb_assets_load: truncated blob"); return FB_ERR_ASSETS; }

    ExactTables &E = g_exact_host;
    memset(&E, 0, sizeof(E));
    const uint8_t *p = blob + 32;
    for (int k = 0; k < 3; k++)
        for (int x = 0; x < bw; x++)
            for (int y = 0; y < bh; y++, p += 4) E.birdPix[k][x][y] = rd32(p);
    for (int x = 0; x < pw; x++)
        for (int y = 0; y < ph; y++, p += 4) E.pipeLo[x][y] = rd32(p);
    for (int x = 0; x < sw; x++)
        for (int y = 0; y < sh; y++, p += 4) E.basePix[x][y] = rd32(p);
    // pygame.transform.rotate(pipe, 180), flappy_bird_utils.py:66-68
    for (int x = 0; x < pw; x++)
        for (int y = 0; y < ph; y++) E.pipeUp[x][y] = E.pipeLo[pw - 1 - x][ph - 1 - y];
    // getHitmask (flappy_bird_utils.py:103-124) as bit rows
    for (int k = 0; k < 3; k++)
        for (int y = 0; y < bh; y++) {
            unsigned long long m = 0;
            for (int x = 0; x < bw; x++) if (E.birdPix[k][x][y] >> 24) m |= 1ull << x;
            E.birdRow[k][y] = m;
        }
    for (int y = 0; y < ph; y++) {
        unsigned long long lo = 0, up = 0;
        for (int x = 0; x < pw; x++) {
            if (E.pipeLo[x][y] >> 24) lo |= 1ull << x;
            if (E.pipeUp[x][y] >> 24) up |= 1ull << x;
        }
        E.pipeRowLo[y] = lo; E.pipeRowUp[y] = up;
    }
    for (int x = 0; x < sw; x++)
        for (int y = 0; y < sh; y++)
            if ((E.basePix[x][y] >> 24) == 0) { fb_set_error("fb_assets_load: base sprite must be opaque"); return FB_ERR_ASSETS; }

    coef_table(kScreenW, kObs, E.sx, E.a0, E.a1);
    coef_table(kScreenH, kObs, E.sy, E.b0, E.b1);
    for (int i = 0; i < kObs; i++) {
        if (E.a0[i] != E.a0[i % 5] || E.a1[i] != E.a1[i % 5] || E.b0[i] != E.b0[i % 5] || E.b1[i] != E.b1[i % 5] ||
            E.sx[i] + 1 >= kScreenW || E.sy[i] + 1 >= kScreenH) {
            fb_set_error("fb_assets_load: resize coefficients are not 5-periodic");
            return FB_ERR_ASSETS;
        }
    }
    // geometry the fast path relies on
    if (E.sy[kBaseJ - 1] + 1 >= kBaseYDraw || E.sy[kBaseJ] < kBaseYDraw) {
        fb_set_error("fb_assets_load: base strip does not start at obs column 63"); return FB_ERR_ASSETS;
    }
    for (int i = 0; i < kObs; i++) {
        bool touches = E.sx[i] + 1 >= kPlayerX && E.sx[i] <= kPlayerX + kBirdW - 1;
        if (touches != (i >= 16 && i < 16 + kBirdRows)) { fb_set_error("fb_assets_load: bird rows are not 16..24"); return FB_ERR_ASSETS; }
    }
    // base strip: every obs pixel with j >= 63 must be 255 for every basex
    for (int basex = -(kBaseShift - 1); basex <= 0; basex++)
        for (int i = 0; i < kObs; i++)
            for (int j = kBaseJ; j < kObs; j++) {
                uint32_t t[4];
                for (int q = 0; q < 4; q++) {
                    int X = E.sx[i] + (q >> 1), Y = E.sy[j] + (q & 1);
                    t[q] = E.basePix[X - basex][Y - kBaseYDraw] & 0xFFFFFFu;
                }
                if (!obs_pixel_gt1(t[0], t[1], t[2], t[3], E.a0[i], E.a1[i], E.b0[j], E.b1[j])) {
                    fb_set_error("fb_assets_load: base strip is not uniformly above the threshold"); return FB_ERR_ASSETS;
                }
            }

    ObsTables &O = g_obs_host;
    memset(&O, 0, sizeof(O));
    // pipe-only rows.  c0 = sx_i - pipe.x is the sprite column of the first tap.
    for (int ci = 0; ci < 54; ci++)
        for (int ph5 = 0; ph5 < 5; ph5++)
            for (int g = 0; g < 8; g++) {
                int c0 = ci - 1, gapY = 100 + 10 * g, uy = gapY - kPipeH, ly = gapY + kGapSize;
                unsigned long long bits = 0;
                for (int j = 0; j < kBaseJ; j++) {
                    uint32_t t[4];
                    for (int q = 0; q < 4; q++) {
                        int c = c0 + (q >> 1), Y = E.sy[j] + (q & 1);
                        uint32_t v = 0;
                        if (c >= 0 && c < kPipeW) {
                            if (Y - uy >= 0 && Y - uy < kPipeH) v = rgb_or_black(E.pipeUp[c][Y - uy]);
                            else if (Y - ly >= 0 && Y - ly < kPipeH) v = rgb_or_black(E.pipeLo[c][Y - ly]);
                        }
                        t[q] = v;
                    }
                    if (obs_pixel_gt1(t[0], t[1], t[2], t[3], E.a0[ph5], E.a1[ph5], E.b0[j], E.b1[j])) bits |= 1ull << j;
                }
                O.pipeObs[ci][ph5][g] = bits;
            }
    for (int i = 0; i < kObs; i++) O.sx[i] = (short)E.sx[i];
    // bird-only 8-column windows
    for (int y = 0; y <= kMaxY; y++) {
        int j0 = 0;
        while (j0 < kObs && E.sy[j0] + 1 < y) j0++;
        int j1 = j0;
        while (j1 + 1 < kObs && E.sy[j1 + 1] <= y + kBirdH - 1) j1++;
        if (j1 - j0 >= 8) { fb_set_error("fb_assets_load: bird spans more than 8 obs columns"); return FB_ERR_ASSETS; }
        O.birdJ0[y] = (unsigned char)j0;
        for (int k = 0; k < 3; k++)
            for (int r = 0; r < kBirdRows; r++) {
                int i = 16 + r;
                unsigned bits = 0;
                for (int b = 0; b < 8; b++) {
                    int j = j0 + b;
                    if (j >= kBaseJ) continue;
                    uint32_t t[4];
                    for (int q = 0; q < 4; q++) {
                        int bx = E.sx[i] + (q >> 1) - kPlayerX, by = E.sy[j] + (q & 1) - y;
                        t[q] = (bx >= 0 && bx < kBirdW && by >= 0 && by < kBirdH) ? rgb_or_black(E.birdPix[k][bx][by]) : 0u;
                    }
                    if (obs_pixel_gt1(t[0], t[1], t[2], t[3], E.a0[i], E.a1[i], E.b0[j], E.b1[j])) bits |= 1u << b;
                }
                O.birdObs[k][y][r] = (unsigned char)bits;
            }
    }

    g_host_ready = true;
    return FB_OK;
}

// test hook: derive the tables on the host only (no device needed); see fb_debug.cu
extern "C" int fb_debug_assets_load_host(const uint8_t *blob, size_t n) { return build_host_tables(blob, n); }

extern "C" int fb_assets_load(const uint8_t *blob, size_t n) {
    int rc = build_host_tables(blob, n);
    if (rc) return rc;
    const ObsTables &O = g_obs_host;
    const ExactTables &E = g_exact_host;
    int dev = 0;
    FB_CUDA_OK(cudaGetDevice(&dev));
    if (g_tables.loaded) { cudaFree(g_tables.obs_dev); cudaFree(g_tables.exact_dev); g_tables.loaded = false; }
    FB_CUDA_OK(cudaMalloc(&g_tables.obs_dev, sizeof(ObsTables)));
    FB_CUDA_OK(cudaMalloc(&g_tables.exact_dev, sizeof(ExactTables)));
    FB_CUDA_OK(cudaMemcpy(g_tables.obs_dev, &O, sizeof(ObsTables), cudaMemcpyHostToDevice));
    FB_CUDA_OK(cudaMemcpy(g_tables.exact_dev, &E, sizeof(ExactTables), cudaMemcpyHostToDevice));
    g_tables.device = dev;
    g_tables.loaded = true;
    return FB_OK;
}
