// Tensor-core path of the DQN-family Q-network (sm_100a): every convolution / fully connected layer, forward,
// data gradient and weight gradient, is ONE warp-specialised kernel shape -- TMA (cp.async.bulk.tensor) feeds
// 128B-swizzled shared-memory stages, one elected thread issues tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), four
// epilogue warps read the accumulator back with tcgen05.ld and fuse bias / ReLU / masks / layout remaps.
//
// Reference graph: BrainDQN.py:119-163; backward = tf.gradients of it (BrainDQN.py:163).
//
// What makes the convolutions plain TMA GEMMs ("implicit GEMM by row offset"): activations live in HBM as bf16
// matrices [position][channel] whose positions are the cells of a zero-padded grid linearised with a FIXED width,
// so that the input row of filter tap (dh,dw) for output position p is simply p + dh*width + dw:
//   conv1 8x8 s4 pad 2  = 2x2 s1 conv over the space-to-depth tensor X2 [B*441][64]  (21x21 blocks of 4x4x4)
//   conv2 4x4 s2 pad 1  = 2x2 s1 conv over the pooled space-to-depth tensor P2 [B*49][128] (7-wide grid of 2x2x32)
//   conv3 3x3 s1 pad 1  = 3x3 s1 conv over A2 [B*49][64] (7-wide grid, the pad ring is the invalid rows, offset -8)
// Each K-block of 64 channels of one tap is one TMA box at a shifted row coordinate; rows outside the tensor are
// zero-filled by TMA.  Invalid grid positions are computed and discarded (conv1 10 %, conv2/3 2x; those layers
// are < 30 % of the FLOPs).  Data gradients are the same kernel over the gradient tensors with re-packed weights;
// weight gradients contract over positions with both operands MN-major (the same row-major tensors, transposed
// by the descriptor, not in memory) and a deterministic split-K.
//
// Numerics: operands bf16 (inputs {0,255} are exact), accumulation fp32, activations stored bf16 between layers,
// head / TD loss / Adam in fp32 (shared with fb_qnet.cu).  Tolerances vs the float64 oracle: tests/test_qnet_tc_gpu.py.
#include <cudaTypedefs.h>

#include <map>
#include <new>
#include <vector>

#include "fb_qnet.cuh"
#include "fb_tc.cuh"

using bf16 = __nv_bfloat16;

namespace {

constexpr int kStages = 4;
constexpr int kMaxKb = 32;
constexpr int kG1 = 21, kP1 = 441;        // conv1 block grid
constexpr int kG2 = 7, kP2 = 49;          // 7-wide grid shared by conv2 / conv3 tensors
constexpr int kThreads = 192;             // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..5 epilogue

// ------------------------------------------------------------------------------------------------ the GEMM
struct GemmParams {
    // MODE 0 (K-major A and B):  D[m][n] = sum_kb sum_k A[m + a_rowoff[kb]][a_col[kb] + k] * Bt[n][64 kb + k]
    int nkb;
    int a_rowoff[kMaxKb];
    int a_col[kMaxKb];
    // MODE 1 (MN-major A and B): D[128 mt + 64 i + j][n] = sum_p A[p + a2_rowoff[mt][i]][a2_col[mt][i] + j] * B[p][n]
    int p_total, klen;
    int a2_rowoff[16][2];
    int a2_col[16][2];
    // shared-memory descriptor strides (bytes); runtime so that the self-test can probe encodings
    uint32_t lbo_a, sbo_a, kstep_a, lbo_b, sbo_b, kstep_b;
};

template <int BN, int MODE, class EP>
__global__ void __launch_bounds__(kThreads) tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA,
                                                           const __grid_constant__ CUtensorMap mapB, const GemmParams g, const EP ep) {
    constexpr uint32_t A_BYTES = 128 * 64 * 2, B_BYTES = BN * 64 * 2, STAGE = A_BYTES + B_BYTES;
    constexpr uint32_t B_LAYOUT = (MODE == 1 && BN == 32) ? tc::kSwizzle64 : tc::kSwizzle128;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kStages], bar_empty[kStages], bar_accum;
    __shared__ uint32_t tmem_slot;
    const uint32_t smem = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x, n0 = blockIdx.y * BN;

    int nkb, p_begin = 0;
    if (MODE == 0) nkb = g.nkb;
    else {
        p_begin = blockIdx.z * g.klen;
        int len = min(g.klen, g.p_total - p_begin);
        nkb = (len + 63) >> 6;
    }

    if (threadIdx.x == 0) {
        tc::tma_prefetch_desc(&mapA);
        tc::tma_prefetch_desc(&mapB);
        for (int s = 0; s < kStages; s++) { tc::mbar_init(tc::smem_u32(&bar_full[s]), 1); tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1); }
        tc::mbar_init(tc::smem_u32(&bar_accum), 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_slot), BN);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ph ^ 1u);
                const uint32_t full = tc::smem_u32(&bar_full[s]);
                const uint32_t sa = smem + s * STAGE, sb = sa + A_BYTES;
                tc::mbar_expect_tx(full, STAGE);
                if (MODE == 0) {
                    tc::tma_load_2d(sa, &mapA, g.a_col[kb], mt * 128 + g.a_rowoff[kb], full);
                    tc::tma_load_2d(sb, &mapB, kb * 64, n0, full);
                } else {
                    const int p = p_begin + kb * 64;
                    tc::tma_load_2d(sa, &mapA, g.a2_col[mt][0], p + g.a2_rowoff[mt][0], full);
                    tc::tma_load_2d(sa + 8192, &mapA, g.a2_col[mt][1], p + g.a2_rowoff[mt][1], full);
                    tc::tma_load_2d(sb, &mapB, n0, p, full);
                    if (BN == 128) tc::tma_load_2d(sb + 8192, &mapB, n0 + 64, p, full);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::instr_desc_bf16(128, BN, MODE, MODE);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                tc::mbar_wait(tc::smem_u32(&bar_full[s]), ph);
                tc::tc_fence_after();
                const uint32_t sa = smem + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint64_t ad = tc::smem_desc(sa + k * g.kstep_a, g.lbo_a, g.sbo_a, tc::kSwizzle128);
                    const uint64_t bd = tc::smem_desc(sb + k * g.kstep_b, g.lbo_b, g.sbo_b, B_LAYOUT);
                    tc::umma_bf16(tmem, ad, bd, idesc, (kb | k) != 0);
                }
                tc::umma_commit(tc::smem_u32(&bar_empty[s]));     // frees the stage once these MMAs have read it
            }
            tc::umma_commit(tc::smem_u32(&bar_accum));
        }
    } else {
        const int q = warp & 3;                                    // TMEM lane quadrant this warp may read
        tc::mbar_wait(tc::smem_u32(&bar_accum), 0);
        tc::tc_fence_after();
        const int row = mt * 128 + q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            float v[16];
            if (nkb > 0) tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            else {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = 0.f;
            }
            ep(row, n0 + c0, v, (int)blockIdx.z);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, BN);
}

template <int BN>
constexpr size_t gemm_smem_bytes() { return (size_t)kStages * (128 * 64 * 2 + BN * 64 * 2) + 1024; }

template <int BN, int MODE, class EP>
static cudaError_t launch_tc_gemm(const CUtensorMap &ma, const CUtensorMap &mb, const GemmParams &g, dim3 grid, EP ep, cudaStream_t st) {
    static bool configured = false;
    auto kern = tc_gemm_kernel<BN, MODE, EP>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<BN>());
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, kThreads, gemm_smem_bytes<BN>(), st>>>(ma, mb, g, ep);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ epilogues
__device__ __forceinline__ void store_bf16x16(bf16 *dst, const float (&v)[16]) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t *>(&h);
    }
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void load_bf16x16(const bf16 *src, float (&v)[16]) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 a = s[0], b = s[1];
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}

struct EpiConv1 {               // rows on the 21-grid -> Z1 [B*441][32] = relu(conv + b), zeros at invalid positions
    bf16 *z1; const float *bias; int rows;
    __device__ void operator()(int row, int col, float (&v)[16], int) const {
        if (row >= rows) return;
        int p = row % kP1;
        bool ok = (p / kG1) < 20 && (p % kG1) < 20;
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = ok ? fmaxf(v[i] + __ldg(bias + col + i), 0.f) : 0.f;
        store_bf16x16(z1 + (size_t)row * kC1 + col, v);
    }
};
struct EpiGrid7 {               // conv2: rows on the 7-grid -> A2 [B*49][64], zeros at invalid positions
    bf16 *out; const float *bias; int rows;
    __device__ void operator()(int row, int col, float (&v)[16], int) const {
        if (row >= rows) return;
        int p = row % kP2;
        bool ok = (p / kG2) < 5 && (p % kG2) < 5;
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = ok ? fmaxf(v[i] + __ldg(bias + col + i), 0.f) : 0.f;
        store_bf16x16(out + (size_t)row * 64 + col, v);
    }
};
struct EpiConv3 {               // rows on the 7-grid -> dense A3 [B][25][64] (TF flatten order h,w,c)
    bf16 *a3; const float *bias; int rows;
    __device__ void operator()(int row, int col, float (&v)[16], int) const {
        if (row >= rows) return;
        int b = row / kP2, p = row - b * kP2, oh = p / kG2, ow = p - oh * kG2;
        if (oh >= 5 || ow >= 5) return;
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i] + __ldg(bias + col + i), 0.f);
        store_bf16x16(a3 + ((size_t)b * 25 + oh * 5 + ow) * 64 + col, v);
    }
};
struct EpiFc1 {                 // H1 fp32 [B][H] = relu(fc + b)
    float *h1; const float *bias; int B, hidden;
    __device__ void operator()(int row, int col, float (&v)[16], int) const {
        if (row >= B) return;
        float4 *d = reinterpret_cast<float4 *>(h1 + (size_t)row * hidden + col);
#pragma unroll
        for (int i = 0; i < 4; i++)
            d[i] = make_float4(fmaxf(v[4 * i] + __ldg(bias + col + 4 * i), 0.f), fmaxf(v[4 * i + 1] + __ldg(bias + col + 4 * i + 1), 0.f),
                               fmaxf(v[4 * i + 2] + __ldg(bias + col + 4 * i + 2), 0.f), fmaxf(v[4 * i + 3] + __ldg(bias + col + 4 * i + 3), 0.f));
    }
};
struct EpiFc1Dgrad {            // dA3 [B][1600] masked by relu(a3) -> dZ3 on the 7-grid [B*49][64]
    bf16 *dz3; const bf16 *a3; int B;
    __device__ void operator()(int row, int col, float (&v)[16], int) const {
        if (row >= B) return;
        float a[16];
        load_bf16x16(a3 + (size_t)row * kFlat + col, a);
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = a[i] > 0.f ? v[i] : 0.f;
        int pix = col >> 6, oh = pix / 5, ow = pix - oh * 5;
        store_bf16x16(dz3 + ((size_t)row * kP2 + oh * kG2 + ow) * 64 + (col & 63), v);
    }
};
struct EpiConv3Dgrad {          // dA2 on the 7-grid masked by relu(a2) -> dZ2 [B*49][64], zeros at invalid positions
    bf16 *dz2; const bf16 *a2; int rows;
    __device__ void operator()(int row, int col, float (&v)[16], int) const {
        if (row >= rows) return;
        int p = row % kP2;
        bool ok = (p / kG2) < 5 && (p % kG2) < 5;
        float a[16];
        load_bf16x16(a2 + (size_t)row * 64 + col, a);
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = (ok && a[i] > 0.f) ? v[i] : 0.f;
        store_bf16x16(dz2 + (size_t)row * 64 + col, v);
    }
};
struct EpiStoreBf16 {           // conv2 dgrad: dP2 [B*49][128]
    bf16 *out; int rows, ld;
    __device__ void operator()(int row, int col, float (&v)[16], int) const {
        if (row >= rows) return;
        store_bf16x16(out + (size_t)row * ld + col, v);
    }
};
struct EpiStoreF32 {            // weight-gradient partials [split][rows][ld] and the self-test
    float *out; int rows, ld; size_t split_stride;
    __device__ void operator()(int row, int col, float (&v)[16], int split) const {
        if (row >= rows) return;
        float4 *d = reinterpret_cast<float4 *>(out + (size_t)split * split_stride + (size_t)row * ld + col);
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
};

// ------------------------------------------------------------------------------------------------ glue kernels
// u8 frames (FrameView) -> X2 [B*441][64]: block (bh,bw) of the input padded by 2, channel j = r*16 + s*4 + c
__global__ void pack_x2_kernel(FrameView fv, int B, bf16 *x2) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * kP1 * 8) return;
    int t8 = (int)(t & 7);
    size_t blk = t >> 3;
    int b = (int)(blk / kP1), pb = (int)(blk - (size_t)b * kP1), bh = pb / kG1, bw = pb - bh * kG1;
    int r = t8 >> 1, sp = t8 & 1;
    int ih = 4 * bh - 2 + r, iw = 4 * bw - 2 + 2 * sp;
    float px[2][4];
    bool in = (unsigned)ih < 80u && (unsigned)iw < 80u;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        unsigned short two = 0;
        if (in) two = __ldg(reinterpret_cast<const unsigned short *>(fv.base + (size_t)b * fv.sample_stride + fv.chan_off[c] + ih * 80 + iw));
        px[0][c] = (float)(two & 255);
        px[1][c] = (float)(two >> 8);
    }
    uint32_t w[4];
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            __nv_bfloat162 v = __floats2bfloat162_rn(px[s][2 * h], px[s][2 * h + 1]);
            w[s * 2 + h] = *reinterpret_cast<uint32_t *>(&v);
        }
    *reinterpret_cast<uint4 *>(x2 + blk * 64 + r * 16 + sp * 8) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
    uint4 r;
    __nv_bfloat162 *pa = reinterpret_cast<__nv_bfloat162 *>(&a), *pb = reinterpret_cast<__nv_bfloat162 *>(&b), *pr = reinterpret_cast<__nv_bfloat162 *>(&r);
#pragma unroll
    for (int i = 0; i < 4; i++) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}

// max_pool 2x2 (BrainDQN.py:128) + space-to-depth for conv2: Z1 (21-grid) -> P2 [B*49][128], channel j = r*64 + s*32 + c
__global__ void pool_pack_kernel(const bf16 *z1, int B, bf16 *p2) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * 36 * 16) return;
    int cg = (int)(t & 3), rs = (int)((t >> 2) & 3);
    size_t blk = t >> 4;
    int b = (int)(blk / 36), q = (int)(blk - (size_t)b * 36), bh = q / 6, bw = q - bh * 6;
    int r = rs >> 1, s = rs & 1;
    int ph = 2 * bh - 1 + r, pw = 2 * bw - 1 + s;
    uint4 m = make_uint4(0, 0, 0, 0);
    if ((unsigned)ph < 10u && (unsigned)pw < 10u) {
        const uint4 *src = reinterpret_cast<const uint4 *>(z1 + ((size_t)b * kP1 + (2 * ph) * kG1 + 2 * pw) * kC1 + cg * 8);
        m = bf16x8_max(bf16x8_max(__ldg(src), __ldg(src + 4)), bf16x8_max(__ldg(src + kG1 * 4), __ldg(src + kG1 * 4 + 4)));
    }
    *reinterpret_cast<uint4 *>(p2 + ((size_t)b * kP2 + bh * kG2 + bw) * 128 + r * 64 + s * 32 + cg * 8) = m;
}

// dZ1 = unpool(dP2) * relu'(z1): the gradient goes to the first maximum of each window (TF MaxPoolGrad)
__global__ void unpool_relu_kernel_tc(const bf16 *z1, const bf16 *dp2, int B, bf16 *dz1) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * 100 * 4) return;
    int cg = (int)(t & 3);
    size_t pp = t >> 2;
    int b = (int)(pp / 100), q = (int)(pp - (size_t)b * 100), ph = q / 10, pw = q - ph * 10;
    int bh = (ph + 1) >> 1, r = (ph + 1) & 1, bw = (pw + 1) >> 1, s = (pw + 1) & 1;
    float g[8], z[4][8];
    {
        uint4 raw = __ldg(reinterpret_cast<const uint4 *>(dp2 + ((size_t)b * kP2 + bh * kG2 + bw) * 128 + r * 64 + s * 32 + cg * 8));
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
        for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); g[2 * i] = f.x; g[2 * i + 1] = f.y; }
    }
    size_t base = ((size_t)b * kP1 + (2 * ph) * kG1 + 2 * pw) * kC1 + cg * 8;
    const size_t off[4] = {0, (size_t)kC1, (size_t)kG1 * kC1, (size_t)kG1 * kC1 + kC1};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint4 raw = __ldg(reinterpret_cast<const uint4 *>(z1 + base + off[k]));
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
        for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); z[k][2 * i] = f.x; z[k][2 * i + 1] = f.y; }
    }
    float o[4][8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float m = fmaxf(fmaxf(z[0][i], z[1][i]), fmaxf(z[2][i], z[3][i]));
        bool done = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            bool hit = !done && z[k][i] == m;
            o[k][i] = (hit && z[k][i] > 0.f) ? g[i] : 0.f;
            done |= hit;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { __nv_bfloat162 h = __floats2bfloat162_rn(o[k][2 * i], o[k][2 * i + 1]); w[i] = *reinterpret_cast<uint32_t *>(&h); }
        *reinterpret_cast<uint4 *>(dz1 + base + off[k]) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// fp32 parameters -> bf16 operand matrices in the K orders the GEMMs use (see the tables in make_plan)
struct PackedWeights {
    bf16 *w1p;      // [32][256]   n, (tap, r, s, c)            conv1 forward  Bt
    bf16 *w2p;      // [64][512]   n, (tap, r, s, c)            conv2 forward  Bt
    bf16 *w3p;      // [64][576]   n, (kh, kw, c)               conv3 forward  Bt
    bf16 *wf1p;     // [H][1600]   n, k                         fc1 forward    Bt
    bf16 *wf1n;     // [1600][H]   k, n   (as stored)           fc1 dgrad      Bt
    bf16 *w3d;      // [64][576]   c, (kh, kw, o)               conv3 dgrad    Bt
    bf16 *w2d;      // [128][256]  (r, s, c), (tap, o)          conv2 dgrad    Bt
};
__global__ void pack_weights_kernel(const float *params, QnetLayout L, PackedWeights pw, int fwd_only) {
    const int H = L.hidden;
    const int n1 = kC1 * kK1, n2 = kC2 * kK2, n3 = kC3 * kK3, nf = kFlat * H;
    const int total = fwd_only ? n1 + n2 + n3 + nf : n1 + n2 + n3 + nf + nf + n3 + 128 * 256;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int e = i;
        if (e < n1) {
            int n = e >> 8, k = e & 255, t = k >> 6, j = k & 63, r = j >> 4, s = (j >> 2) & 3, c = j & 3;
            int kh = 4 * (t >> 1) + r, kw = 4 * (t & 1) + s;
            pw.w1p[e] = __float2bfloat16(params[L.w1 + ((kh * 8 + kw) * 4 + c) * kC1 + n]);
            continue;
        }
        e -= n1;
        if (e < n2) {
            int n = e >> 9, k = e & 511, t = k >> 7, j = k & 127, r = j >> 6, s = (j >> 5) & 1, c = j & 31;
            int kh = 2 * (t >> 1) + r, kw = 2 * (t & 1) + s;
            pw.w2p[e] = __float2bfloat16(params[L.w2 + ((kh * 4 + kw) * 32 + c) * kC2 + n]);
            continue;
        }
        e -= n2;
        if (e < n3) {
            int n = e / kK3, k = e - n * kK3;
            pw.w3p[e] = __float2bfloat16(params[L.w3 + k * kC3 + n]);
            continue;
        }
        e -= n3;
        if (e < nf) {
            int n = e / kFlat, k = e - n * kFlat;
            pw.wf1p[e] = __float2bfloat16(params[L.wf1 + (size_t)k * H + n]);
            continue;
        }
        e -= nf;
        if (e < nf) { pw.wf1n[e] = __float2bfloat16(params[L.wf1 + e]); continue; }
        e -= nf;
        if (e < n3) {
            int c = e / kK3, k = e - c * kK3, t = k >> 6, o = k & 63;
            pw.w3d[e] = __float2bfloat16(params[L.w3 + (t * 64 + c) * kC3 + o]);
            continue;
        }
        e -= n3;
        {
            int nn = e >> 8, k = e & 255, t = k >> 6, o = k & 63, r = nn >> 6, s = (nn >> 5) & 1, c = nn & 31;
            int kh = 2 * (t >> 1) + r, kw = 2 * (t & 1) + s;
            pw.w2d[e] = __float2bfloat16(params[L.w2 + ((kh * 4 + kw) * 32 + c) * kC2 + o]);
        }
    }
}

// column sums of a bf16 matrix [rows][N] in chunks -> part[chunk][N] (bias gradients, summed in order later)
__global__ void colsum_kernel(const bf16 *x, int rows, int N, int chunk_rows, float *part) {
    __shared__ float red[256];
    const int chunk = blockIdx.x, r0 = chunk * chunk_rows, r1 = min(rows, r0 + chunk_rows);
    for (int c0 = 0; c0 < N; c0 += 256) {
        const int ncol = min(N - c0, 256), lanes = 256 / ncol;          // N in {32, 64, 512}: lanes in {8, 4, 1}
        const int col = threadIdx.x % ncol, rl = threadIdx.x / ncol;
        float s = 0.f;
        if (rl < lanes)
            for (int r = r0 + rl; r < r1; r += lanes) s += __bfloat162float(x[(size_t)r * N + c0 + col]);
        red[threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.x < ncol) {
            float tot = 0.f;
            for (int k = 0; k < lanes; k++) tot += red[k * ncol + threadIdx.x];
            part[(size_t)chunk * N + c0 + threadIdx.x] = tot;
        }
        __syncthreads();
    }
}

// split-K partials + bias partials -> the flat gradient vector in TF variable order (HWIO)
struct FinalizeArgs {
    const float *part1, *part2, *part3, *partf;     // [splits][rows][N]
    int s1, s2, s3;                                  // number of splits (fc1 has one)
    const float *bp1, *bp2, *bp3, *bpf;             // bias partials [chunks][N]
    int c1, c2, c3, cf;
};
__global__ void finalize_grads_kernel(FinalizeArgs a, QnetLayout L, float *grads) {
    const int H = L.hidden;
    const int end = L.bf1 + H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
        float s = 0.f;
        if (i < L.b1) {                     // W1 [kh][kw][c][n]
            int e = i - L.w1, n = e & 31, c = (e >> 5) & 3, kw = (e >> 7) & 7, kh = e >> 10;
            int row = ((kh >> 2) * 2 + (kw >> 2)) * 64 + (kh & 3) * 16 + (kw & 3) * 4 + c;
            for (int z = 0; z < a.s1; z++) s += a.part1[((size_t)z * 256 + row) * 32 + n];
        } else if (i < L.w2) {
            int n = i - L.b1;
            for (int k = 0; k < a.c1; k++) s += a.bp1[k * 32 + n];
        } else if (i < L.b2) {              // W2 [kh][kw][c][n]
            int e = i - L.w2, n = e & 63, c = (e >> 6) & 31, kw = (e >> 11) & 3, kh = e >> 13;
            int row = ((kh >> 1) * 2 + (kw >> 1)) * 128 + (kh & 1) * 64 + (kw & 1) * 32 + c;
            for (int z = 0; z < a.s2; z++) s += a.part2[((size_t)z * 512 + row) * 64 + n];
        } else if (i < L.w3) {
            int n = i - L.b2;
            for (int k = 0; k < a.c2; k++) s += a.bp2[k * 64 + n];
        } else if (i < L.b3) {              // W3 [k][n], k natural
            int e = i - L.w3;
            for (int z = 0; z < a.s3; z++) s += a.part3[(size_t)z * 640 * 64 + e];
        } else if (i < L.wf1) {
            int n = i - L.b3;
            for (int k = 0; k < a.c3; k++) s += a.bp3[k * 64 + n];
        } else if (i < L.bf1) {
            s = a.partf[i - L.wf1];
        } else {
            int n = i - L.bf1;
            for (int k = 0; k < a.cf; k++) s += a.bpf[k * H + n];
        }
        grads[i] = s;
    }
}

// ------------------------------------------------------------------------------------------------ TMA maps
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int ensure_encode() {
    if (g_encode) return FB_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FB_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    FB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
    g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    return FB_OK;
}

// bf16 matrix [rows][cols] row-major, box [box_rows][box_cols]; swizzle 128B (64B when the box is 32 columns wide)
int make_map(CUtensorMap *m, const bf16 *base, long long rows, long long cols, int box_cols, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    FB_REQUIRE(box_cols * 2 == 128 || box_cols * 2 == 64, "make_map: box must be 64 or 32 bf16 wide");
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16 *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fb_set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"); return FB_ERR_CUDA; }
    return FB_OK;
}

void set_kmajor(GemmParams &g) { g.lbo_a = 16; g.sbo_a = 1024; g.kstep_a = 32; g.lbo_b = 16; g.sbo_b = 1024; g.kstep_b = 32; }
void set_mnmajor(GemmParams &g, int bn) {
    g.lbo_a = 8192; g.sbo_a = 1024; g.kstep_a = 2048;
    if (bn == 32) { g.lbo_b = 4096; g.sbo_b = 512; g.kstep_b = 1024; }
    else { g.lbo_b = 8192; g.sbo_b = 1024; g.kstep_b = 2048; }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ state
struct TcWeightMaps { CUtensorMap w1p, w2p, w3p, wf1p, wf1n, w3d, w2d; };
struct TcPlan {                 // everything that depends on the batch size
    int B;
    CUtensorMap x2_k, p2_k, a2_k, a3_k, dh1_k, dz3_k, dz2_k;        // K-major A operands, box 64 x 128
    CUtensorMap x2_m, p2_m, a2_m, a3_m;                              // MN-major A operands, box 64 x 64
    CUtensorMap dz1_b, dz2_b, dz3_b, dh1_b;                          // MN-major B operands
    GemmParams conv1, conv2, conv3, fc1, fc1_d, conv3_d, conv2_d;
    GemmParams conv1_w, conv2_w, conv3_w, fc1_w;
    int s1, s2, s3;                                                  // weight-gradient splits
};
struct TcState {
    // activations / gradients (bf16 unless noted), sized for max_batch
    bf16 *x2, *z1, *p2, *a2, *a3, *dh1, *dz3, *dz2, *dp2, *dz1;
    float *part1, *part2, *part3, *partf;
    size_t cap1, cap2, cap3;
    float *bp1, *bp2, *bp3, *bpf;
    PackedWeights pw[2];
    TcWeightMaps wm[2];
    std::map<int, TcPlan> plans;
};

namespace {

constexpr int kChunk1 = 768, kChunk23 = 256, kChunkF = 64;

int plan_splits(long long P, int tiles, int *klen_out) {
    int target = 148 / tiles; if (target < 1) target = 1;
    long long klen = (P + target - 1) / target;
    klen = (klen + 63) / 64 * 64;
    if (klen < 64) klen = 64;
    *klen_out = (int)klen;
    return (int)((P + klen - 1) / klen);
}

int make_plan(fb_qnet *n, int B, TcPlan **out) {
    TcState *t = n->tc;
    auto it = t->plans.find(B);
    if (it != t->plans.end()) { *out = &it->second; return FB_OK; }
    if (t->plans.size() > 16) t->plans.clear();
    TcPlan p{};
    p.B = B;
    const int H = n->L.hidden;
    const long long P1 = (long long)B * kP1, P2 = (long long)B * kP2;
    int rc;
#define MAP(dst, base, rows, cols, bc, br) if ((rc = make_map(&p.dst, base, rows, cols, bc, br))) return rc
    MAP(x2_k, t->x2, P1, 64, 64, 128); MAP(p2_k, t->p2, P2, 128, 64, 128); MAP(a2_k, t->a2, P2, 64, 64, 128);
    MAP(a3_k, t->a3, B, kFlat, 64, 128); MAP(dh1_k, t->dh1, B, H, 64, 128); MAP(dz3_k, t->dz3, P2, 64, 64, 128);
    MAP(dz2_k, t->dz2, P2, 64, 64, 128);
    MAP(x2_m, t->x2, P1, 64, 64, 64); MAP(p2_m, t->p2, P2, 128, 64, 64); MAP(a2_m, t->a2, P2, 64, 64, 64);
    MAP(a3_m, t->a3, B, kFlat, 64, 64);
    MAP(dz1_b, t->dz1, P1, 32, 32, 64); MAP(dz2_b, t->dz2, P2, 64, 64, 64); MAP(dz3_b, t->dz3, P2, 64, 64, 64);
    MAP(dh1_b, t->dh1, B, H, 64, 64);
#undef MAP
    // ---- forward
    set_kmajor(p.conv1); p.conv1.nkb = 4;
    for (int k = 0; k < 4; k++) { p.conv1.a_rowoff[k] = (k >> 1) * kG1 + (k & 1); p.conv1.a_col[k] = 0; }
    set_kmajor(p.conv2); p.conv2.nkb = 8;
    for (int k = 0; k < 8; k++) { int tp = k >> 1; p.conv2.a_rowoff[k] = (tp >> 1) * kG2 + (tp & 1); p.conv2.a_col[k] = (k & 1) * 64; }
    set_kmajor(p.conv3); p.conv3.nkb = 9;
    for (int k = 0; k < 9; k++) { p.conv3.a_rowoff[k] = (k / 3) * kG2 + (k % 3) - 8; p.conv3.a_col[k] = 0; }
    set_kmajor(p.fc1); p.fc1.nkb = 25;
    for (int k = 0; k < 25; k++) { p.fc1.a_rowoff[k] = 0; p.fc1.a_col[k] = k * 64; }
    // ---- data gradients
    set_kmajor(p.fc1_d); p.fc1_d.nkb = H / 64;
    FB_REQUIRE(H / 64 <= kMaxKb, "tensor-core path: hidden must be <= 2048");
    for (int k = 0; k < H / 64; k++) { p.fc1_d.a_rowoff[k] = 0; p.fc1_d.a_col[k] = k * 64; }
    set_kmajor(p.conv3_d); p.conv3_d.nkb = 9;
    for (int k = 0; k < 9; k++) { p.conv3_d.a_rowoff[k] = (1 - k / 3) * kG2 + (1 - k % 3); p.conv3_d.a_col[k] = 0; }
    set_kmajor(p.conv2_d); p.conv2_d.nkb = 4;
    for (int k = 0; k < 4; k++) { p.conv2_d.a_rowoff[k] = -(k >> 1) * kG2 - (k & 1); p.conv2_d.a_col[k] = 0; }
    // ---- weight gradients (contract over positions)
    set_mnmajor(p.conv1_w, 32); p.conv1_w.p_total = (int)P1;
    p.s1 = plan_splits(P1, 2, &p.conv1_w.klen);
    for (int mt = 0; mt < 2; mt++) for (int i = 0; i < 2; i++) { int tp = 2 * mt + i; p.conv1_w.a2_rowoff[mt][i] = (tp >> 1) * kG1 + (tp & 1); p.conv1_w.a2_col[mt][i] = 0; }
    set_mnmajor(p.conv2_w, 64); p.conv2_w.p_total = (int)P2;
    p.s2 = plan_splits(P2, 4, &p.conv2_w.klen);
    for (int mt = 0; mt < 4; mt++) for (int i = 0; i < 2; i++) { p.conv2_w.a2_rowoff[mt][i] = (mt >> 1) * kG2 + (mt & 1); p.conv2_w.a2_col[mt][i] = i * 64; }
    set_mnmajor(p.conv3_w, 64); p.conv3_w.p_total = (int)P2;
    p.s3 = plan_splits(P2, 5, &p.conv3_w.klen);
    for (int mt = 0; mt < 5; mt++) for (int i = 0; i < 2; i++) { int tp = min(2 * mt + i, 8); p.conv3_w.a2_rowoff[mt][i] = (tp / 3) * kG2 + (tp % 3) - 8; p.conv3_w.a2_col[mt][i] = 0; }
    set_mnmajor(p.fc1_w, 128); p.fc1_w.p_total = B; p.fc1_w.klen = (B + 63) / 64 * 64;
    for (int mt = 0; mt < 13; mt++) for (int i = 0; i < 2; i++) { p.fc1_w.a2_rowoff[mt][i] = 0; p.fc1_w.a2_col[mt][i] = mt * 128 + i * 64; }
    FB_REQUIRE((size_t)p.s1 <= t->cap1 && (size_t)p.s2 <= t->cap2 && (size_t)p.s3 <= t->cap3, "tensor-core path: split-K workspace too small");
    auto res = t->plans.emplace(B, p);
    *out = &res.first->second;
    return FB_OK;
}

}  // namespace

int tc_state_create(fb_qnet *n) {
    if (n->tc) return FB_OK;
    FB_REQUIRE(n->L.hidden % 128 == 0 && n->L.hidden <= 2048, "tensor-core path needs hidden % 128 == 0 and hidden <= 2048");
    int rc = ensure_encode();
    if (rc) return rc;
    TcState *t = new (std::nothrow) TcState();
    FB_REQUIRE(t != nullptr, "tc_state_create: out of host memory");
    const size_t B = (size_t)n->max_batch, H = (size_t)n->L.hidden;
    auto alloc_bf = [](bf16 **p, size_t elems) {
        cudaError_t e = cudaMalloc(p, elems * sizeof(bf16));
        if (e == cudaSuccess) e = cudaMemset(*p, 0, elems * sizeof(bf16));
        return e;
    };
    auto alloc_f = [](float **p, size_t elems) {
        cudaError_t e = cudaMalloc(p, elems * sizeof(float));
        if (e == cudaSuccess) e = cudaMemset(*p, 0, elems * sizeof(float));
        return e;
    };
    FB_CUDA_OK(alloc_bf(&t->x2, B * kP1 * 64)); FB_CUDA_OK(alloc_bf(&t->z1, B * kP1 * 32)); FB_CUDA_OK(alloc_bf(&t->p2, B * kP2 * 128));
    FB_CUDA_OK(alloc_bf(&t->a2, B * kP2 * 64)); FB_CUDA_OK(alloc_bf(&t->a3, B * kFlat)); FB_CUDA_OK(alloc_bf(&t->dh1, B * H));
    FB_CUDA_OK(alloc_bf(&t->dz3, B * kP2 * 64)); FB_CUDA_OK(alloc_bf(&t->dz2, B * kP2 * 64)); FB_CUDA_OK(alloc_bf(&t->dp2, B * kP2 * 128));
    FB_CUDA_OK(alloc_bf(&t->dz1, B * kP1 * 32));
    int klen;
    t->cap1 = (size_t)plan_splits((long long)B * kP1, 2, &klen) + 1;
    t->cap2 = (size_t)plan_splits((long long)B * kP2, 4, &klen) + 1;
    t->cap3 = (size_t)plan_splits((long long)B * kP2, 5, &klen) + 1;
    // smaller batches never need more splits than max_batch does, except through rounding: keep generous caps
    if (t->cap1 < 80) t->cap1 = 80;
    if (t->cap2 < 40) t->cap2 = 40;
    if (t->cap3 < 32) t->cap3 = 32;
    FB_CUDA_OK(alloc_f(&t->part1, t->cap1 * 256 * 32)); FB_CUDA_OK(alloc_f(&t->part2, t->cap2 * 512 * 64));
    FB_CUDA_OK(alloc_f(&t->part3, t->cap3 * 640 * 64)); FB_CUDA_OK(alloc_f(&t->partf, (size_t)1664 * H));
    FB_CUDA_OK(alloc_f(&t->bp1, ((B * kP1 + kChunk1 - 1) / kChunk1) * 32)); FB_CUDA_OK(alloc_f(&t->bp2, ((B * kP2 + kChunk23 - 1) / kChunk23) * 64));
    FB_CUDA_OK(alloc_f(&t->bp3, ((B * kP2 + kChunk23 - 1) / kChunk23) * 64)); FB_CUDA_OK(alloc_f(&t->bpf, ((B + kChunkF - 1) / kChunkF) * H));
    for (int s = 0; s < 2; s++) {
        PackedWeights &w = t->pw[s];
        FB_CUDA_OK(alloc_bf(&w.w1p, kC1 * kK1)); FB_CUDA_OK(alloc_bf(&w.w2p, kC2 * kK2)); FB_CUDA_OK(alloc_bf(&w.w3p, kC3 * kK3));
        FB_CUDA_OK(alloc_bf(&w.wf1p, kFlat * H)); FB_CUDA_OK(alloc_bf(&w.wf1n, kFlat * H)); FB_CUDA_OK(alloc_bf(&w.w3d, kC3 * kK3));
        FB_CUDA_OK(alloc_bf(&w.w2d, 128 * 256));
        TcWeightMaps &m = t->wm[s];
        if ((rc = make_map(&m.w1p, w.w1p, kC1, kK1, 64, 32))) return rc;
        if ((rc = make_map(&m.w2p, w.w2p, kC2, kK2, 64, 64))) return rc;
        if ((rc = make_map(&m.w3p, w.w3p, kC3, kK3, 64, 64))) return rc;
        if ((rc = make_map(&m.wf1p, w.wf1p, (long long)H, kFlat, 64, 128))) return rc;
        if ((rc = make_map(&m.wf1n, w.wf1n, kFlat, (long long)H, 64, 64))) return rc;
        if ((rc = make_map(&m.w3d, w.w3d, kC3, kK3, 64, 64))) return rc;
        if ((rc = make_map(&m.w2d, w.w2d, 128, 256, 64, 128))) return rc;
    }
    n->tc = t;
    return FB_OK;
}

void tc_state_destroy(fb_qnet *n) {
    TcState *t = n->tc;
    if (!t) return;
    void *ps[] = {t->x2, t->z1, t->p2, t->a2, t->a3, t->dh1, t->dz3, t->dz2, t->dp2, t->dz1, t->part1, t->part2, t->part3, t->partf,
                  t->bp1, t->bp2, t->bp3, t->bpf};
    for (void *p : ps) cudaFree(p);
    for (int s = 0; s < 2; s++) {
        PackedWeights &w = t->pw[s];
        void *ws[] = {w.w1p, w.w2p, w.w3p, w.wf1p, w.wf1n, w.w3d, w.w2d};
        for (void *p : ws) cudaFree(p);
    }
    delete t;
    n->tc = nullptr;
}

int tc_pack_weights(fb_qnet *n, const float *params_dev, int slot, cudaStream_t st) {
    FB_REQUIRE(n->tc != nullptr && (slot == 0 || slot == 1), "tc_pack_weights: bad argument");
    pack_weights_kernel<<<592, 256, 0, st>>>(params_dev, n->L, n->tc->pw[slot], slot == 1 ? 1 : 0);
    FB_CUDA_OK(cudaGetLastError());
    return FB_OK;
}

int tc_forward(fb_qnet *n, int slot, const float *params_dev, FrameView fv, int B, float *q_out, cudaStream_t st) {
    TcState *t = n->tc;
    FB_REQUIRE(t != nullptr && B > 0 && B <= n->max_batch, "tc_forward: bad argument");
    TcPlan *p;
    int rc = make_plan(n, B, &p);
    if (rc) return rc;
    const QnetLayout &L = n->L;
    const TcWeightMaps &wm = t->wm[slot];
    const int P1 = B * kP1, P2 = B * kP2;
    pack_x2_kernel<<<(unsigned)(((size_t)P1 * 8 + 255) / 256), 256, 0, st>>>(fv, B, t->x2);
    FB_CUDA_OK((launch_tc_gemm<32, 0>(p->x2_k, wm.w1p, p->conv1, dim3((P1 + 127) / 128, 1, 1), EpiConv1{t->z1, params_dev + L.b1, P1}, st)));
    pool_pack_kernel<<<(unsigned)(((size_t)B * 36 * 16 + 255) / 256), 256, 0, st>>>(t->z1, B, t->p2);
    FB_CUDA_OK((launch_tc_gemm<64, 0>(p->p2_k, wm.w2p, p->conv2, dim3((P2 + 127) / 128, 1, 1), EpiGrid7{t->a2, params_dev + L.b2, P2}, st)));
    FB_CUDA_OK((launch_tc_gemm<64, 0>(p->a2_k, wm.w3p, p->conv3, dim3((P2 + 127) / 128, 1, 1), EpiConv3{t->a3, params_dev + L.b3, P2}, st)));
    FB_CUDA_OK((launch_tc_gemm<128, 0>(p->a3_k, wm.wf1p, p->fc1, dim3((B + 127) / 128, L.hidden / 128, 1),
                                       EpiFc1{n->h1, params_dev + L.bf1, B, L.hidden}, st)));
    qnet_launch_head_forward(n->h1, params_dev, L, B, q_out, st);
    FB_CUDA_OK(cudaGetLastError());
    return FB_OK;
}

// backward of the LAST tc_forward (online net on s); n->dq holds dLoss/dQ.  Fills grads_dev completely.
int tc_backward(fb_qnet *n, const float *params_dev, int B, float *grads_dev, cudaStream_t st) {
    TcState *t = n->tc;
    FB_REQUIRE(t != nullptr && B > 0 && B <= n->max_batch, "tc_backward: bad argument");
    TcPlan *p;
    int rc = make_plan(n, B, &p);
    if (rc) return rc;
    const QnetLayout &L = n->L;
    const TcWeightMaps &wm = t->wm[0];
    const int H = L.hidden, P1 = B * kP1, P2 = B * kP2;
    // head: fp32 gradients of the head variables straight into grads, dh1 (masked by relu) as bf16
    qnet_launch_head_backward(n->h1, n->dq, params_dev, L, B, grads_dev, nullptr, t->dh1, st);
    // fc1: dW = a3^T dh1, dz3 = (dh1 Wf1^T) * relu'(a3)
    FB_CUDA_OK((launch_tc_gemm<128, 1>(p->a3_m, p->dh1_b, p->fc1_w, dim3(13, H / 128, 1), EpiStoreF32{t->partf, kFlat, H, 0}, st)));
    FB_CUDA_OK((launch_tc_gemm<64, 0>(p->dh1_k, wm.wf1n, p->fc1_d, dim3((B + 127) / 128, kFlat / 64, 1), EpiFc1Dgrad{t->dz3, t->a3, B}, st)));
    // conv3
    FB_CUDA_OK((launch_tc_gemm<64, 1>(p->a2_m, p->dz3_b, p->conv3_w, dim3(5, 1, p->s3), EpiStoreF32{t->part3, 640, 64, (size_t)640 * 64}, st)));
    FB_CUDA_OK((launch_tc_gemm<64, 0>(p->dz3_k, wm.w3d, p->conv3_d, dim3((P2 + 127) / 128, 1, 1), EpiConv3Dgrad{t->dz2, t->a2, P2}, st)));
    // conv2
    FB_CUDA_OK((launch_tc_gemm<64, 1>(p->p2_m, p->dz2_b, p->conv2_w, dim3(4, 1, p->s2), EpiStoreF32{t->part2, 512, 64, (size_t)512 * 64}, st)));
    FB_CUDA_OK((launch_tc_gemm<128, 0>(p->dz2_k, wm.w2d, p->conv2_d, dim3((P2 + 127) / 128, 1, 1), EpiStoreBf16{t->dp2, P2, 128}, st)));
    unpool_relu_kernel_tc<<<(unsigned)(((size_t)B * 400 + 255) / 256), 256, 0, st>>>(t->z1, t->dp2, B, t->dz1);
    // conv1 (no input gradient)
    FB_CUDA_OK((launch_tc_gemm<32, 1>(p->x2_m, p->dz1_b, p->conv1_w, dim3(2, 1, p->s1), EpiStoreF32{t->part1, 256, 32, (size_t)256 * 32}, st)));
    // bias gradients = column sums of the dZ tensors
    const int c1 = (P1 + kChunk1 - 1) / kChunk1, c23 = (P2 + kChunk23 - 1) / kChunk23, cf = (B + kChunkF - 1) / kChunkF;
    colsum_kernel<<<c1, 256, 0, st>>>(t->dz1, P1, 32, kChunk1, t->bp1);
    colsum_kernel<<<c23, 256, 0, st>>>(t->dz2, P2, 64, kChunk23, t->bp2);
    colsum_kernel<<<c23, 256, 0, st>>>(t->dz3, P2, 64, kChunk23, t->bp3);
    colsum_kernel<<<cf, 256, 0, st>>>(t->dh1, B, H, kChunkF, t->bpf);
    FinalizeArgs fa{t->part1, t->part2, t->part3, t->partf, p->s1, p->s2, p->s3, t->bp1, t->bp2, t->bp3, t->bpf, c1, c23, c23, cf};
    finalize_grads_kernel<<<592, 256, 0, st>>>(fa, L, grads_dev);
    FB_CUDA_OK(cudaGetLastError());
    return FB_OK;
}

// ------------------------------------------------------------------------------------------------ self-test hook
// Raw GEMM on caller matrices through the same kernel, so that the descriptor encodings are pinned on their own:
//   mode 0: D[M][N] = A[M][K] * Bt[N][K]^T          (both K-major)
//   mode 1: D[M][N] = A[K][M]^T * B[K][N]           (both MN-major; contraction over rows)
// bn in {32, 64, 128}; N % bn == 0; strides6 (may be NULL) overrides {lbo_a, sbo_a, kstep_a, lbo_b, sbo_b, kstep_b}.
extern "C" int fb_debug_tc_gemm(int mode, int bn, int M, int N, int K, const void *a_dev, const void *b_dev, float *d_dev,
                                const uint32_t *strides6, void *stream) {
    FB_REQUIRE(a_dev && b_dev && d_dev && (mode == 0 || mode == 1) && (bn == 32 || bn == 64 || bn == 128) && N % bn == 0,
               "fb_debug_tc_gemm: bad argument");
    int rc = ensure_encode();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap ma, mb;
    GemmParams g{};
    const bf16 *A = (const bf16 *)a_dev, *Bm = (const bf16 *)b_dev;
    dim3 grid;
    if (mode == 0) {
        FB_REQUIRE(K % 64 == 0 && K / 64 <= kMaxKb, "fb_debug_tc_gemm: K must be a multiple of 64, at most 2048");
        if ((rc = make_map(&ma, A, M, K, 64, 128))) return rc;
        if ((rc = make_map(&mb, Bm, N, K, 64, bn))) return rc;
        set_kmajor(g);
        g.nkb = K / 64;
        for (int k = 0; k < g.nkb; k++) { g.a_rowoff[k] = 0; g.a_col[k] = k * 64; }
        grid = dim3((M + 127) / 128, N / bn, 1);
    } else {
        FB_REQUIRE((M + 127) / 128 <= 16, "fb_debug_tc_gemm: M must be at most 2048 in mode 1");
        if ((rc = make_map(&ma, A, K, M, 64, 64))) return rc;
        if ((rc = make_map(&mb, Bm, K, N, bn == 32 ? 32 : 64, 64))) return rc;
        set_mnmajor(g, bn);
        g.p_total = K; g.klen = (K + 63) / 64 * 64;
        for (int mt = 0; mt < (M + 127) / 128; mt++) for (int i = 0; i < 2; i++) { g.a2_rowoff[mt][i] = 0; g.a2_col[mt][i] = mt * 128 + i * 64; }
        grid = dim3((M + 127) / 128, N / bn, 1);
    }
    if (strides6) { g.lbo_a = strides6[0]; g.sbo_a = strides6[1]; g.kstep_a = strides6[2]; g.lbo_b = strides6[3]; g.sbo_b = strides6[4]; g.kstep_b = strides6[5]; }
    EpiStoreF32 ep{d_dev, M, N, 0};
    cudaError_t e;
    if (mode == 0) e = bn == 32 ? launch_tc_gemm<32, 0>(ma, mb, g, grid, ep, st) : bn == 64 ? launch_tc_gemm<64, 0>(ma, mb, g, grid, ep, st) : launch_tc_gemm<128, 0>(ma, mb, g, grid, ep, st);
    else e = bn == 32 ? launch_tc_gemm<32, 1>(ma, mb, g, grid, ep, st) : bn == 64 ? launch_tc_gemm<64, 1>(ma, mb, g, grid, ep, st) : launch_tc_gemm<128, 1>(ma, mb, g, grid, ep, st);
    FB_CUDA_OK(e);
    return FB_OK;
}
