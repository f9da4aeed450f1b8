// Slab convolution kernels of the tensor-core Q-network path (included by fb_qnet_tc.cu).
//
// tcgen05.mma applies the 128-byte swizzle to ABSOLUTE shared-memory address bits, exactly as TMA does when it
// writes a tile (pinned on hardware by fb_debug_tc_slab / tests/test_qnet_tc_gpu.py::test_slab_descriptor_shift).
// A matrix descriptor may therefore start ANY number of 128-byte rows into a TMA-written slab.  With activations
// stored [grid position][channel] (see fb_qnet_tc.cu) all filter taps of a convolution read the SAME rows shifted by
// a constant, so one slab of 128 + halo rows is loaded once per tile and every tap is just another descriptor:
//   forward / data gradient  (K-major A):   rows = output positions, tap = row offset of the A descriptor
//   weight gradient          (MN-major A):  rows = contraction index, tap = row offset; two taps form one M=128
//                                           accumulator whose second 64-row atom is LBO bytes after the first
// Operand bytes from L2 per tile drop 4x (conv1) to 8x (conv3) against one TMA box per tap.
#pragma once
#include "fb_tc.cuh"

namespace {

constexpr int kConvThreads = 192;          // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..5 epilogue

struct ConvParams {
    int n_tiles;                           // tiles of 128 output positions
    int slab_row0;                         // the slab of tile t starts at row 128 t + slab_row0 (<= 0)
    int kb_rowoff[16];                     // K-block kb reads slab rows [kb_rowoff, kb_rowoff + 128) ...
    int kb_half[16];                       // ... of column half kb_half (64 channels each)
};

// D[p][n] = sum_kb sum_c A[p + tap(kb)][64 half(kb) + c] * Bt[n][64 kb + c], persistent over tiles of 128 positions.
// The whole Bt (NKB x BN x 64) stays resident in shared memory; slabs stream through S stages; two TMEM accumulators
// let the epilogue of tile i overlap the MMAs of tile i+1.
template <int BN, int SLAB_ROWS, int NHALF, int NKB, int S, int T, class EP>
__global__ void __launch_bounds__(kConvThreads, 1) tc_conv_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                   const __grid_constant__ CUtensorMap mapB, const ConvParams g, const EP ep) {
    constexpr uint32_t HALF_BYTES = SLAB_ROWS * 128, STAGE = NHALF * HALF_BYTES, BBLK = BN * 128, B_BYTES = NKB * BBLK;
    static_assert(SLAB_ROWS % 8 == 0 && SLAB_ROWS <= 256, "slab rows: multiple of 8 (1024-byte stage alignment), one TMA box");
    static_assert(S % T == 0 && 2 * T * BN <= 512, "a group of T tiles occupies T consecutive stages and T accumulators of a set");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[S], bar_empty[S], bar_b, bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float bias_s[BN];
    const uint32_t smem_b = (tc::smem_u32(smem_raw) + 1023u) & ~1023u, smem_a = smem_b + B_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t TMEM_COLS = 2 * T * BN <= 32 ? 32 : 2 * T * BN <= 64 ? 64 : 2 * T * BN <= 128 ? 128 : 2 * T * BN <= 256 ? 256 : 512;

    if (threadIdx.x == 0) {
        tc::tma_prefetch_desc(&mapA);
        tc::tma_prefetch_desc(&mapB);
        for (int s = 0; s < S; s++) { tc::mbar_init(tc::smem_u32(&bar_full[s]), 1); tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1); }
        tc::mbar_init(tc::smem_u32(&bar_b), 1);
        for (int a = 0; a < 2; a++) { tc::mbar_init(tc::smem_u32(&bar_acc_full[a]), 1); tc::mbar_init(tc::smem_u32(&bar_acc_empty[a]), 4); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_slot), TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (threadIdx.x == 0) {                          // the weights do not depend on the previous kernel: fetch them first
        tc::mbar_expect_tx(tc::smem_u32(&bar_b), B_BYTES);
        for (int kb = 0; kb < NKB; kb++) tc::tma_load_2d(smem_b + kb * BBLK, &mapB, kb * 64, 0, tc::smem_u32(&bar_b));
    }
    if (ep.bias != nullptr && threadIdx.x >= 64 && (int)threadIdx.x - 64 < BN) bias_s[threadIdx.x - 64] = ep.bias[threadIdx.x - 64];
    tc::pdl_wait();
    tc::pdl_launch();
    __syncthreads();

    if (warp == 0) {
        const bool leader = tc::elect_one();
        int i = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, i++) {
            const int s = i % S;
            tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ((i / S) & 1) ^ 1u);
            if (leader) {
                const uint32_t full = tc::smem_u32(&bar_full[s]);
                tc::mbar_expect_tx(full, STAGE);
#pragma unroll
                for (int h = 0; h < NHALF; h++)
                    tc::tma_load_2d(smem_a + s * STAGE + h * HALF_BYTES, &mapA, h * 64, tile * 128 + g.slab_row0, full);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        {
            const bool leader = tc::elect_one();
            constexpr uint32_t idesc = tc::instr_desc_bf16(128, BN, 0, 0);
            constexpr uint32_t dhi = tc::smem_desc_hi(1024, tc::kSwizzle128);
            // descriptor start fields, in 16-byte units: per K-block for A (tap row offset, column half), constant for B
            uint32_t a_rel[NKB];
#pragma unroll
            for (int kb = 0; kb < NKB; kb++) a_rel[kb] = (uint32_t)(g.kb_half[kb] * (int)HALF_BYTES + g.kb_rowoff[kb] * 128) >> 4;
            const uint32_t a_lo0 = tc::smem_desc_lo(smem_a, 16), b_lo0 = tc::smem_desc_lo(smem_b, 16);
            tc::mbar_wait(tc::smem_u32(&bar_b), 0);
            // Tiles are taken T at a time, their MMAs issued round-robin over T accumulators.  Measured on the B200: T > 1
            // does NOT help (conv1 forward 45 -> 50 us at 2048 samples) -- the per-MMA cost at N = 32 is the 4 KB A-operand
            // read from shared memory (sm__pipe_tc_cycles_active 54 % vs tensor math 14 %), not an accumulator dependency --
            // so every launch uses T = 1; the grouping is kept for wider-N layers.
            int grp = 0;
            for (int tile0 = blockIdx.x; tile0 < g.n_tiles; tile0 += T * gridDim.x, grp++) {
                const int as = grp & 1;
                int nt = 0;
#pragma unroll
                for (int t = 0; t < T; t++) nt += (tile0 + t * (int)gridDim.x < g.n_tiles) ? 1 : 0;
                tc::mbar_wait(tc::smem_u32(&bar_acc_empty[as]), ((grp >> 1) & 1) ^ 1u);
                for (int t = 0; t < nt; t++) {
                    const int it = grp * T + t;
                    tc::mbar_wait(tc::smem_u32(&bar_full[it % S]), (it / S) & 1);
                }
                tc::tc_fence_after();
                const uint32_t s0 = (uint32_t)((grp * T) % S);
                if (leader) {
#pragma unroll
                    for (int kb = 0; kb < NKB; kb++)
#pragma unroll
                        for (int k = 0; k < 4; k++)
#pragma unroll
                            for (int t = 0; t < T; t++)
                                if (t < nt)
                                    tc::umma_bf16_lohi(tmem + (as * T + t) * BN, a_lo0 + (s0 + t) * (STAGE >> 4) + a_rel[kb] + 2 * k, dhi,
                                                       b_lo0 + ((kb * BBLK + k * 32) >> 4), dhi, idesc, (kb | k) != 0);
                    for (int t = 0; t < nt; t++) tc::umma_commit(tc::smem_u32(&bar_empty[(s0 + t) % S]));
                    tc::umma_commit(tc::smem_u32(&bar_acc_full[as]));
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        int grp = 0;
        for (int tile0 = blockIdx.x; tile0 < g.n_tiles; tile0 += T * gridDim.x, grp++) {
            const int as = grp & 1;
            tc::mbar_wait(tc::smem_u32(&bar_acc_full[as]), (grp >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll 1
            for (int t = 0; t < T; t++) {
                const int tile = tile0 + t * (int)gridDim.x;
                const bool last = t == T - 1 || tile + (int)gridDim.x >= g.n_tiles;
                if (tile < g.n_tiles) {
                    const int row = tile * 128 + q * 32 + lane;
#pragma unroll 1
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        float v[32];
                        tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * T + t) * BN + c0), v);
                        if (last && c0 + 32 >= BN) {         // accumulator set drained: hand it back before the stores
                            tc::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_acc_empty[as]));
                        }
                        ep(row, c0, *reinterpret_cast<float(*)[16]>(&v[0]), 0, bias_s);
                        ep(row, c0 + 16, *reinterpret_cast<float(*)[16]>(&v[16]), 0, bias_s);
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, TMEM_COLS);
}

template <int BN, int SLAB_ROWS, int NHALF, int NKB, int S>
constexpr size_t conv_smem_bytes() { return (size_t)NKB * BN * 128 + (size_t)S * NHALF * SLAB_ROWS * 128 + 1024; }

template <int BN, int SLAB_ROWS, int NHALF, int NKB, int S, int T, class EP>
static cudaError_t launch_tc_conv(const CUtensorMap &ma, const CUtensorMap &mb, const ConvParams &g, int max_ctas, EP ep, cudaStream_t st) {
    static bool configured = false;
    auto kern = tc_conv_kernel<BN, SLAB_ROWS, NHALF, NKB, S, T, EP>;
    constexpr size_t smem = conv_smem_bytes<BN, SLAB_ROWS, NHALF, NKB, S>();
    static_assert(smem <= 227 * 1024, "shared memory budget");
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    int grid = g.n_tiles < max_ctas ? g.n_tiles : max_ctas;
    return tc::launch_pdl(kern, dim3(grid), dim3(kConvThreads), smem, st, ma, mb, g, ep);
}

// ---- weight gradient: D[128 a + 64 i + j][n] = sum_p A[p + off(a, i)][j] * B[p][n] over this CTA's rows p ---------
struct WgradParams {
    int p_total, klen;                     // CTA z contracts rows [z klen, min(p_total, (z+1) klen)), klen % 64 == 0
    int slab_row0;                         // the A slab of K-block p starts at row p + slab_row0
    int acc_rowoff[8];                     // accumulator a: first atom starts acc_rowoff rows into the slab (column half 0),
    uint32_t acc_lbo[8];                   // its second 64-row atom acc_lbo bytes later
};

template <int BN, int NACC, int SLAB_ROWS, int NHALF, int S, class EP>
__global__ void __launch_bounds__(kConvThreads, 1) tc_wgrad_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                    const __grid_constant__ CUtensorMap mapB, const WgradParams g, const EP ep) {
    constexpr uint32_t HALF_BYTES = SLAB_ROWS * 128, A_BYTES = NHALF * HALF_BYTES, B_BYTES = 64 * BN * 2, STAGE = A_BYTES + B_BYTES;
    constexpr uint32_t B_LAYOUT = BN == 32 ? tc::kSwizzle64 : tc::kSwizzle128;
    constexpr uint32_t B_SBO = BN == 32 ? 512 : 1024, B_KSTEP = BN == 32 ? 1024 : 2048;
    constexpr uint32_t TMEM_COLS = NACC * BN <= 32 ? 32 : NACC * BN <= 64 ? 64 : NACC * BN <= 128 ? 128 : NACC * BN <= 256 ? 256 : 512;
    static_assert(SLAB_ROWS % 8 == 0 && NACC * BN <= 512 && (BN == 32 || BN == 64), "wgrad tile shape");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[S], bar_empty[S], bar_acc;
    __shared__ uint32_t tmem_slot;
    const uint32_t smem = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p_begin = blockIdx.x * g.klen;
    const int nkb = (min(g.klen, g.p_total - p_begin) + 63) >> 6;

    if (threadIdx.x == 0) {
        tc::tma_prefetch_desc(&mapA);
        tc::tma_prefetch_desc(&mapB);
        for (int s = 0; s < S; s++) { tc::mbar_init(tc::smem_u32(&bar_full[s]), 1); tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1); }
        tc::mbar_init(tc::smem_u32(&bar_acc), 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_slot), TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    tc::pdl_wait();
    tc::pdl_launch();

    if (warp == 0) {
        const bool leader = tc::elect_one();
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % S;
            tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ((kb / S) & 1) ^ 1u);
            if (leader) {
                const uint32_t full = tc::smem_u32(&bar_full[s]), sa = smem + s * STAGE;
                const int p = p_begin + kb * 64;
                tc::mbar_expect_tx(full, STAGE);
#pragma unroll
                for (int h = 0; h < NHALF; h++) tc::tma_load_2d(sa + h * HALF_BYTES, &mapA, h * 64, p + g.slab_row0, full);
                tc::tma_load_2d(sa + A_BYTES, &mapB, 0, p, full);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        {
            const bool leader = tc::elect_one();
            constexpr uint32_t idesc = tc::instr_desc_bf16(128, BN, 1, 1);
            constexpr uint32_t ahi = tc::smem_desc_hi(1024, tc::kSwizzle128), bhi = tc::smem_desc_hi(B_SBO, B_LAYOUT);
            uint32_t a_lo0[NACC];                 // start (16-byte units) + LBO of each accumulator's A descriptor, stage 0
#pragma unroll
            for (int a = 0; a < NACC; a++)
                a_lo0[a] = tc::smem_desc_lo(smem + g.acc_rowoff[a] * 128, NHALF == 2 ? HALF_BYTES : g.acc_lbo[a]);
            const uint32_t b_lo0 = tc::smem_desc_lo(smem + A_BYTES, 8192);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % S;
                tc::mbar_wait(tc::smem_u32(&bar_full[s]), (kb / S) & 1);
                tc::tc_fence_after();
                const uint32_t soff = s * (STAGE >> 4);
                if (leader) {
                    // k outer, accumulators inner (neutral on the B200: the MMA rate is set by operand reads, see tc_conv_kernel)
#pragma unroll
                    for (int k = 0; k < 4; k++)
#pragma unroll
                        for (int a = 0; a < NACC; a++)
                            tc::umma_bf16_lohi(tmem + a * BN, a_lo0[a] + soff + k * 128, ahi, b_lo0 + soff + k * (B_KSTEP >> 4), bhi, idesc,
                                               (kb | k) != 0);
                    tc::umma_commit(tc::smem_u32(&bar_empty[s]));
                }
                __syncwarp();
            }
            if (leader) tc::umma_commit(tc::smem_u32(&bar_acc));
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        tc::mbar_wait(tc::smem_u32(&bar_acc), 0);
        tc::tc_fence_after();
#pragma unroll 1
        for (int a = 0; a < NACC; a++)
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float v[16];
                if (nkb > 0) tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c0), v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = 0.f;
                }
                ep(a * 128 + q * 32 + lane, c0, v, (int)blockIdx.x, nullptr);
            }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, TMEM_COLS);
}

template <int BN, int NACC, int SLAB_ROWS, int NHALF, int S, class EP>
static cudaError_t launch_tc_wgrad(const CUtensorMap &ma, const CUtensorMap &mb, const WgradParams &g, int splits, EP ep, cudaStream_t st) {
    static bool configured = false;
    auto kern = tc_wgrad_kernel<BN, NACC, SLAB_ROWS, NHALF, S, EP>;
    constexpr size_t smem = (size_t)S * (NHALF * SLAB_ROWS * 128 + 64 * BN * 2) + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    return tc::launch_pdl(kern, dim3(splits), dim3(kConvThreads), smem, st, ma, mb, g, ep);
}

}  // namespace
