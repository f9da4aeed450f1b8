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

// Every tcgen05 kernel asks for more than half of an SM's shared memory, so that two tensor-core CTAs (of the same or of
// different, concurrently running kernels) never share an SM.  The block scheduler knows nothing about TMEM: a CTA placed
// beside one that holds the columns it needs spins in tcgen05.alloc with its shared memory and registers taken -- measured in
// the training step's backward pass, where three weight-gradient kernels and two data-gradient kernels overlap: the conv2
// weight gradient took 24-29 us for 2 us of tensor work.  With exclusive SMs the scheduler queues CTAs instead.
constexpr size_t kExclusiveSmem = 116 * 1024;
constexpr size_t exclusive_smem(size_t need) { return need > kExclusiveSmem ? need : kExclusiveSmem; }

struct ConvParams {
    int n_tiles;                           // tiles of 128 output positions
    int slab_row0;                         // the slab of tile t starts at row 128 t + slab_row0 (<= 0)
    int kb_rowoff[16];                     // K-block kb reads slab rows [kb_rowoff, kb_rowoff + 128) ...
    int kb_half[16];                       // ... of column half kb_half (64 channels each)
    int f16;                               // operand format (tc::pack2): 0 bf16, 1 fp16
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
    __shared__ __align__(16) float bias_s[BN];
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
            const uint32_t idesc = tc::instr_desc_16(128, BN, 0, 0, g.f16);
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
    constexpr size_t smem = exclusive_smem(conv_smem_bytes<BN, SLAB_ROWS, NHALF, NKB, S>());
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
    int f16;
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
            const uint32_t idesc = tc::instr_desc_16(128, BN, 1, 1, g.f16);
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
        // The partial sums go out as whole 128-byte lines: a lane holds 32 columns of ITS row, the warp transposes them through
        // shared memory (the operand stages are free by now; 16-byte pieces XOR-swizzled by row, conflict-free both ways) so that
        // eight lanes store one row's 128 bytes.  Straight from the registers every store instruction touched 32 rows with 16
        // bytes each -- 17 MB of half-sector writes per update from the three weight-gradient kernels, at the update's very end.
        // (EP must have EpiStoreF32's fields: out, rows, ld, split_stride, scale.)
        float4 *stage = reinterpret_cast<float4 *>(smem_raw + (smem - tc::smem_u32(smem_raw))) + q * 256;      // 4 KB per warp
        float *outp = ep.out + (size_t)blockIdx.x * ep.split_stride;
#pragma unroll 1
        for (int a = 0; a < NACC; a++)
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                if (nkb > 0) tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c0), v);
                else {
#pragma unroll
                    for (int i = 0; i < 32; i++) v[i] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; j++)
                    stage[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j] * ep.scale, v[4 * j + 1] * ep.scale, v[4 * j + 2] * ep.scale, v[4 * j + 3] * ep.scale);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int r = (lane >> 3) + 4 * j, c4 = lane & 7, row = a * 128 + q * 32 + r;
                    const float4 x = stage[r * 8 + (c4 ^ (r & 7))];
                    if (row < ep.rows) *reinterpret_cast<float4 *>(outp + (size_t)row * ep.ld + c0 + 4 * c4) = x;
                }
                __syncwarp();
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
    constexpr size_t smem = exclusive_smem((size_t)S * (NHALF * SLAB_ROWS * 128 + 64 * BN * 2) + 1024);
    static_assert(smem <= 227 * 1024, "shared memory budget");
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    return tc::launch_pdl(kern, dim3(splits), dim3(kConvThreads), smem, st, ma, mb, g, ep);
}

}  // namespace

// ---- conv1 fused with its neighbours ---------------------------------------------------------------------------------
// frames (u8, FrameView) -> [space-to-depth slab built in shared memory] -> conv 8x8 s4 + bias + ReLU -> Z1 (optional,
// kept for backward) -> 2x2 max-pool -> P2 in conv2's space-to-depth layout.  Neither the bf16 input matrix X2
// (56 KB/sample) nor, when acting, Z1 (28 KB/sample) ever exists in HBM: 25.6 KB of u8 in, 12.5 KB of P2 out per sample.
//
// Tile = 6 rows of one sample's 21-wide block grid (126 positions, M = 128), so that every 2x2 pooling window is inside
// one tile; 4 tiles per sample.  Warps: 0 weights (TMA), 1 MMA issue, 2-9 epilogue (+ pooling through shared memory; two groups
// on alternate tiles, staging buffer i & 3 so that a group's next tile never overwrites rows still being pooled),
// 10-17 slab builders, one tile per group in flight (u8 -> bf16, written with the 128-byte swizzle TMA would have used, fence.proxy.async, mbarrier).
namespace {

constexpr int kBuilders = 256;                                   // 8 slab-builder warps in two groups
constexpr int kRoleThreads1 = 64 + 256;                          // TMA warp, MMA warp, two epilogue groups of 4 warps
constexpr int kFusedThreads = kRoleThreads1 + kBuilders;
constexpr int kTileRows1 = 6, kTilePos1 = kTileRows1 * 21;      // 126 positions per tile
constexpr int kSlabF = 152;                                     // 126 + 22 halo, rounded up to 8

// rows [row0, row0 + nrows) of the virtual matrix X2 [B*441][64] (block (bh,bw) of the input padded by 2, channel
// j = r*16 + s*4 + c) into a swizzled slab: 16-byte chunk c16 of slab row r lands at r*128 + ((c16 ^ (r & 7)) << 4)
// Slab building, two steps, by the 256 builder threads of the fused conv1 kernel:
//   (A) the tile's raw pixels -- 4 frames x 32 pixel rows x 80 bytes, this sample only; 32 rows of a frame are one
//       contiguous 2,560-byte run -- arrive as four cp.async.bulk copies per tile into a ring of raw stages, issued
//       several tiles ahead by warp 0 (one tile in flight at a time is bound by the HBM latency, ~3,400 cycles a tile);
//   (B) every thread gathers 2 pixels x 4 frames per 16-byte chunk from there, converts u8 -> bf16 and stores the chunk
//       where TMA's 128-byte swizzle would have put it (row r, chunk c16 -> r*128 + ((c16 ^ (r & 7)) << 4)).
// Gathering straight from global memory with 2-byte loads asks L1 for 7x the sectors the tile holds and was the bound.
constexpr int kRawRows = 32, kRawBytes = 4 * kRawRows * 80;     // raw pixels per tile (10,240 bytes)
constexpr int kRawStages = 4;

// 1,216 chunks of 16 bytes per slab over the builder threads (every warp equally loaded); r / 21 by multiply-shift;
// loads unconditional (clamped address) and zeroing by a select, so that no divergent block surrounds the LDS
template <int NT>
__device__ __forceinline__ void slab_from_raw(uint8_t *slab, const uint8_t *raw, int tq, int tid, int f16) {
    constexpr int K = (kSlabF * 8 + NT - 1) / NT, LAST = kSlabF * 8 - 1;
    // all of a thread's loads first (one warp per scheduler: the conversions of chunk k would otherwise wait out the
    // shared-memory latency of chunk k's own loads, ~350 cycles a chunk), then convert and store
    uint32_t t[K][4];
    uint32_t dst[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        int item = tid + k * NT;
        item = item < LAST ? item : LAST;
        const int r = item >> 3, c16 = item & 7;
        const int bh_l = (r * 3121) >> 16, bw = r - bh_l * 21;              // r / 21, r % 21 for r < 400
        const int rr = 4 * bh_l + (c16 >> 1), ih = 24 * tq - 2 + rr, iw = 4 * bw - 2 + 2 * (c16 & 1);
        const bool col_ok = (unsigned)iw < 80u;
        const bool valid = col_ok && kTileRows1 * tq + bh_l < 21 && (unsigned)ih < 80u;
        const uint8_t *src = raw + rr * 80 + (col_ok ? iw : 0);
#pragma unroll
        for (int c = 0; c < 4; c++) t[k][c] = *reinterpret_cast<const unsigned short *>(src + c * (kRawRows * 80));
        dst[k] = (uint32_t)(r * 128 + ((c16 ^ (r & 7)) << 4)) | (valid ? 0x80000000u : 0u);
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        const bool valid = (dst[k] & 0x80000000u) != 0;
        uint32_t w[4];
#pragma unroll
        for (int s2 = 0; s2 < 2; s2++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float lo = (float)((t[k][2 * h] >> (8 * s2)) & 255), hi = (float)((t[k][2 * h + 1] >> (8 * s2)) & 255);
                w[s2 * 2 + h] = valid ? tc::pack2(lo, hi, f16) : 0u;
            }
        if (tid + k * NT <= LAST) *reinterpret_cast<uint4 *>(slab + (dst[k] & 0x7fffffffu)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

struct Conv1FusedParams {
    FrameView fv;
    int B;
    const float *bias;
    __nv_bfloat16 *z1;          // [B*441][32] or nullptr (acting / target forward: not needed)
    __nv_bfloat16 *p2;          // [B*49][128]
    int f16;
};

// FROM_X2: the slab comes by TMA from a materialised X2 (pack_x2_kernel) instead of the in-kernel builders -- conv1 with
// only the pooling fused (no Z1 round trip, no pool kernel); 192 threads.
template <int S, bool FROM_X2>
__global__ void __launch_bounds__(FROM_X2 ? kRoleThreads1 : kFusedThreads, 1) tc_conv1_fused_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                                          const __grid_constant__ CUtensorMap mapB,
                                                                                          const Conv1FusedParams g) {
    constexpr int BN = 32, NKB = 4;
    constexpr uint32_t STAGE = kSlabF * 128, BBLK = BN * 128, B_BYTES = NKB * BBLK, ZS_BYTES = 128 * BN * 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[S], bar_empty[S], bar_b, bar_acc_full[2], bar_acc_empty[2], bar_raw_full[kRawStages],
        bar_raw_empty[kRawStages];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float bias_s[BN];
    uint8_t *smem_gen = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-aligned, generic pointer
    const uint32_t smem_b = tc::smem_u32(smem_gen), smem_a = smem_b + B_BYTES;
    uint8_t *slab_gen = smem_gen + B_BYTES, *zs_gen = slab_gen + S * STAGE, *raw_gen = zs_gen + 4 * ZS_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = 4 * g.B;
    const long long total_rows = (long long)g.B * 441;

    if (threadIdx.x == 0) {
        tc::tma_prefetch_desc(&mapB);
        for (int s = 0; s < S; s++) { tc::mbar_init(tc::smem_u32(&bar_full[s]), 1); tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1); }
        tc::mbar_init(tc::smem_u32(&bar_b), 1);
        for (int a = 0; a < 2; a++) { tc::mbar_init(tc::smem_u32(&bar_acc_full[a]), 1); tc::mbar_init(tc::smem_u32(&bar_acc_empty[a]), 4); }
        for (int k = 0; k < kRawStages; k++) { tc::mbar_init(tc::smem_u32(&bar_raw_full[k]), 1); tc::mbar_init(tc::smem_u32(&bar_raw_empty[k]), 1); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_slot), 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(tc::smem_u32(&bar_b), B_BYTES);
        for (int kb = 0; kb < NKB; kb++) tc::tma_load_2d(smem_b + kb * BBLK, &mapB, kb * 64, 0, tc::smem_u32(&bar_b));
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + BN) bias_s[threadIdx.x - 64] = g.bias[threadIdx.x - 64];
    tc::pdl_wait();
    tc::pdl_launch();
    __syncthreads();

    if (warp == 0 && FROM_X2) {
        // ===== slab loader: one TMA box of 152 rows of X2 per tile
        const bool leader = tc::elect_one();
        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, i++) {
            const int s = i % S;
            tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ((i / S) & 1) ^ 1u);
            if (leader) {
                const uint32_t full = tc::smem_u32(&bar_full[s]);
                tc::mbar_expect_tx(full, STAGE);
                tc::tma_load_2d(smem_a + s * STAGE, &mapA, 0, (tile >> 2) * 441 + (tile & 3) * kTilePos1, full);
            }
            __syncwarp();
        }
    } else if (warp == 0) {
        // ===== raw pixel loader: four bulk copies per tile (one contiguous run of rows per frame), kRawStages tiles ahead
        const bool leader = tc::elect_one();
        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, i++) {
            const int rs = i % kRawStages, b = tile >> 2, tq = tile & 3;
            tc::mbar_wait(tc::smem_u32(&bar_raw_empty[rs]), ((i / kRawStages) & 1) ^ 1u);
            if (leader) {
                const int ih0 = 24 * tq - 2, lo = ih0 < 0 ? 0 : ih0, hi = ih0 + kRawRows > 80 ? 80 : ih0 + kRawRows;       // valid rows [lo, hi)
                const uint32_t bytes = (uint32_t)(hi - lo) * 80u, full = tc::smem_u32(&bar_raw_full[rs]);
                const uint32_t dst = tc::smem_u32(raw_gen) + rs * kRawBytes + (uint32_t)(lo - ih0) * 80u;
                tc::mbar_expect_tx(full, 4 * bytes);
#pragma unroll
                for (int c = 0; c < 4; c++) tc::bulk_load_1d(dst + c * (kRawRows * 80), g.fv.chan(b, c) + lo * 80, bytes, full);
            }
            __syncwarp();
        }
    } else if (warp >= kRoleThreads1 / 32) {
        // ===== slab builders: two groups of kBuilders / 2 threads take alternate tiles (named barriers 2 and 3), so that one
        // group's fence / barrier / hand-over latency overlaps the other's conversion
        constexpr int GT = kBuilders / 2;
        const int grp = (threadIdx.x - kRoleThreads1) / GT, tid = (threadIdx.x - kRoleThreads1) - grp * GT;
        int i = grp;
        for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < n_tiles; tile += 2 * (int)gridDim.x, i += 2) {
            const int s = i % S, rs = i % kRawStages;
            tc::mbar_wait(tc::smem_u32(&bar_raw_full[rs]), (i / kRawStages) & 1);
            tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ((i / S) & 1) ^ 1u);
            slab_from_raw<GT>(slab_gen + s * STAGE, raw_gen + rs * kRawBytes, tile & 3, tid, g.f16);
            tc::fence_proxy_async();                      // generic-proxy writes -> visible to the tensor core's async proxy
            if (grp == 0) asm volatile("bar.sync 2, %0;" ::"n"(GT) : "memory");
            else asm volatile("bar.sync 3, %0;" ::"n"(GT) : "memory");
            if (tid == 0) { tc::mbar_arrive(tc::smem_u32(&bar_full[s])); tc::mbar_arrive(tc::smem_u32(&bar_raw_empty[rs])); }
        }
    } else if (warp == 1) {
        // ===== MMA issue
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::instr_desc_16(128, BN, 0, 0, g.f16);
        constexpr uint32_t dhi = tc::smem_desc_hi(1024, tc::kSwizzle128);
        const uint32_t a_lo0 = tc::smem_desc_lo(smem_a, 16), b_lo0 = tc::smem_desc_lo(smem_b, 16);
        tc::mbar_wait(tc::smem_u32(&bar_b), 0);
        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, i++) {
            const int s = i % S, as = i & 1;
            tc::mbar_wait(tc::smem_u32(&bar_acc_empty[as]), ((i >> 1) & 1) ^ 1u);
            tc::mbar_wait(tc::smem_u32(&bar_full[s]), (i / S) & 1);
            tc::tc_fence_after();
            if (leader) {
                const uint32_t a_lo = a_lo0 + s * (STAGE >> 4);
#pragma unroll
                for (int kb = 0; kb < NKB; kb++) {
                    const uint32_t tap = (uint32_t)(((kb >> 1) * 21 + (kb & 1)) * 128) >> 4;       // taps at rows 0, 1, 21, 22
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        tc::umma_bf16_lohi(tmem + as * BN, a_lo + tap + 2 * k, dhi, b_lo0 + ((kb * BBLK + k * 32) >> 4), dhi, idesc, (kb | k) != 0);
                }
                tc::umma_commit(tc::smem_u32(&bar_empty[s]));
                tc::umma_commit(tc::smem_u32(&bar_acc_full[as]));
            }
            __syncwarp();
        }
    } else if (warp >= 2 && warp < kRoleThreads1 / 32) {
        // ===== epilogue: bias + ReLU, Z1, max-pool through shared memory, P2.  Two groups of four warps on alternate tiles
        // (group e owns accumulator e and staging buffer e): at N = 32 a tile's MMAs take ~640 cycles, one group's
        // epilogue about twice that
        const int eg = (warp - 2) >> 2;
        const int q = warp & 3, r = q * 32 + lane;          // accumulator row = position r of the tile (TMEM lanes of warp % 4)
        const int et = threadIdx.x - 64 - 128 * eg;         // 0..127 within the group
        int i = eg;
        for (int tile = blockIdx.x + eg * (int)gridDim.x; tile < n_tiles; tile += 2 * (int)gridDim.x, i += 2) {
            const int as = i & 1, b = tile >> 2, tq = tile & 3;
            tc::mbar_wait(tc::smem_u32(&bar_acc_full[as]), (i >> 1) & 1);
            tc::tc_fence_after();
            float v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN), v);
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_acc_empty[as]));
            const int oh_l = r / 21, ow = r - oh_l * 21, oh = kTileRows1 * tq + oh_l;
            const bool ok = r < kTilePos1 && oh < 20 && ow < 20;
            // bias first, unconditionally and vectorised: with the loads inside the `ok ? :` the compiler emitted sixteen
            // divergent blocks per tile, each re-deriving the shared window (S2UR) before two dependent LDS (3,000 cycles a tile)
            float bia[32];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 t4 = reinterpret_cast<const float4 *>(bias_s)[k];
                bia[4 * k] = t4.x; bia[4 * k + 1] = t4.y; bia[4 * k + 2] = t4.z; bia[4 * k + 3] = t4.w;
            }
            const float keep = ok ? 1.f : 0.f;
            uint32_t w[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const float x0 = fmaxf(v[2 * k] + bia[2 * k], 0.f) * keep, x1 = fmaxf(v[2 * k + 1] + bia[2 * k + 1], 0.f) * keep;
                w[k] = tc::pack2(x0, x1, g.f16);
            }
            uint4 *zs = reinterpret_cast<uint4 *>(zs_gen + (i & 3) * ZS_BYTES + r * 64);
            if (g.p2 != nullptr)
#pragma unroll
            for (int k = 0; k < 4; k++) zs[(k + (r >> 1)) & 3] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);   // rotate: fewer bank conflicts
            if (ok && g.z1 != nullptr) {
                uint4 *d = reinterpret_cast<uint4 *>(g.z1 + ((size_t)b * 441 + oh * 21 + ow) * BN);
#pragma unroll
                for (int k = 0; k < 4; k += 2)
                    tc::st_global_256(d + k, make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]), make_uint4(w[4 * k + 4], w[4 * k + 5], w[4 * k + 6], w[4 * k + 7]));
            }
            if (g.p2 == nullptr) continue;               // (measurement only: no pooling, no output)
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 4, 128;" ::: "memory");
            if (et < 120) {
                const int ph_l = et / 40, rem = et - ph_l * 40, pw = rem >> 2, cg = rem & 3, ph = 3 * tq + ph_l;
                if (ph < 10) {
                    const uint8_t *zb = zs_gen + (i & 3) * ZS_BYTES;
                    auto ld = [&](int row) { return *reinterpret_cast<const uint4 *>(zb + row * 64 + (((cg + (row >> 1)) & 3) << 4)); };
                    const int r0 = (2 * ph_l) * 21 + 2 * pw;
                    uint4 m = ld(r0), o1 = ld(r0 + 1), o2 = ld(r0 + 21), o3 = ld(r0 + 22);
                    uint32_t *pm = reinterpret_cast<uint32_t *>(&m);
                    const uint32_t *p1 = reinterpret_cast<const uint32_t *>(&o1), *p2 = reinterpret_cast<const uint32_t *>(&o2),
                                   *p3 = reinterpret_cast<const uint32_t *>(&o3);
#pragma unroll
                    for (int k = 0; k < 4; k++) pm[k] = tc::max2(tc::max2(pm[k], p1[k], g.f16), tc::max2(p2[k], p3[k], g.f16), g.f16);
                    const int bh = (ph + 1) >> 1, rr = (ph + 1) & 1, bw = (pw + 1) >> 1, ss = (pw + 1) & 1;
                    *reinterpret_cast<uint4 *>(g.p2 + ((size_t)b * 49 + bh * 7 + bw) * 128 + rr * 64 + ss * 32 + cg * 8) = m;
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 64);
}

template <int S, bool FROM_X2>
static cudaError_t launch_tc_conv1_fused(const CUtensorMap &ma, const CUtensorMap &mb, const Conv1FusedParams &g, int max_ctas, cudaStream_t st) {
    static bool configured = false;
    auto kern = tc_conv1_fused_kernel<S, FROM_X2>;
    constexpr size_t smem = exclusive_smem(4 * 32 * 128 + (size_t)S * kSlabF * 128 + 4 * 128 * 32 * 2 + (FROM_X2 ? 0 : kRawStages * kRawBytes) + 1024);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    int grid = 4 * g.B < max_ctas ? 4 * g.B : max_ctas;
    return tc::launch_pdl(kern, dim3(grid), dim3(FROM_X2 ? kRoleThreads1 : kFusedThreads), smem, st, ma, mb, g);
}

}  // namespace
