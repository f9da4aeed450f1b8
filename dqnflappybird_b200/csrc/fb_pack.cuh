// bf16 operand copies of the parameters (tensor-core path): their layouts and the one place that knows where parameter i
// goes.  Shared by the pack kernel, the Adam kernels of fb_qnet_tc.cu and the exchange + Adam kernel of fb_dist.cu, so that
// whoever updates a parameter also refreshes its operand copies and no pack kernel runs in steady state.
#pragma once
#include "fb_qnet.cuh"
#include "fb_tc.cuh"

using bf16 = __nv_bfloat16;

// fp32 parameters -> bf16 operand matrices in the K orders the GEMMs use (see the tables in make_plan, fb_qnet_tc.cu)
struct PackedWeights {
    bf16 *w1p;      // [32][256]   n, (tap, r, s, c)            conv1 forward  Bt
    bf16 *w2p;      // [64][512]   n, (tap, r, s, c)            conv2 forward  Bt
    bf16 *w3p;      // [64][576]   n, (kh, kw, c)               conv3 forward  Bt
    bf16 *wf1n;     // [1600][H]   k, n   (as stored)           fc1 forward B (MN-major) and fc1 dgrad Bt
    bf16 *w3d;      // [64][576]   c, (kh, kw, o)               conv3 dgrad    Bt
    bf16 *w2d;      // [128][256]  (r, s, c), (tap, o)          conv2 dgrad    Bt
    int f16;        // operand format of the copies: 0 bf16, 1 fp16 (the pointers are 16-bit storage either way)
};
// parameter i (TF variable order) -> its bf16 operand copies
__device__ __forceinline__ void scatter_packed(int i, float val, const QnetLayout &L, const PackedWeights &pw, int fwd_only) {
    const unsigned short bits = tc::pack1(val, pw.f16);
    const bf16 v = *reinterpret_cast<const bf16 *>(&bits);
    if (i < L.b1) {                         // W1 [kh][kw][c][n]
        int e = i - L.w1, n = e & 31, c = (e >> 5) & 3, kw = (e >> 7) & 7, kh = e >> 10;
        int k = ((kh >> 2) * 2 + (kw >> 2)) * 64 + (kh & 3) * 16 + (kw & 3) * 4 + c;
        pw.w1p[n * kK1 + k] = v;
    } else if (i >= L.w2 && i < L.b2) {     // W2 [kh][kw][c][o]
        int e = i - L.w2, o = e & 63, c = (e >> 6) & 31, kw = (e >> 11) & 3, kh = e >> 13;
        int t = (kh >> 1) * 2 + (kw >> 1), j = (kh & 1) * 64 + (kw & 1) * 32 + c;
        pw.w2p[o * kK2 + t * 128 + j] = v;
        if (!fwd_only) pw.w2d[j * 256 + t * 64 + o] = v;
    } else if (i >= L.w3 && i < L.b3) {     // W3 [t][c][o]
        int e = i - L.w3, o = e & 63, c = (e >> 6) & 63, t = e >> 12;
        pw.w3p[o * kK3 + t * 64 + c] = v;
        if (!fwd_only) pw.w3d[c * kK3 + t * 64 + o] = v;
    } else if (i >= L.wf1 && i < L.bf1) {   // W_fc1 [k][n]: same index
        pw.wf1n[i - L.wf1] = v;
    }
}

// the online net's operand copies (slot 0) if the net runs the tensor-core path; false otherwise (fb_qnet_tc.cu)
bool tc_online_operands(fb_qnet *n, PackedWeights *out);
