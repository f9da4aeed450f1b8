// Q-network geometry shared by the fp32 and tensor-core paths.
//
// Reference graph: BrainDQN.py:119-163 (identical in every Brain; dueling head
// BrainDuelingDQN_CC.py:68-77).  TensorFlow NHWC activations, HWIO weights, SAME padding:
//   x [B,80,80,4] -> conv 8x8 s4 (pad 2) +b relu [B,20,20,32] -> maxpool 2x2 s2 [B,10,10,32]
//     -> conv 4x4 s2 (pad 1) +b relu [B,5,5,64] -> conv 3x3 s1 (pad 1) +b relu [B,5,5,64]
//     -> flatten (h,w,c) 1600 -> fc H (512) relu -> fc A (2)   [dueling: V (1) and A (2) heads]
// Here x is never materialised: channel c of sample b is the 80x80 u8 frame at
// frames + b * sample_stride + chan_off[c] (a view of the frame ring or of a gathered batch),
// H = obs axis 0 (game x), W = obs axis 1 (game y), exactly the array the reference feeds.
#pragma once
#include "fb_common.cuh"

constexpr int kC1 = 32, kC2 = 64, kC3 = 64, kFlat = 1600, kActions = 2;
constexpr int kK1 = 8 * 8 * 4, kK2 = 4 * 4 * 32, kK3 = 3 * 3 * 64;

struct QnetLayout {             // offsets (in floats) into the flat parameter vector, TF variable creation order
    int hidden, dueling;
    int w1, b1, w2, b2, w3, b3, wf1, bf1;
    int wf2, bf2;               // plain head [H,2], [2]
    int wv, bv, wa, ba;         // dueling heads [H,1],[1],[H,2],[2]
    int total;
};

inline QnetLayout qnet_layout(int hidden, int dueling) {
    QnetLayout L{};
    L.hidden = hidden; L.dueling = dueling;
    int o = 0;
    L.w1 = o; o += kK1 * kC1; L.b1 = o; o += kC1;
    L.w2 = o; o += kK2 * kC2; L.b2 = o; o += kC2;
    L.w3 = o; o += kK3 * kC3; L.b3 = o; o += kC3;
    L.wf1 = o; o += kFlat * hidden; L.bf1 = o; o += hidden;
    if (!dueling) { L.wf2 = o; o += hidden * kActions; L.bf2 = o; o += kActions; L.wv = L.bv = L.wa = L.ba = -1; }
    else { L.wv = o; o += hidden; L.bv = o; o += 1; L.wa = o; o += hidden * kActions; L.ba = o; o += kActions; L.wf2 = L.bf2 = -1; }
    L.total = o;
    return L;
}

struct FrameView {              // where the 4 input channels of each sample live
    const uint8_t *base;
    long long sample_stride;    // bytes between consecutive samples
    int chan_off[4];            // byte offset of channel c (oldest frame first, newest last: BrainDQN.py:68)
};
