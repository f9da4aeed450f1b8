// Multi-GPU gradient exchange fused with the optimizer (the learner is replicated, the ONE exchange step of
// the path is the sum of the per-shard gradients before tf.train.AdamOptimizer, BrainDQN.py:163).
//
// One process per GPU.  Every rank keeps its gradient vector in a cudaMalloc'd exchange buffer that the other ranks
// map through CUDA IPC (NVLink peer access).  One kernel per step, on every rank:
//   1. publish: write this step's number into every peer's flag array (system-scope release; the gradients were
//      produced by earlier kernels of the same stream);
//   2. wait until every rank's flag shows the step (polling LOCAL memory);
//   3. for each parameter: sum the gradients of ranks 0..G-1 in rank order straight from peer memory (identical
//      order on every rank => bitwise identical parameters everywhere, no broadcast ever needed) and apply TF-1 Adam.
// With four or more ranks step 3 is split (two-shot): each rank first reduces only ITS 1/G slice of the vector from all
// peers into a `red` buffer, a grid barrier + second flag publishes it, and every rank then reads the G reduced slices
// (its own locally) while applying Adam: 2 (G-1)/G vector reads per rank over NVLink instead of G-1.
// No separate all-reduce, no reduced-gradient round trip through HBM.  Exchange buffers are double-buffered by step
// parity: a rank may already write step s+1's gradients while a slow peer still reads step s's; it cannot reach s+2
// before every peer has published s+1, i.e. has finished reading s.  Waits are bounded and end in a trap, not a hang.
#include <string.h>

#include <new>

#include "fb_qnet.cuh"
#include "fb_pack.cuh"

constexpr int kMaxRanks = 8;
constexpr int kBuckets = 2;                  // 0: W_fc1 (its gradient is final early in the backward pass), 1: everything else
constexpr int kBucketFlag0 = 64;             // flags[kBucketFlag0 + 16 b + q] = newest step rank q has published for bucket b

struct fb_dist {
    int debug_mask;                          // timing experiments only (fb_dist_debug_mask): bit b: bucket b sums its own gradient alone,
                                             // bit 2 + b: bucket b skips the publish / wait handshake -- the results are then WRONG
    int rank, world;
    size_t n;
    float *xgrads[2];
    uint32_t *flags;                         // local; flags[q] = newest step rank q has published, flags[32 + q] = its reduced slice
    float *red;                              // this rank's reduced slice (two-shot)
    unsigned int *grid_count;                // grid barrier counter (two-shot)
    int two_shot;
    uint32_t ts_steps;                       // two-shot exchanges done (the grid barrier counter is monotonic)
    const float *peer_grads[2][kMaxRanks];
    uint32_t *peer_flags[kMaxRanks];
    const float *peer_red[kMaxRanks];
    void *opened[4 * kMaxRanks];
    int n_opened;
    uint32_t step;
    unsigned long long *stamps;              // debug: %globaltimer at the phase boundaries of the last exchange (block 0)
    // exchange INSIDE the training step's graph (dist_launch_bucket): the step number lives in device memory, one counter
    // per bucket, so that a captured step replays without per-step arguments
    uint32_t *dev_steps;                     // [kBuckets] exchanges completed, per bucket
    unsigned int *dev_done;                  // [kBuckets] CTAs of the running bucket kernel that have finished (self-resetting)
};

namespace {

struct XArgs {
    const float *g[kMaxRanks];
    uint32_t *peer_flags[kMaxRanks];
    const float *red[kMaxRanks];             // two-shot: every rank's reduced slice
    float *my_red;
    unsigned int *grid_count;
    const uint32_t *my_flags;
    int rank, world, wait, two_shot;
    size_t slice4;                           // float4s per slice (two-shot)
    uint32_t step, ts_step;                  // exchange number; number of two-shot exchanges including this one
    unsigned long long *stamps;              // optional phase time stamps
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer(const float *p) {          // never served from a stale L1 line
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void stamp(const XArgs &x, int k) {
    if (x.stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        x.stamps[k] = t;
    }
}
__device__ __forceinline__ float adam_x(float &p, float g, float &m, float &v, float alpha, float beta1, float beta2, float eps) {
    m += (g - m) * (1.f - beta1);            // TF 1.12 ApplyAdam functor, as adam_kernel in fb_qnet.cu
    v += (g * g - v) * (1.f - beta2);
    p -= (m * alpha) / (sqrtf(v) + eps);
    return p;
}

__global__ void __launch_bounds__(256, 2) adam_xreduce_kernel(float *__restrict__ params, float *__restrict__ am, float *__restrict__ av, size_t n,
                                                           const XArgs x, float alpha, float beta1, float beta2, float eps, float grad_scale,
                                                           float *__restrict__ reduced_out, int repack, const QnetLayout L,
                                                           const PackedWeights pw) {
    stamp(x, 0);
    if (x.wait) {
        if (blockIdx.x == 0 && (int)threadIdx.x < x.world) {
            __threadfence_system();
            st_release_sys(x.peer_flags[threadIdx.x] + x.rank, x.step);
        }
        if ((int)threadIdx.x < x.world) {
            uint32_t spins = 0;
            while ((int32_t)(ld_acquire_sys(x.my_flags + threadIdx.x) - x.step) < 0)
                if (++spins > (1u << 26)) __trap();
        }
        __syncthreads();
    }
    stamp(x, 1);
    const size_t n4 = n >> 2;
    if (x.two_shot) {
        // shot 1: this rank's slice, summed over the ranks in rank order
        const size_t lo = x.slice4 * (size_t)x.rank, hi = min(n4, lo + x.slice4);
        for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
            float4 g = ld_peer(x.g[0] + 4 * i);
            for (int q = 1; q < x.world; q++) {
                float4 h = ld_peer(x.g[q] + 4 * i);
                g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
            }
            reinterpret_cast<float4 *>(x.my_red)[i - lo] = g;
        }
        stamp(x, 2);
        // grid barrier (every CTA of this grid is resident: 2 per SM), then publish the slice and wait for everyone's
        // Measured on 8 GPUs: one thread publishing to the eight peers in a loop cost ~4 us per system-scope release store, so the
        // last peer saw the slice ~30 us after the first and the skew carried into the next step; now one thread per peer.
        // One system-scope fence per CTA, by the thread that then arrives at the grid barrier: it follows the CTA barrier, so it
        // is cumulative over the slice stores of the CTA's other threads (a fence in every thread cost ~10 us here).
        __shared__ int s_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            s_last = atomicAdd(x.grid_count, 1u) + 1u == gridDim.x * x.ts_step;  // the last CTA of this grid
            __threadfence_system();
        }
        __syncthreads();
        if (s_last && (int)threadIdx.x < x.world) st_release_sys(x.peer_flags[threadIdx.x] + 32 + x.rank, x.step);
        if ((int)threadIdx.x < x.world) {
            uint32_t spins = 0;
            while ((int32_t)(ld_acquire_sys(x.my_flags + 32 + threadIdx.x) - x.step) < 0)
                if (++spins > (1u << 26)) __trap();
        }
        __syncthreads();
    }
    stamp(x, 3);
    // every remote load of a thread is issued before the first Adam update (the vector is ~3 float4 per thread: three
    // dependent NVLink round trips otherwise)
    constexpr int U = 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += U * stride) {
        float4 gs[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = base + u * stride;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n4) {
                if (x.two_shot) {
                    const int q = (int)(i / x.slice4);
                    g = ld_peer(x.red[q] + 4 * (i - (size_t)q * x.slice4));
                } else {
                    g = ld_peer(x.g[0] + 4 * i);
                    for (int q = 1; q < x.world; q++) {
                        float4 h = ld_peer(x.g[q] + 4 * i);
                        g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
                    }
                }
            }
            gs[u] = g;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = base + u * stride;
            if (i >= n4) break;
            float4 g = gs[u];
            g.x *= grad_scale; g.y *= grad_scale; g.z *= grad_scale; g.w *= grad_scale;
            float4 p = reinterpret_cast<float4 *>(params)[i], m = reinterpret_cast<float4 *>(am)[i], v = reinterpret_cast<float4 *>(av)[i];
            adam_x(p.x, g.x, m.x, v.x, alpha, beta1, beta2, eps); adam_x(p.y, g.y, m.y, v.y, alpha, beta1, beta2, eps);
            adam_x(p.z, g.z, m.z, v.z, alpha, beta1, beta2, eps); adam_x(p.w, g.w, m.w, v.w, alpha, beta1, beta2, eps);
            reinterpret_cast<float4 *>(params)[i] = p; reinterpret_cast<float4 *>(am)[i] = m; reinterpret_cast<float4 *>(av)[i] = v;
            if (repack) {                         // tensor-core path: the bf16 operand copies of what was just updated
                const int e = (int)(4 * i);
                scatter_packed(e, p.x, L, pw, 0); scatter_packed(e + 1, p.y, L, pw, 0); scatter_packed(e + 2, p.z, L, pw, 0); scatter_packed(e + 3, p.w, L, pw, 0);
            }
            if (reduced_out) reinterpret_cast<float4 *>(reduced_out)[i] = g;
        }
    }
    stamp(x, 4);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t i = n4 * 4; i < n; i++) {                        // tail (n % 4 parameters)
            float g = 0.f;
            for (int q = 0; q < x.world; q++) g += __ldcv(x.g[q] + i);
            g *= grad_scale;
            adam_x(params[i], g, am[i], av[i], alpha, beta1, beta2, eps);
            if (repack) scatter_packed((int)i, params[i], L, pw, 0);
            if (reduced_out) reduced_out[i] = g;
        }
}

// ---- the exchange as a node of the training step's graph --------------------------------------------------------------
// One kernel per BUCKET of the parameter vector: publish "my gradients of this bucket for step s are final" into every peer
// (system-scope release), wait until every peer has published the same, then for each parameter of the bucket sum the ranks'
// gradients in rank order straight from peer memory and apply TF-1 Adam (bitwise identical replicas).  One-shot only: every
// CTA depends on the peers' flags and on nothing else of its own grid -- no grid barrier, no co-residency requirement (round 1's
// two-shot form needed all 296 CTAs resident).  Nothing about a step is a launch argument: the step number is a device
// counter advanced by the CTA that finishes last, alpha was left in device memory by the step's head kernel, so the captured
// step replays as is.  Bucket 0 (W_fc1: 91 % of the vector, final ~50 us before the step ends) runs beside the convolution
// gradients; bucket 1 (79,522 parameters, 2.2 MB over NVLink at eight ranks) is all that is left at the tail.
struct BucketArgs {
    const float *g[kMaxRanks];
    uint32_t *peer_flags[kMaxRanks];         // each rank's flag block for this bucket
    const uint32_t *my_flags;
    uint32_t *dev_step;
    unsigned int *dev_done;
    int rank, world;
    int lo4, hi4;                            // float4 range of the bucket ...
    int skip_lo4, skip4;                     // ... from which [skip_lo4, skip_lo4 + skip4) is left out (bucket 1 = everything but W_fc1)
    int tail_lo, tail_hi;                    // scalar tail (total % 4 parameters), bucket 1 only
    int sum_world, handshake;                // = world, 1 -- except under fb_dist_debug_mask (timing experiments, wrong sums)
};

__global__ void __launch_bounds__(256) adam_xbucket_kernel(float *__restrict__ params, float *__restrict__ am, float *__restrict__ av, const BucketArgs x,
                                                           const float *__restrict__ alpha_dev, float beta1, float beta2, float eps, float grad_scale,
                                                           int repack, const QnetLayout L, const PackedWeights pw) {
    const uint32_t step = *x.dev_step + 1u;
    if (x.handshake && blockIdx.x == 0 && (int)threadIdx.x < x.world) {
        __threadfence_system();
        st_release_sys(x.peer_flags[threadIdx.x] + x.rank, step);
    }
    if (x.handshake && (int)threadIdx.x < x.world) {
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(x.my_flags + threadIdx.x) - step) < 0)
            if (++spins > (1u << 26)) __trap();
    }
    __syncthreads();
    const float alpha = *alpha_dev;
    constexpr int U = 4;
    const int stride = gridDim.x * blockDim.x;
    const int n4 = x.hi4 - x.lo4 - x.skip4;                            // compact index k -> i = lo4 + k (+ skip4 past the hole)
    for (int base = blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += U * stride) {
        float4 gs[U];
#pragma unroll
        for (int u = 0; u < U; u++) {                                  // every remote load first, then the updates
            const int k = base + u * stride;
            const int i = x.lo4 + k + (x.lo4 + k >= x.skip_lo4 ? x.skip4 : 0);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < n4) {
                g = ld_peer(x.g[0] + 4 * (size_t)i);
                for (int q = 1; q < x.sum_world; q++) {
                    float4 h = ld_peer(x.g[q] + 4 * (size_t)i);
                    g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
                }
            }
            gs[u] = g;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int k = base + u * stride;
            if (k >= n4) break;
            const int i = x.lo4 + k + (x.lo4 + k >= x.skip_lo4 ? x.skip4 : 0);
            float4 g = gs[u];
            g.x *= grad_scale; g.y *= grad_scale; g.z *= grad_scale; g.w *= grad_scale;
            float4 p = reinterpret_cast<float4 *>(params)[i], m = reinterpret_cast<float4 *>(am)[i], v = reinterpret_cast<float4 *>(av)[i];
            adam_x(p.x, g.x, m.x, v.x, alpha, beta1, beta2, eps); adam_x(p.y, g.y, m.y, v.y, alpha, beta1, beta2, eps);
            adam_x(p.z, g.z, m.z, v.z, alpha, beta1, beta2, eps); adam_x(p.w, g.w, m.w, v.w, alpha, beta1, beta2, eps);
            reinterpret_cast<float4 *>(params)[i] = p; reinterpret_cast<float4 *>(am)[i] = m; reinterpret_cast<float4 *>(av)[i] = v;
            if (repack) {
                const int e = 4 * i;
                scatter_packed(e, p.x, L, pw, 0); scatter_packed(e + 1, p.y, L, pw, 0); scatter_packed(e + 2, p.z, L, pw, 0); scatter_packed(e + 3, p.w, L, pw, 0);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int i = x.tail_lo; i < x.tail_hi; i++) {
            float g = 0.f;
            for (int q = 0; q < x.world; q++) g += __ldcv(x.g[q] + i);
            g *= grad_scale;
            adam_x(params[i], g, am[i], av[i], alpha, beta1, beta2, eps);
            if (repack) scatter_packed(i, params[i], L, pw, 0);
        }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(x.dev_done, 1u) == gridDim.x - 1) { *x.dev_step = step; *x.dev_done = 0u; }   // every CTA has read the step
}

}  // namespace

// bucket 0: W_fc1 [L.wf1, L.bf1); bucket 1: [0, L.wf1) and [L.bf1, total).  Called by the tensor-core training step
// (fb_qnet_tc.cu) for the exchange buffer of the CURRENT parity; fb_dist_advance flips it once the step has been enqueued.
int dist_launch_bucket(fb_dist *d, fb_qnet *net, int bucket, float *params_dev, float *m_dev, float *v_dev, const float *alpha_dev, float beta1,
                       float beta2, float eps, float grad_scale, cudaStream_t st) {
    FB_REQUIRE(d && net && (bucket == 0 || bucket == 1) && (size_t)net->L.total == d->n, "dist_launch_bucket: bad argument");
    const QnetLayout &L = net->L;
    FB_REQUIRE(L.wf1 % 4 == 0 && L.bf1 % 4 == 0, "dist_launch_bucket: bucket boundaries must be 16-byte aligned");
    const int par = (int)(d->step & 1);
    PackedWeights pw{};
    const int repack = tc_online_operands(net, &pw) ? 1 : 0;
    auto launch = [&](int lo4, int hi4, int skip_lo4, int skip4, int tail_lo, int tail_hi, int b, int ctas) -> int {
        BucketArgs x{};
        for (int q = 0; q < d->world; q++) {
            FB_REQUIRE(d->peer_grads[par][q] != nullptr && d->peer_flags[q] != nullptr, "dist_launch_bucket: call fb_dist_connect first");
            x.g[q] = d->peer_grads[par][q]; x.peer_flags[q] = d->peer_flags[q] + kBucketFlag0 + 16 * b;
        }
        x.my_flags = d->flags + kBucketFlag0 + 16 * b; x.dev_step = d->dev_steps + b; x.dev_done = d->dev_done + b;
        x.sum_world = (d->debug_mask >> b) & 1 ? 1 : d->world; x.handshake = (d->debug_mask >> (2 + b)) & 1 ? 0 : 1;
        x.rank = d->rank; x.world = d->world; x.lo4 = lo4; x.hi4 = hi4; x.skip_lo4 = skip_lo4; x.skip4 = skip4; x.tail_lo = tail_lo; x.tail_hi = tail_hi;
        adam_xbucket_kernel<<<ctas, 256, 0, st>>>(params_dev, m_dev, v_dev, x, alpha_dev, beta1, beta2, eps, grad_scale, repack, L, pw);
        FB_CUDA_OK(cudaGetLastError());
        return FB_OK;
    };
    if (bucket == 0) return launch(L.wf1 / 4, L.bf1 / 4, L.bf1 / 4, 0, 0, 0, 0, 200);
    // bucket 1: [0, total) without W_fc1, one kernel (one handshake)
    return launch(0, L.total / 4, L.wf1 / 4, (L.bf1 - L.wf1) / 4, (L.total / 4) * 4, L.total, 1, 78);
}

const float *dist_current_grads(const fb_dist *d) { return d ? d->xgrads[d->step & 1] : nullptr; }

extern "C" int fb_dist_create(int rank, int world, long long n_floats, fb_dist **out) {
    FB_REQUIRE(out != nullptr && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world && n_floats > 0, "fb_dist_create: bad argument");
    fb_dist *d = new (std::nothrow) fb_dist();
    FB_REQUIRE(d != nullptr, "fb_dist_create: out of host memory");
    d->rank = rank; d->world = world; d->n = (size_t)n_floats; d->n_opened = 0; d->step = 0; d->debug_mask = 0;
    const size_t bytes = ((size_t)n_floats * sizeof(float) + 255) / 256 * 256;
    for (int k = 0; k < 2; k++) { FB_CUDA_OK(cudaMalloc(&d->xgrads[k], bytes)); FB_CUDA_OK(cudaMemset(d->xgrads[k], 0, bytes)); }
    FB_CUDA_OK(cudaMalloc(&d->flags, 512));
    FB_CUDA_OK(cudaMemset(d->flags, 0, 512));
    FB_CUDA_OK(cudaMalloc(&d->red, bytes));
    FB_CUDA_OK(cudaMemset(d->red, 0, bytes));
    FB_CUDA_OK(cudaMalloc(&d->grid_count, 256));
    FB_CUDA_OK(cudaMemset(d->grid_count, 0, 256));
    FB_CUDA_OK(cudaMalloc(&d->dev_steps, 256));
    FB_CUDA_OK(cudaMemset(d->dev_steps, 0, 256));
    d->dev_done = reinterpret_cast<unsigned int *>(d->dev_steps) + 16;
    d->stamps = nullptr;
    d->two_shot = world >= 4; d->ts_steps = 0;
    for (int q = 0; q < kMaxRanks; q++) { d->peer_grads[0][q] = d->peer_grads[1][q] = nullptr; d->peer_flags[q] = nullptr; d->peer_red[q] = nullptr; }
    d->peer_grads[0][rank] = d->xgrads[0]; d->peer_grads[1][rank] = d->xgrads[1]; d->peer_flags[rank] = d->flags; d->peer_red[rank] = d->red;
    *out = d;
    return FB_OK;
}

extern "C" int fb_dist_destroy(fb_dist *d) {
    if (!d) return FB_OK;
    for (int k = 0; k < d->n_opened; k++) cudaIpcCloseMemHandle(d->opened[k]);
    cudaFree(d->xgrads[0]); cudaFree(d->xgrads[1]); cudaFree(d->flags); cudaFree(d->red); cudaFree(d->grid_count); cudaFree(d->stamps); cudaFree(d->dev_steps);
    delete d;
    return FB_OK;
}

extern "C" int fb_dist_handle_bytes(void) { return 4 * (int)sizeof(cudaIpcMemHandle_t); }

// timing experiments: see fb_dist::debug_mask.  Takes effect for steps captured / launched afterwards.
extern "C" int fb_dist_debug_mask(fb_dist *d, int mask) {
    FB_REQUIRE(d != nullptr, "fb_dist_debug_mask: NULL argument");
    d->debug_mask = mask;
    return FB_OK;
}

// force the one-shot (0) or two-shot (1) exchange (default: two-shot from four ranks up)
extern "C" int fb_dist_set_two_shot(fb_dist *d, int on) {
    FB_REQUIRE(d != nullptr, "fb_dist_set_two_shot: NULL argument");
    d->two_shot = on ? 1 : 0;
    return FB_OK;
}

// this rank's IPC handles (exchange buffer 0, exchange buffer 1, flags, reduced slice), fb_dist_handle_bytes() bytes
extern "C" int fb_dist_handles(fb_dist *d, uint8_t *out_host) {
    FB_REQUIRE(d != nullptr && out_host != nullptr, "fb_dist_handles: NULL argument");
    cudaIpcMemHandle_t h[4];
    FB_CUDA_OK(cudaIpcGetMemHandle(&h[0], d->xgrads[0]));
    FB_CUDA_OK(cudaIpcGetMemHandle(&h[1], d->xgrads[1]));
    FB_CUDA_OK(cudaIpcGetMemHandle(&h[2], d->flags));
    FB_CUDA_OK(cudaIpcGetMemHandle(&h[3], d->red));
    memcpy(out_host, h, sizeof(h));
    return FB_OK;
}

// all ranks' handles, rank-major (world x fb_dist_handle_bytes()); maps every peer's buffers into this process
extern "C" int fb_dist_connect(fb_dist *d, const uint8_t *all_handles_host) {
    FB_REQUIRE(d != nullptr && all_handles_host != nullptr, "fb_dist_connect: NULL argument");
    for (int q = 0; q < d->world; q++) {
        if (q == d->rank) continue;
        cudaIpcMemHandle_t h[4];
        memcpy(h, all_handles_host + (size_t)q * sizeof(h), sizeof(h));
        void *p[4];
        for (int k = 0; k < 4; k++) {
            FB_CUDA_OK(cudaIpcOpenMemHandle(&p[k], h[k], cudaIpcMemLazyEnablePeerAccess));
            d->opened[d->n_opened++] = p[k];
        }
        d->peer_grads[0][q] = (const float *)p[0]; d->peer_grads[1][q] = (const float *)p[1]; d->peer_flags[q] = (uint32_t *)p[2];
        d->peer_red[q] = (const float *)p[3];
    }
    return FB_OK;
}

// test hook: several "ranks" living in one process (no IPC): rank q's buffers are given directly
extern "C" int fb_dist_connect_local(fb_dist *d, int q, fb_dist *peer) {
    FB_REQUIRE(d && peer && q >= 0 && q < d->world && peer->rank == q && peer->n == d->n, "fb_dist_connect_local: bad argument");
    d->peer_grads[0][q] = peer->xgrads[0]; d->peer_grads[1][q] = peer->xgrads[1]; d->peer_flags[q] = peer->flags; d->peer_red[q] = peer->red;
    return FB_OK;
}

// device pointer of the exchange buffer the NEXT fb_dist_adam will read (write this step's gradients there)
extern "C" int fb_dist_grads(fb_dist *d, int parity, float **out) {
    FB_REQUIRE(d != nullptr && out != nullptr && (parity == 0 || parity == 1), "fb_dist_grads: bad argument");
    *out = d->xgrads[parity];
    return FB_OK;
}
extern "C" int fb_dist_parity(const fb_dist *d) { return d ? (int)(d->step & 1) : -1; }
// the training step that carried this exchange in its graph has been enqueued: the next step writes the other buffer
extern "C" int fb_dist_advance(fb_dist *d) {
    FB_REQUIRE(d != nullptr, "fb_dist_advance: NULL argument");
    d->step++;
    return FB_OK;
}

// debug: nanosecond %globaltimer stamps of block 0 at the phase boundaries of the most recent exchange (start, peers'
// gradients published, own slice reduced, slices published, Adam done); the first call switches the stamping on
extern "C" int fb_dist_debug_stamps(fb_dist *d, unsigned long long *out_host5) {
    FB_REQUIRE(d != nullptr && out_host5 != nullptr, "fb_dist_debug_stamps: NULL argument");
    if (!d->stamps) {
        FB_CUDA_OK(cudaMalloc(&d->stamps, 8 * sizeof(unsigned long long)));
        FB_CUDA_OK(cudaMemset(d->stamps, 0, 8 * sizeof(unsigned long long)));
    }
    FB_CUDA_OK(cudaDeviceSynchronize());
    FB_CUDA_OK(cudaMemcpy(out_host5, d->stamps, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return FB_OK;
}

// One optimizer step over the SUM of all ranks' gradients (exchange buffer of the current parity).  wait = 0 skips the
// publish / wait handshake (single-process tests whose "ranks" run one after the other on one GPU).
extern "C" int fb_dist_adam(fb_dist *d, fb_qnet *net, float *params_dev, float *m_dev, float *v_dev, float alpha, float beta1, float beta2,
                            float eps, float grad_scale, float *reduced_out_dev, int wait, void *stream) {
    FB_REQUIRE(d && net && params_dev && m_dev && v_dev, "fb_dist_adam: NULL argument");
    FB_REQUIRE((size_t)net->L.total == d->n, "fb_dist_adam: the exchange buffers were sized for another network");
    const int par = (int)(d->step & 1);
    XArgs x{};
    for (int q = 0; q < d->world; q++) {
        FB_REQUIRE(d->peer_grads[par][q] != nullptr && d->peer_flags[q] != nullptr, "fb_dist_adam: call fb_dist_connect first");
        x.g[q] = d->peer_grads[par][q]; x.peer_flags[q] = d->peer_flags[q]; x.red[q] = d->peer_red[q];
    }
    x.my_flags = d->flags; x.rank = d->rank; x.world = d->world; x.wait = wait; x.step = d->step + 1;
    x.my_red = d->red; x.grid_count = d->grid_count; x.stamps = d->stamps;
    // two-shot needs every rank's slice before anyone reads it: without the cross-rank handshake (single-process test,
    // ranks run one after the other) only the one-shot form is meaningful
    x.two_shot = d->two_shot && d->world > 1 && wait;
    x.slice4 = ((d->n >> 2) + d->world - 1) / d->world;
    if (x.two_shot) x.ts_step = ++d->ts_steps;
    PackedWeights pw{};
    const int repack = tc_online_operands(net, &pw) ? 1 : 0;
    // the two-shot form has a grid barrier: every CTA must be resident.  Two CTAs per SM fit (launch bounds); the grid is sized from
    // the device, not hard-coded (the graph-resident bucket kernels, adam_xbucket_kernel, have no barrier at all)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    adam_xreduce_kernel<<<2 * sms, 256, 0, (cudaStream_t)stream>>>(params_dev, m_dev, v_dev, d->n, x, alpha, beta1, beta2, eps, grad_scale, reduced_out_dev,
                                                               repack, net->L, pw);
    FB_CUDA_OK(cudaGetLastError());
    d->step++;
    for (int s = 0; s < 2; s++) if (net->packed_src[s] == params_dev) net->packed_src[s] = nullptr;
    if (repack) net->packed_src[0] = params_dev;          // slot 0 mirrors the updated vector: the next step needs no pack kernel
    return FB_OK;
}
