// urnn_common.cuh -- shared helpers for liburnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/urnn_b200.h"

namespace urnn {

// ---- error reporting (thread-local message, C-ABI return codes) -------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define URNN_CHECK_ARG(cond, ...)                                  \
    do { if (!(cond)) { ::urnn::set_error(__VA_ARGS__); return URNN_E_INVALID; } } while (0)

#define URNN_CUDA(call)                                                                     \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                                  \
        ::urnn::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        return URNN_E_CUDA; } } while (0)

#define URNN_LAUNCH_CHECK()                                                                 \
    do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) {                      \
        ::urnn::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
        return URNN_E_CUDA; } ::urnn::count_launch(); } while (0)

#define URNN_TRY(call) do { int r_ = (call); if (r_ != URNN_OK) return r_; } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
    char* base; size_t size; size_t off;
    Arena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
    template <class T> T* take(size_t count) {
        off = align_up(off, 256);
        T* r = (T*)(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= size; }
};

// ---- device math -------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// reduced-precision modes only: ex2.approx + rcp.approx (about 2 ulp), the result is rounded to bf16 anyway
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// ex2.approx / rcp.approx have a relative error of ~2 ulp: |err| of silu/sigmoid stays below 1e-6 on O(1) values,
// inside the fp32 parity tolerance (atol 1e-5)
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// tanh(x) = 1 - 2/(1 + e^{2x}) with ex2.approx / rcp.approx: absolute error ~1e-6, saturates correctly at +-1
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }
__device__ __forceinline__ float lrelu(float x, float slope) { return x >= 0.f ? x : x * slope; }
__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- normalisation statistics ------------------------------------------------------------------
// A "stat set" is one normalisation group: (sum, sum of squares) over its elements.  Producers write one
// float2 partial per CTA into partial[set * stride + cta]; the last CTA to finish (ticket counter) sums the
// partials of every set in a fixed order in double and writes total[set]; the result is therefore
// run-to-run deterministic.  The counter resets itself so the same control block can be reused.
// Spatial sharding (one process per GPU, the H x W grid split into row bands): normalisation statistics are the only
// quantity exchanged per step.  Every rank owns an exchange buffer that its peers map through CUDA IPC; the last CTA of
// a statistics-producing kernel stores its totals into slot `rank` of EVERY peer's buffer over NVLink, publishes a
// flag (st.release.sys), waits for the flags of all ranks (ld.acquire.sys) and sums the slots in rank order, so all
// ranks obtain bit-identical totals with one NVLink round trip and no host involvement (one-shot all-reduce).
constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_RING = 64;        // exchanges in flight are at most 1 apart between ranks; the ring avoids resets
constexpr int COMM_MAX_SETS = 8;     // statistics sets per exchange (<= 8 GroupNorm groups / 2 LayerNorms)
constexpr int COMM_WORDS = 8;        // 32-bit payload words per set (6 used: three doubles)
struct CommDev {
    int world, rank;                               // world <= 1: no exchange
    unsigned long long* slots[COMM_MAX_WORLD];     // peer p: slots[p][((ring*world + src_rank)*COMM_MAX_SETS + set)*COMM_WORDS + word]
    unsigned* flags[COMM_MAX_WORLD];               // (unused by the LL protocol; kept in the buffer layout)
    unsigned* seq;                                 // local exchange counter (device memory)
};

struct StatSink {
    float2*   partial;   // [nsets][stride]
    double2*  total;     // [nsets]
    unsigned* counter;   // one ticket counter for this launch
    int       nsets;
    int       stride;    // >= number of CTAs contributing to a set
    CommDev   comm;      // cross-GPU exchange of the totals (world <= 1: none)
};

// One-shot all-gather of `nw` 32-bit words per set from every rank, "LL" style: every 8-byte store carries 4 bytes of
// payload and the 4-byte epoch of this exchange, so a word is valid as soon as its epoch matches -- one NVLink hop, no
// system fence, no separate flag (the fence + flag round trip was ~2/3 of the exposed exchange latency).  The epoch grows
// with every wrap of the ring, so stale words never match.  Called by all threads of ONE CTA.
//   in : words_in[set*nw + w]           (shared or global memory of this CTA)
//   out: words_out[(r*nsets + set)*nw + w] for every rank r (shared memory, >= world*nsets*nw words)
__device__ __forceinline__ void ll_allgather(const CommDev& c, const unsigned* words_in, int nsets, int nw, unsigned* words_out) {
    __shared__ unsigned sh_seq;
    if (threadIdx.x == 0) sh_seq = *c.seq;
    __syncthreads();
    const unsigned seq = sh_seq, ring = seq % COMM_RING, epoch = seq / COMM_RING + 1;
    const int per_rank = nsets * nw;
    // 1. my words -> slot `rank` of every rank (including myself)
    for (int i = threadIdx.x; i < c.world * per_rank; i += blockDim.x) {
        const int peer = i / per_rank, j = i - peer * per_rank, set = j / nw, w = j - set * nw;
        unsigned long long* dst = c.slots[peer] + ((size_t)(ring * c.world + c.rank) * COMM_MAX_SETS + set) * COMM_WORDS + w;
        const unsigned long long v = (unsigned long long)words_in[j] | ((unsigned long long)epoch << 32);
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
    }
    // 2. every rank's words from MY buffer
    for (int i = threadIdx.x; i < c.world * per_rank; i += blockDim.x) {
        const int r = i / per_rank, j = i - r * per_rank, set = j / nw, w = j - set * nw;
        const unsigned long long* src = c.slots[c.rank] + ((size_t)(ring * c.world + r) * COMM_MAX_SETS + set) * COMM_WORDS + w;
        unsigned long long v;
        for (;;) {
            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
            if ((unsigned)(v >> 32) == epoch) break;
            __nanosleep(32);
        }
        words_out[i] = (unsigned)v;
    }
    if (threadIdx.x == 0) *c.seq = seq + 1;
    __syncthreads();
}

// One-shot all-reduce of s.total[0..nsets) across ranks; called by all threads of the last CTA after the local
// totals are written (and a __syncthreads()).  Deterministic: fixed summation order (rank 0, 1, ...).
__device__ __forceinline__ void stats_exchange(const StatSink& s) {
    const CommDev& c = s.comm;
    if (c.world <= 1) return;
    __shared__ unsigned xin[COMM_MAX_SETS * 4], xout[COMM_MAX_WORLD * COMM_MAX_SETS * 4];
    if (threadIdx.x < s.nsets) {
        const double2 t = s.total[threadIdx.x];
        const unsigned long long a = (unsigned long long)__double_as_longlong(t.x), b = (unsigned long long)__double_as_longlong(t.y);
        xin[4 * threadIdx.x] = (unsigned)a; xin[4 * threadIdx.x + 1] = (unsigned)(a >> 32);
        xin[4 * threadIdx.x + 2] = (unsigned)b; xin[4 * threadIdx.x + 3] = (unsigned)(b >> 32);
    }
    __syncthreads();
    ll_allgather(c, xin, s.nsets, 4, xout);
    if (threadIdx.x < s.nsets) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < c.world; ++r) {
            const unsigned* w = xout + (r * s.nsets + threadIdx.x) * 4;
            a += __longlong_as_double((long long)((unsigned long long)w[0] | ((unsigned long long)w[1] << 32)));
            b += __longlong_as_double((long long)((unsigned long long)w[2] | ((unsigned long long)w[3] << 32)));
        }
        s.total[threadIdx.x] = make_double2(a, b);
    }
    __syncthreads();
}

// Called by ALL threads of the CTA after the CTA's partials are written and made visible.
// ncontrib = CTAs contributing per set, ncta_total = CTAs in the launch.  Optionally the last CTA also
// folds the totals into per-channel affine (scale, shift) so consumers apply y = x*scale + shift:
//   scale[c] = gamma[c]*rstd[g(c)],  shift[c] = beta[c] - mean[g(c)]*scale[c],  g(c) = c / ch_per_set.
struct AffineOut {
    float* scale; float* shift; const float* gamma; const float* beta;
    int channels; int ch_per_set; double count; float eps;
};

__device__ __forceinline__ void stats_finalize_last_cta(const StatSink& s, int ncontrib, unsigned ncta_total,
                                                        const AffineOut* aff) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(s.counter, 1u);
        is_last = (t == ncta_total - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // one warp per set (sets are few: <= 8 GroupNorm groups / 2 LayerNorms): lane-strided double sums in a fixed order,
    // then a fixed shuffle tree -> deterministic, and no block-wide barrier per set
    const int nwarp = blockDim.x >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int set = warp; set < s.nsets; set += nwarp) {
        double a = 0.0, b = 0.0;
        for (int i = lane; i < ncontrib; i += 32) {
            float2 v = __ldcg(&s.partial[(size_t)set * s.stride + i]);
            a += (double)v.x; b += (double)v.y;
        }
        a = warp_sum(a); b = warp_sum(b);
        if (lane == 0) s.total[set] = make_double2(a, b);
    }
    __syncthreads();
    stats_exchange(s);
    if (aff != nullptr) {
        __threadfence();
        for (int c = threadIdx.x; c < aff->channels; c += blockDim.x) {
            double2 t = s.total[c / aff->ch_per_set];
            double mean = t.x / aff->count;
            double var = t.y / aff->count - mean * mean;
            if (var < 0.0) var = 0.0;
            double rstd = 1.0 / sqrt(var + (double)aff->eps);
            double sc = (double)aff->gamma[c] * rstd;
            aff->scale[c] = (float)sc;
            aff->shift[c] = (float)((double)aff->beta[c] - mean * sc);
        }
    }
    if (threadIdx.x == 0) *s.counter = 0u;
}

}  // namespace urnn
