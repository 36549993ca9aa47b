// head2.cuh -- streaming forward of the dual-output head (head/flood_head.py:131-202) for the f16x3 step.
//
// Same four dependent sweeps as head.cuh (three LayerNorm([16,H,W]) levels, network_blocks.py:93-94), restructured for
// bandwidth: a thread owns TWO consecutive pixels and walks the INPUT channels (one 8-byte load per map and channel, the
// 16 x 16 convolution accumulated as rank-1 updates), and from the second level on the independent cls / reg branches run
// in different blocks (blockIdx.y).  Few registers per thread -> many loads in flight; measured ~2x faster than the
// thread-per-pixel kernels at 500 x 500 (profiles/).  Statistics: (sum, sum of squares) partials per block, merged in a
// fixed order by the last block (stats_finalize_last_cta), all-reduced over NVLink when the grid is sharded.
#pragma once
#include "head.cuh"

namespace urnn {
namespace head2 {

constexpr int PX = 2;
constexpr int TPB = 256;

__device__ __forceinline__ void block_stats(float s, float q, const StatSink& sink, int set, int nsets_level, int first_set) {
    __shared__ float red[2][TPB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s = warp_sum(s); q = warp_sum(q);
    if (lane == 0) { red[0][warp] = s; red[1][warp] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < TPB / 32; ++i) { a += red[0][i]; b += red[1][i]; }
        sink.partial[(size_t)set * sink.stride + blockIdx.x] = make_float2(a, b);
    }
    StatSink t = sink;                      // this level's contiguous range of sets; all blocks of the launch take a ticket
    t.partial += (size_t)first_set * t.stride;
    t.total += first_set;
    t.nsets = nsets_level;
    stats_finalize_last_cta(t, gridDim.x, gridDim.x * gridDim.y, nullptr);
}

// level 0: statistics of the stem convolution
__global__ void __launch_bounds__(TPB) stem_stats_kernel(HeadDev hd, const float* __restrict__ feat, long npix) {
    __shared__ __align__(16) float w[256];
    w[threadIdx.x] = __ldg(hd.p.conv_w[0] + threadIdx.x);
    __syncthreads();
    const long p0 = ((long)blockIdx.x * TPB + threadIdx.x) * PX;
    float s = 0.f, q = 0.f;
    if (p0 < npix) {
        float acc[16][PX];
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r][0] = acc[r][1] = 0.f;
#pragma unroll 4
        for (int c = 0; c < 16; ++c) {
            const float2 v = __ldcs(reinterpret_cast<const float2*>(feat + (long)c * hd.plane + p0));
#pragma unroll
            for (int r = 0; r < 16; ++r) { const float ww = w[r * 16 + c]; acc[r][0] = fmaf(ww, v.x, acc[r][0]); acc[r][1] = fmaf(ww, v.y, acc[r][1]); }
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) { s += acc[r][0] + acc[r][1]; q = fmaf(acc[r][0], acc[r][0], fmaf(acc[r][1], acc[r][1], q)); }
    }
    block_stats(s, q, hd.sink, 0, 1, 0);
}

// level 1: stem conv -> LayerNorm 0 -> SiLU -> cls_convs.0 / reg_convs.0 (blockIdx.y) -> pre-norm map + statistics (sets 1, 2)
__global__ void __launch_bounds__(TPB) stage1_kernel(HeadDev hd, const float* __restrict__ feat, float* __restrict__ bufA, float* __restrict__ bufB, long npix) {
    __shared__ __align__(16) float w0[256], w1[256];
    const int br = blockIdx.y;                                   // 0: cls, 1: reg
    w0[threadIdx.x] = __ldg(hd.p.conv_w[0] + threadIdx.x);
    w1[threadIdx.x] = __ldg(hd.p.conv_w[br == 0 ? 1 : 3] + threadIdx.x);
    __syncthreads();
    float mean, rstd;
    mean_rstd(hd.sink.total, 0, hd.count, hd.eps, mean, rstd);
    const long p0 = ((long)blockIdx.x * TPB + threadIdx.x) * PX;
    float s = 0.f, q = 0.f;
    if (p0 < npix) {
        float t[16][PX];
#pragma unroll
        for (int r = 0; r < 16; ++r) t[r][0] = t[r][1] = 0.f;
#pragma unroll 4
        for (int c = 0; c < 16; ++c) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(feat + (long)c * hd.plane + p0));
#pragma unroll
            for (int r = 0; r < 16; ++r) { const float ww = w0[r * 16 + c]; t[r][0] = fmaf(ww, v.x, t[r][0]); t[r][1] = fmaf(ww, v.y, t[r][1]); }
        }
        float acc[16][PX];
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r][0] = acc[r][1] = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float2 g = __ldg(reinterpret_cast<const float2*>(hd.p.ln_w[0] + (long)c * hd.plane + p0));
            const float2 b = __ldg(reinterpret_cast<const float2*>(hd.p.ln_b[0] + (long)c * hd.plane + p0));
            const float a0 = silu_fast(fmaf((t[c][0] - mean) * rstd, g.x, b.x)), a1 = silu_fast(fmaf((t[c][1] - mean) * rstd, g.y, b.y));
#pragma unroll
            for (int r = 0; r < 16; ++r) { const float ww = w1[r * 16 + c]; acc[r][0] = fmaf(ww, a0, acc[r][0]); acc[r][1] = fmaf(ww, a1, acc[r][1]); }
        }
        float* out = br == 0 ? bufA : bufB;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            *reinterpret_cast<float2*>(out + (long)r * hd.plane + p0) = make_float2(acc[r][0], acc[r][1]);
            s += acc[r][0] + acc[r][1]; q = fmaf(acc[r][0], acc[r][0], fmaf(acc[r][1], acc[r][1], q));
        }
    }
    block_stats(s, q, hd.sink, 1 + br, 2, 1);
}

// level 2: LayerNorm (cls0 | reg0) -> SiLU -> cls_convs.1 / reg_convs.1 -> pre-norm map (in place) + statistics (sets 3, 4)
__global__ void __launch_bounds__(TPB) stage2_kernel(HeadDev hd, float* __restrict__ bufA, float* __restrict__ bufB, long npix) {
    __shared__ __align__(16) float w[256];
    const int br = blockIdx.y;
    w[threadIdx.x] = __ldg(hd.p.conv_w[br == 0 ? 2 : 4] + threadIdx.x);
    __syncthreads();
    float mean, rstd;
    mean_rstd(hd.sink.total, 1 + br, hd.count, hd.eps, mean, rstd);
    const float* lw = hd.p.ln_w[br == 0 ? 1 : 3];
    const float* lb = hd.p.ln_b[br == 0 ? 1 : 3];
    float* buf = br == 0 ? bufA : bufB;
    const long p0 = ((long)blockIdx.x * TPB + threadIdx.x) * PX;
    float s = 0.f, q = 0.f;
    if (p0 < npix) {
        float acc[16][PX];
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r][0] = acc[r][1] = 0.f;
#pragma unroll 8
        for (int c = 0; c < 16; ++c) {
            const float2 x = *reinterpret_cast<const float2*>(buf + (long)c * hd.plane + p0);
            const float2 g = __ldg(reinterpret_cast<const float2*>(lw + (long)c * hd.plane + p0));
            const float2 b = __ldg(reinterpret_cast<const float2*>(lb + (long)c * hd.plane + p0));
            const float a0 = silu_fast(fmaf((x.x - mean) * rstd, g.x, b.x)), a1 = silu_fast(fmaf((x.y - mean) * rstd, g.y, b.y));
#pragma unroll
            for (int r = 0; r < 16; ++r) { const float ww = w[r * 16 + c]; acc[r][0] = fmaf(ww, a0, acc[r][0]); acc[r][1] = fmaf(ww, a1, acc[r][1]); }
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            *reinterpret_cast<float2*>(buf + (long)r * hd.plane + p0) = make_float2(acc[r][0], acc[r][1]);
            s += acc[r][0] + acc[r][1]; q = fmaf(acc[r][0], acc[r][0], fmaf(acc[r][1], acc[r][1], q));
        }
    }
    block_stats(s, q, hd.sink, 3 + br, 2, 3);
}

// level 3: LayerNorm (cls1 | reg1) -> SiLU -> prediction convs -> sigmoid / LeakyReLU -> depth * (prob >= thr), prob
__global__ void __launch_bounds__(TPB) stage3_kernel(HeadDev hd, const float* __restrict__ bufA, const float* __restrict__ bufB,
                                                     float* __restrict__ out, long npix) {
    __shared__ float pw[2][16];
    if (threadIdx.x < 16) { pw[0][threadIdx.x] = __ldg(hd.p.cls_pred_w + threadIdx.x); pw[1][threadIdx.x] = __ldg(hd.p.reg_pred_w + threadIdx.x); }
    __syncthreads();
    float mc, rc, mr, rr;
    mean_rstd(hd.sink.total, 3, hd.count, hd.eps, mc, rc);
    mean_rstd(hd.sink.total, 4, hd.count, hd.eps, mr, rr);
    const long p0 = ((long)blockIdx.x * TPB + threadIdx.x) * PX;
    if (p0 >= npix) return;
    const float bc = __ldg(hd.p.cls_pred_b), brg = __ldg(hd.p.reg_pred_b);
    float pc[PX] = {bc, bc}, pr[PX] = {brg, brg};
#pragma unroll 8
    for (int c = 0; c < 16; ++c) {
        const long o = (long)c * hd.plane + p0;
        const float2 xa = __ldcs(reinterpret_cast<const float2*>(bufA + o)), xb = __ldcs(reinterpret_cast<const float2*>(bufB + o));
        const float2 ga = __ldg(reinterpret_cast<const float2*>(hd.p.ln_w[2] + o)), ba = __ldg(reinterpret_cast<const float2*>(hd.p.ln_b[2] + o));
        const float2 gb = __ldg(reinterpret_cast<const float2*>(hd.p.ln_w[4] + o)), bb = __ldg(reinterpret_cast<const float2*>(hd.p.ln_b[4] + o));
        pc[0] = fmaf(pw[0][c], silu_fast(fmaf((xa.x - mc) * rc, ga.x, ba.x)), pc[0]);
        pc[1] = fmaf(pw[0][c], silu_fast(fmaf((xa.y - mc) * rc, ga.y, ba.y)), pc[1]);
        pr[0] = fmaf(pw[1][c], silu_fast(fmaf((xb.x - mr) * rr, gb.x, bb.x)), pr[0]);
        pr[1] = fmaf(pw[1][c], silu_fast(fmaf((xb.y - mr) * rr, gb.y, bb.y)), pr[1]);
    }
    float d[PX], pb[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        pb[i] = sigmoid_fast(pc[i]);
        const float depth = lrelu(pr[i], hd.slope);
        d[i] = (pb[i] >= hd.cls_thred) ? depth : depth * 0.0f;      // depth * mask, flood_head.py:201-202
    }
    *reinterpret_cast<float2*>(out + p0) = make_float2(d[0], d[1]);
    *reinterpret_cast<float2*>(out + hd.plane + p0) = make_float2(pb[0], pb[1]);
}

}  // namespace head2
}  // namespace urnn
