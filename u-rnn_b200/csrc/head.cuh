// head.cuh -- the dual-output head (head/flood_head.py:131-202) as four recompute sweeps.
// Every op is pixel-local except the five LayerNorm([16,H,W]) statistics, which form three dependent
// levels (stem | cls0,reg0 | cls1,reg1).  Each sweep recomputes the 16-wide chain from the decoder
// features (16 floats per cell) up to the next un-normalised tensor and reduces its (sum, sumsq); the last
// sweep produces the outputs.  Nothing but the 2-channel result and ~100 bytes of statistics is written.
#pragma once
#include "urnn_common.cuh"

namespace urnn {

struct HeadDev {
    urnn_head_params p;
    float cls_thred, eps, slope;
    long plane;          // H*W
    double count;        // 16*H*W  (global element count of one LayerNorm)
    StatSink sink;       // 5 sets by level: 0 stem | 1 cls0, 2 reg0 | 3 cls1, 4 reg1
};

__device__ __forceinline__ void matvec16(const float* __restrict__ w /*smem 16x16 row-major*/,
                                         const float (&v)[16], float (&o)[16]) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            float4 ww = *reinterpret_cast<const float4*>(w + r * 16 + c);
            a = fmaf(ww.x, v[c], a); a = fmaf(ww.y, v[c + 1], a);
            a = fmaf(ww.z, v[c + 2], a); a = fmaf(ww.w, v[c + 3], a);
        }
        o[r] = a;
    }
}

__device__ __forceinline__ void ln_silu16(float (&v)[16], float mean, float rstd,
                                          const float* __restrict__ lw, const float* __restrict__ lb,
                                          long plane, long pix) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        float g = __ldg(lw + c * plane + pix), b = __ldg(lb + c * plane + pix);
        float y = fmaf((v[c] - mean) * rstd, g, b);
        v[c] = silu_fast(y);
    }
}

__device__ __forceinline__ void mean_rstd(const double2* total, int set, double count, float eps,
                                          float& mean, float& rstd) {
    double2 t = total[set];
    double m = t.x / count;
    double var = t.y / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// LEVEL 0: stats of stem conv; 1: stats of cls0/reg0 convs; 2: stats of cls1/reg1 convs; 3: outputs.
template <int LEVEL>
__global__ void __launch_bounds__(128) head_kernel(HeadDev hd, const float* __restrict__ feat,
                                                   float* __restrict__ out, int npix) {
    __shared__ __align__(16) float w[5][256];
    __shared__ float pw[2][16];
    __shared__ float red[2][2][4];
    // only the mat-vecs this level evaluates: stem | + cls_convs.0, reg_convs.0 (indices 1, 3) | all five
    constexpr int NEED = (LEVEL == 0) ? 0x01 : (LEVEL == 1 ? 0x0B : 0x1F);
    for (int i = threadIdx.x; i < 5 * 256; i += blockDim.x)
        if ((NEED >> (i / 256)) & 1) w[i / 256][i % 256] = __ldg(hd.p.conv_w[i / 256] + (i % 256));
    if (threadIdx.x < 16) {
        pw[0][threadIdx.x] = __ldg(hd.p.cls_pred_w + threadIdx.x);
        pw[1][threadIdx.x] = __ldg(hd.p.reg_pred_w + threadIdx.x);
    }
    __syncthreads();

    const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = pix < npix;
    const long plane = hd.plane;
    float sA = 0.f, ssA = 0.f, sB = 0.f, ssB = 0.f;    // stat accumulators (A: stem/cls, B: reg)

    if (valid) {
        float x[16], t[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) x[c] = __ldg(feat + c * plane + pix);
        matvec16(w[0], x, t);
        if constexpr (LEVEL == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) { sA += t[c]; ssA = fmaf(t[c], t[c], ssA); }
        } else {
            float mean, rstd;
            mean_rstd(hd.sink.total, 0, hd.count, hd.eps, mean, rstd);
            ln_silu16(t, mean, rstd, hd.p.ln_w[0], hd.p.ln_b[0], plane, pix);   // t = stem output
            float uc[16], ur[16];
            matvec16(w[1], t, uc);
            matvec16(w[3], t, ur);
            if constexpr (LEVEL == 1) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    sA += uc[c]; ssA = fmaf(uc[c], uc[c], ssA);
                    sB += ur[c]; ssB = fmaf(ur[c], ur[c], ssB);
                }
            } else {
                mean_rstd(hd.sink.total, 1, hd.count, hd.eps, mean, rstd);
                ln_silu16(uc, mean, rstd, hd.p.ln_w[1], hd.p.ln_b[1], plane, pix);
                mean_rstd(hd.sink.total, 2, hd.count, hd.eps, mean, rstd);
                ln_silu16(ur, mean, rstd, hd.p.ln_w[3], hd.p.ln_b[3], plane, pix);
                float vc[16], vr[16];
                matvec16(w[2], uc, vc);
                matvec16(w[4], ur, vr);
                if constexpr (LEVEL == 2) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        sA += vc[c]; ssA = fmaf(vc[c], vc[c], ssA);
                        sB += vr[c]; ssB = fmaf(vr[c], vr[c], ssB);
                    }
                } else {
                    mean_rstd(hd.sink.total, 3, hd.count, hd.eps, mean, rstd);
                    ln_silu16(vc, mean, rstd, hd.p.ln_w[2], hd.p.ln_b[2], plane, pix);
                    mean_rstd(hd.sink.total, 4, hd.count, hd.eps, mean, rstd);
                    ln_silu16(vr, mean, rstd, hd.p.ln_w[4], hd.p.ln_b[4], plane, pix);
                    float pc = __ldg(hd.p.cls_pred_b), pr = __ldg(hd.p.reg_pred_b);
#pragma unroll
                    for (int c = 0; c < 16; ++c) { pc = fmaf(pw[0][c], vc[c], pc); pr = fmaf(pw[1][c], vr[c], pr); }
                    float prob = sigmoid_fast(pc);
                    float depth = lrelu(pr, hd.slope);
                    out[pix] = (prob >= hd.cls_thred) ? depth : depth * 0.0f;   // depth * mask, flood_head.py:201-202
                    out[plane + pix] = prob;
                }
            }
        }
    }
    if constexpr (LEVEL < 3) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        sA = warp_sum(sA); ssA = warp_sum(ssA); sB = warp_sum(sB); ssB = warp_sum(ssB);
        if (lane == 0) { red[0][0][warp] = sA; red[0][1][warp] = ssA; red[1][0][warp] = sB; red[1][1][warp] = ssB; }
        __syncthreads();
        constexpr int FIRST = (LEVEL == 0) ? 0 : (LEVEL == 1 ? 1 : 3);
        constexpr int NSET = (LEVEL == 0) ? 1 : 2;
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
            for (int i = 0; i < 4; ++i) { a += red[0][0][i]; b += red[0][1][i]; c += red[1][0][i]; d += red[1][1][i]; }
            hd.sink.partial[(size_t)FIRST * hd.sink.stride + blockIdx.x] = make_float2(a, b);
            if (NSET == 2) hd.sink.partial[(size_t)(FIRST + 1) * hd.sink.stride + blockIdx.x] = make_float2(c, d);
        }
        StatSink s = hd.sink;                      // this level's contiguous range of sets
        s.partial += (size_t)FIRST * s.stride;
        s.total += FIRST;
        s.nsets = NSET;
        stats_finalize_last_cta(s, gridDim.x, gridDim.x, nullptr);
    }
}

}  // namespace urnn

// ------------------------------------------------------------------------------------------------ staged forward
// Same arithmetic as head_kernel<LEVEL> (bit-identical per pixel: same matvec / LayerNorm / SiLU order), but every level
// starts from the previous level's un-normalised 2 x 16 channels (stored fp32 in the workspace) instead of recomputing
// the chain from the decoder features: 6 instead of 13 16x16 mat-vecs per pixel over the three sweeps, one LayerNorm
// level of registers per kernel, two pixels per thread (shared weight loads, independent dependency chains).
namespace urnn {


template <int PX>
__device__ __forceinline__ void matvec16_px(const float* __restrict__ w, const float (&v)[PX][16], float (&o)[PX][16]) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        float a[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) a[p] = 0.f;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            const float4 ww = *reinterpret_cast<const float4*>(w + r * 16 + c);
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                a[p] = fmaf(ww.x, v[p][c], a[p]); a[p] = fmaf(ww.y, v[p][c + 1], a[p]);
                a[p] = fmaf(ww.z, v[p][c + 2], a[p]); a[p] = fmaf(ww.w, v[p][c + 3], a[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) o[p][r] = a[p];
    }
}

// STAGE 1: feat -> (cls0, reg0) pre-norm maps + their statistics (needs set 0)
// STAGE 2: (cls0, reg0) -> (cls1, reg1) pre-norm maps, in place, + statistics (needs sets 1, 2)
// STAGE 3: (cls1, reg1) -> outputs (needs sets 3, 4)
template <int STAGE, int HEAD_PX>
__global__ void __launch_bounds__(128, HEAD_PX == 2 ? 3 : 5) head_stage_kernel(HeadDev hd, const float* __restrict__ feat, float* __restrict__ bufA,
                                                         float* __restrict__ bufB, float* __restrict__ out, int npix) {
    __shared__ __align__(16) float w[3][256];
    __shared__ float pw[2][16];
    __shared__ float red[2][2][4];
    constexpr int W0 = (STAGE == 1) ? 0 : 2, W1 = (STAGE == 1) ? 1 : 2, W2 = (STAGE == 1) ? 3 : 4;   // conv_w indices used
    if constexpr (STAGE < 3) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            w[0][i] = __ldg(hd.p.conv_w[W0] + i); w[1][i] = __ldg(hd.p.conv_w[W1] + i); w[2][i] = __ldg(hd.p.conv_w[W2] + i);
        }
    } else if (threadIdx.x < 16) {
        pw[0][threadIdx.x] = __ldg(hd.p.cls_pred_w + threadIdx.x);
        pw[1][threadIdx.x] = __ldg(hd.p.reg_pred_w + threadIdx.x);
    }
    __syncthreads();
    const long plane = hd.plane;
    long pix[HEAD_PX]; bool valid[HEAD_PX];
#pragma unroll
    for (int p = 0; p < HEAD_PX; ++p) {
        pix[p] = (long)blockIdx.x * (128 * HEAD_PX) + p * 128 + threadIdx.x;
        valid[p] = pix[p] < npix;
        if (!valid[p]) pix[p] = 0;                          // compute on a valid address, mask the results
    }
    float sA = 0.f, ssA = 0.f, sB = 0.f, ssB = 0.f;
    float a[HEAD_PX][16], b[HEAD_PX][16];

    if constexpr (STAGE == 1) {
        float x[HEAD_PX][16];
#pragma unroll
        for (int p = 0; p < HEAD_PX; ++p)
#pragma unroll
            for (int c = 0; c < 16; ++c) x[p][c] = __ldg(feat + c * plane + pix[p]);
        matvec16_px<HEAD_PX>(w[0], x, a);                   // stem conv
        float mean, rstd;
        mean_rstd(hd.sink.total, 0, hd.count, hd.eps, mean, rstd);
#pragma unroll
        for (int p = 0; p < HEAD_PX; ++p) ln_silu16(a[p], mean, rstd, hd.p.ln_w[0], hd.p.ln_b[0], plane, pix[p]);
        float (&uc)[HEAD_PX][16] = x;
        matvec16_px<HEAD_PX>(w[1], a, uc);                  // cls_convs.0
        matvec16_px<HEAD_PX>(w[2], a, b);                   // reg_convs.0
#pragma unroll
        for (int p = 0; p < HEAD_PX; ++p) {
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                s0 += uc[p][c]; q0 = fmaf(uc[p][c], uc[p][c], q0); s1 += b[p][c]; q1 = fmaf(b[p][c], b[p][c], q1);
                if (valid[p]) { bufA[c * plane + pix[p]] = uc[p][c]; bufB[c * plane + pix[p]] = b[p][c]; }
            }
            if (valid[p]) { sA += s0; ssA += q0; sB += s1; ssB += q1; }
        }
    } else {
#pragma unroll
        for (int p = 0; p < HEAD_PX; ++p)
#pragma unroll
            for (int c = 0; c < 16; ++c) { a[p][c] = __ldg(bufA + c * plane + pix[p]); b[p][c] = __ldg(bufB + c * plane + pix[p]); }
        constexpr int LA = (STAGE == 2) ? 1 : 2, LB = (STAGE == 2) ? 3 : 4;      // LayerNorm parameter sets (cls, reg)
        constexpr int SA = (STAGE == 2) ? 1 : 3, SB = (STAGE == 2) ? 2 : 4;      // statistic sets
        float mean, rstd;
        mean_rstd(hd.sink.total, SA, hd.count, hd.eps, mean, rstd);
#pragma unroll
        for (int p = 0; p < HEAD_PX; ++p) ln_silu16(a[p], mean, rstd, hd.p.ln_w[LA], hd.p.ln_b[LA], plane, pix[p]);
        mean_rstd(hd.sink.total, SB, hd.count, hd.eps, mean, rstd);
#pragma unroll
        for (int p = 0; p < HEAD_PX; ++p) ln_silu16(b[p], mean, rstd, hd.p.ln_w[LB], hd.p.ln_b[LB], plane, pix[p]);
        if constexpr (STAGE == 2) {
            float vc[HEAD_PX][16], vr[HEAD_PX][16];
            matvec16_px<HEAD_PX>(w[0], a, vc);              // cls_convs.1
            matvec16_px<HEAD_PX>(w[2], b, vr);              // reg_convs.1
#pragma unroll
            for (int p = 0; p < HEAD_PX; ++p) {
                float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    s0 += vc[p][c]; q0 = fmaf(vc[p][c], vc[p][c], q0); s1 += vr[p][c]; q1 = fmaf(vr[p][c], vr[p][c], q1);
                    if (valid[p]) { bufA[c * plane + pix[p]] = vc[p][c]; bufB[c * plane + pix[p]] = vr[p][c]; }
                }
                if (valid[p]) { sA += s0; ssA += q0; sB += s1; ssB += q1; }
            }
        } else {
#pragma unroll
            for (int p = 0; p < HEAD_PX; ++p) {
                float pc = __ldg(hd.p.cls_pred_b), pr = __ldg(hd.p.reg_pred_b);
#pragma unroll
                for (int c = 0; c < 16; ++c) { pc = fmaf(pw[0][c], a[p][c], pc); pr = fmaf(pw[1][c], b[p][c], pr); }
                const float prob = sigmoid_fast(pc);
                const float depth = lrelu(pr, hd.slope);
                if (valid[p]) {
                    out[pix[p]] = (prob >= hd.cls_thred) ? depth : depth * 0.0f;   // depth * mask, flood_head.py:201-202
                    out[plane + pix[p]] = prob;
                }
            }
        }
    }
    if constexpr (STAGE < 3) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        sA = warp_sum(sA); ssA = warp_sum(ssA); sB = warp_sum(sB); ssB = warp_sum(ssB);
        if (lane == 0) { red[0][0][warp] = sA; red[0][1][warp] = ssA; red[1][0][warp] = sB; red[1][1][warp] = ssB; }
        __syncthreads();
        constexpr int FIRST = (STAGE == 1) ? 1 : 3;
        if (threadIdx.x == 0) {
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
            for (int i = 0; i < 4; ++i) { s0 += red[0][0][i]; q0 += red[0][1][i]; s1 += red[1][0][i]; q1 += red[1][1][i]; }
            hd.sink.partial[(size_t)FIRST * hd.sink.stride + blockIdx.x] = make_float2(s0, q0);
            hd.sink.partial[(size_t)(FIRST + 1) * hd.sink.stride + blockIdx.x] = make_float2(s1, q1);
        }
        StatSink s = hd.sink;
        s.partial += (size_t)FIRST * s.stride;
        s.total += FIRST;
        s.nsets = 2;
        stats_finalize_last_cta(s, gridDim.x, gridDim.x, nullptr);
    }
}

}  // namespace urnn
