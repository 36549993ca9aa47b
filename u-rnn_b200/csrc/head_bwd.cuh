// head_bwd.cuh -- backward of the dual-output head (head/flood_head.py:131-202) as recompute sweeps.
// LayerNorm([16,H,W]) backward needs two grid-wide sums per normalisation (sum g, sum g*xhat with g = dy*gamma), so the
// three dependent LayerNorm levels of the forward become three more statistic levels in reverse.  Every sweep
// recomputes the 16-wide forward chain of its pixel from the decoder features (thread = pixel), walks the backward
// chain down to its level, accumulates the per-element LayerNorm affine gradients in place and the sums with double
// atomics, and materialises only the 16-channel maps the weight-gradient GEMMs need.
#pragma once
#include "head.cuh"

namespace urnn {

struct HeadBwdDev {
    HeadDev fwd;                 // parameters + forward statistics (sets: 0 stem | 1 cls0, 2 reg0 | 3 cls1, 4 reg1)
    urnn_head_grads g;           // accumulated into
    double* bsum;                // [5][2] backward sums (sum g, sum g*xhat) per LayerNorm set
    double* psum;                // [34] prediction-conv gradients: wcp[16], bcp, wrp[16], brp
    const float* dout;           // (2,H,W)
    float* dfeat;                // (16,H,W)
    float *m0, *m1, *m2, *m3;    // scratch maps (16,H,W) for the weight-gradient GEMMs
};

__device__ __forceinline__ void matvec16T(const float* __restrict__ w, const float (&v)[16], float (&o)[16]) {
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = fmaf(w[r * 16 + c], v[r], o[c]);
    }
}

// forward LayerNorm + SiLU keeping what the backward needs: xhat and the pre-activation y
__device__ __forceinline__ void ln_silu16_keep(const float (&t)[16], float mean, float rstd, const float* lw, const float* lb,
                                               long plane, long pix, float (&xh)[16], float (&y)[16], float (&a)[16]) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        xh[c] = (t[c] - mean) * rstd;
        y[c] = fmaf(xh[c], __ldg(lw + c * plane + pix), __ldg(lb + c * plane + pix));
        a[c] = silu_fast(y[c]);
    }
}
__device__ __forceinline__ float dsilu(float y) { const float s = sigmoid_fast(y); return s * (1.f + y * (1.f - s)); }

// da -> g = dy*gamma (returned in da), accumulates dgamma/dbeta in place (when `write`) and the two sums
__device__ __forceinline__ void ln_bwd_first(float (&da)[16], const float (&y)[16], const float (&xh)[16], const float* lw,
                                             float* dlw, float* dlb, long plane, long pix, bool write, double& s1, double& s2) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float dy = da[c] * dsilu(y[c]);
        if (write) { dlw[c * plane + pix] += dy * xh[c]; dlb[c * plane + pix] += dy; }
        const float g = dy * __ldg(lw + c * plane + pix);
        da[c] = g; s1 += (double)g; s2 += (double)g * (double)xh[c];
    }
}
__device__ __forceinline__ void ln_bwd_second(float (&g)[16], const float (&xh)[16], float rstd, const double* bsum, int set, double count) {
    const float m1 = (float)(bsum[2 * set] / count), m2 = (float)(bsum[2 * set + 1] / count);
#pragma unroll
    for (int c = 0; c < 16; ++c) g[c] = rstd * (g[c] - m1 - xh[c] * m2);
}

// LV 1: sums of LN3/LN4 + prediction-conv gradients; 2: maps for dWc1/dWr1, sums of LN1/LN2; 3: maps for dWc0/dWr0, sums
// of LN0; 4: map for dWs, dfeat.
template <int LV>
__global__ void __launch_bounds__(128) head_bwd_kernel(HeadBwdDev hb, const float* __restrict__ feat, int npix) {
    __shared__ __align__(16) float w[5][256];
    __shared__ float pw[2][16];
    __shared__ double red[34];
    const HeadDev& hd = hb.fwd;
    for (int i = threadIdx.x; i < 5 * 256; i += blockDim.x) w[i / 256][i % 256] = __ldg(hd.p.conv_w[i / 256] + (i % 256));
    if (threadIdx.x < 16) { pw[0][threadIdx.x] = __ldg(hd.p.cls_pred_w + threadIdx.x); pw[1][threadIdx.x] = __ldg(hd.p.reg_pred_w + threadIdx.x); }
    if (threadIdx.x < 34) red[threadIdx.x] = 0.0;
    __syncthreads();
    const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long plane = hd.plane;
    double sA1 = 0.0, sA2 = 0.0, sB1 = 0.0, sB2 = 0.0;
    if (pix < npix) {
        float mean[5], rstd[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) mean_rstd(hd.sink.total, s, hd.count, hd.eps, mean[s], rstd[s]);
        // ---- forward chain (kept values: xhat, y of every LayerNorm; activations feeding a conv)
        float x[16], t[16], xh0[16], y0[16], s_[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) x[c] = __ldg(feat + c * plane + pix);
        matvec16(w[0], x, t);
        ln_silu16_keep(t, mean[0], rstd[0], hd.p.ln_w[0], hd.p.ln_b[0], plane, pix, xh0, y0, s_);
        float xh1[16], y1[16], ac[16], xh2[16], y2[16], ar[16];
        matvec16(w[1], s_, t); ln_silu16_keep(t, mean[1], rstd[1], hd.p.ln_w[1], hd.p.ln_b[1], plane, pix, xh1, y1, ac);
        matvec16(w[3], s_, t); ln_silu16_keep(t, mean[2], rstd[2], hd.p.ln_w[3], hd.p.ln_b[3], plane, pix, xh2, y2, ar);
        float xh3[16], y3[16], fc[16], xh4[16], y4[16], fr[16];
        matvec16(w[2], ac, t); ln_silu16_keep(t, mean[3], rstd[3], hd.p.ln_w[2], hd.p.ln_b[2], plane, pix, xh3, y3, fc);
        matvec16(w[4], ar, t); ln_silu16_keep(t, mean[4], rstd[4], hd.p.ln_w[4], hd.p.ln_b[4], plane, pix, xh4, y4, fr);
        float pc = __ldg(hd.p.cls_pred_b), pr = __ldg(hd.p.reg_pred_b);
#pragma unroll
        for (int c = 0; c < 16; ++c) { pc = fmaf(pw[0][c], fc[c], pc); pr = fmaf(pw[1][c], fr[c], pr); }
        const float prob = sigmoid_fast(pc);
        // ---- backward through the predictions (flood_head.py:164-175,201-202: the mask is a constant)
        const float mask = (prob >= hd.cls_thred) ? 1.f : 0.f;
        const float dpr = __ldg(hb.dout + pix) * mask * (pr > 0.f ? 1.f : hd.slope);
        const float dpc = __ldg(hb.dout + plane + pix) * prob * (1.f - prob);
        float gc[16], grr[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { gc[c] = pw[0][c] * dpc; grr[c] = pw[1][c] * dpr; }
        if constexpr (LV == 1) {
#pragma unroll
            for (int c = 0; c < 16; ++c) { atomicAdd(&red[c], (double)(dpc * fc[c])); atomicAdd(&red[17 + c], (double)(dpr * fr[c])); }
            atomicAdd(&red[16], (double)dpc); atomicAdd(&red[33], (double)dpr);
        }
        ln_bwd_first(gc, y3, xh3, hd.p.ln_w[2], hb.g.ln_w[2], hb.g.ln_b[2], plane, pix, LV == 1, sA1, sA2);
        ln_bwd_first(grr, y4, xh4, hd.p.ln_w[4], hb.g.ln_w[4], hb.g.ln_b[4], plane, pix, LV == 1, sB1, sB2);
        if constexpr (LV >= 2) {
            ln_bwd_second(gc, xh3, rstd[3], hb.bsum, 3, hd.count);      // gc = d vc, grr = d vr
            ln_bwd_second(grr, xh4, rstd[4], hb.bsum, 4, hd.count);
            if constexpr (LV == 2) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    hb.m0[c * plane + pix] = gc[c]; hb.m1[c * plane + pix] = ac[c];
                    hb.m2[c * plane + pix] = grr[c]; hb.m3[c * plane + pix] = ar[c];
                }
            }
            float dac[16], dar[16];
            matvec16T(w[2], gc, dac); matvec16T(w[4], grr, dar);
            sA1 = sA2 = sB1 = sB2 = 0.0;
            ln_bwd_first(dac, y1, xh1, hd.p.ln_w[1], hb.g.ln_w[1], hb.g.ln_b[1], plane, pix, LV == 2, sA1, sA2);
            ln_bwd_first(dar, y2, xh2, hd.p.ln_w[3], hb.g.ln_w[3], hb.g.ln_b[3], plane, pix, LV == 2, sB1, sB2);
            if constexpr (LV >= 3) {
                ln_bwd_second(dac, xh1, rstd[1], hb.bsum, 1, hd.count);   // dac = d uc, dar = d ur
                ln_bwd_second(dar, xh2, rstd[2], hb.bsum, 2, hd.count);
                if constexpr (LV == 3) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        hb.m0[c * plane + pix] = dac[c]; hb.m1[c * plane + pix] = s_[c]; hb.m2[c * plane + pix] = dar[c];
                    }
                }
                float ds[16], ds2[16];
                matvec16T(w[1], dac, ds); matvec16T(w[3], dar, ds2);
#pragma unroll
                for (int c = 0; c < 16; ++c) ds[c] += ds2[c];
                sA1 = sA2 = sB1 = sB2 = 0.0;
                ln_bwd_first(ds, y0, xh0, hd.p.ln_w[0], hb.g.ln_w[0], hb.g.ln_b[0], plane, pix, LV == 3, sA1, sA2);
                if constexpr (LV == 4) {
                    ln_bwd_second(ds, xh0, rstd[0], hb.bsum, 0, hd.count);    // ds = d t0
                    float dx[16];
                    matvec16T(w[0], ds, dx);
#pragma unroll
                    for (int c = 0; c < 16; ++c) { hb.m0[c * plane + pix] = ds[c]; hb.dfeat[c * plane + pix] = dx[c]; }
                }
            }
        }
    }
    // ---- grid-wide sums of this level
    if constexpr (LV <= 3) {
        __shared__ double sh[4][4];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        sA1 = warp_sum(sA1); sA2 = warp_sum(sA2); sB1 = warp_sum(sB1); sB2 = warp_sum(sB2);
        if (lane == 0) { sh[0][warp] = sA1; sh[1][warp] = sA2; sh[2][warp] = sB1; sh[3][warp] = sB2; }
        __syncthreads();
        if (threadIdx.x < 4) {
            double v = sh[threadIdx.x][0] + sh[threadIdx.x][1] + sh[threadIdx.x][2] + sh[threadIdx.x][3];
            // LV 1 -> sets (3,4); LV 2 -> sets (1,2); LV 3 -> set 0
            const int setA = (LV == 1) ? 3 : (LV == 2 ? 1 : 0), setB = (LV == 1) ? 4 : 2;
            const int set = (threadIdx.x < 2) ? setA : setB;
            if (LV != 3 || threadIdx.x < 2) atomicAdd(hb.bsum + 2 * set + (threadIdx.x & 1), v);
        }
        if constexpr (LV == 1) {
            __syncthreads();
            if (threadIdx.x < 34) atomicAdd(hb.psum + threadIdx.x, red[threadIdx.x]);
        }
    }
}

// psum (double) -> prediction-conv gradients (+=)
__global__ void head_pred_grads_kernel(const double* __restrict__ psum, urnn_head_grads g) {
    int i = threadIdx.x;
    if (i < 16) { g.cls_pred_w[i] += (float)psum[i]; g.reg_pred_w[i] += (float)psum[17 + i]; }
    if (i == 16) { g.cls_pred_b[0] += (float)psum[16]; g.reg_pred_b[0] += (float)psum[33]; }
}

}  // namespace urnn
