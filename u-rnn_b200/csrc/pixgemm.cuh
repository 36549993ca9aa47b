// pixgemm.cuh -- fp32 "pixel GEMM": out[M x pixels] = A[M x K] * B[K x pixels] with pluggable B-tile loaders
// (virtual channel concatenation, gating, pooling gather, k x k taps) and fused epilogues.
// This is the URNN_MATH_FP32 parity path (FFMA, fp32 accumulate); the tcgen05 path lives in cgru_tc.cu.
//
// Tiling: CTA = 256 threads = 8 (channel dim, ty) x 32 (pixel dim, tx); CTA tile = (8*TM) outputs x 128 pixels;
// thread tile = TM outputs x 4 consecutive pixels; K is consumed in slabs of 16 through double-buffered
// shared memory with register prefetch.  For TM % 4 == 0 a thread's outputs are 4 consecutive channels in each
// 32-channel block (ch = (i/4)*32 + ty*4 + i%4) so that one GroupNorm group (32 channels) is spread over the
// whole CTA and group statistics fall out of a CTA-wide reduction.
#pragma once
#include "urnn_common.cuh"

namespace urnn {

constexpr int PG_BN = 128;   // pixels per CTA tile
constexpr int PG_BK = 16;    // K slab

template <int TM>
__device__ __forceinline__ int pg_channel(int ty, int i) {
    if constexpr (TM % 4 == 0) return (i >> 2) * 32 + ty * 4 + (i & 3);
    else return ty * TM + i;
}

// A operand addressing: A(m,k) = A[m*sm + k*sk]
struct AView { const float* ptr; long sm; long sk; };

template <int TM, class BLoader, class Epilogue>
__global__ void __launch_bounds__(256)
pixgemm_kernel(AView A, int M, int K, int N, BLoader bl, Epilogue epi) {
    constexpr int BM = TM * 8;
    constexpr int AS = BM + 4;                       // padded row (keeps 16B alignment)
    constexpr int AR = (BM * PG_BK + 255) / 256;     // A elements per thread per slab
    __shared__ __align__(16) float As[2][PG_BK][AS];
    __shared__ __align__(16) float Bs[2][PG_BK][PG_BN];

    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int p0 = blockIdx.x * PG_BN;
    const int m0 = blockIdx.y * BM;

    bl.init(p0, N);   // per-CTA loader setup (may use shared memory + __syncthreads)

    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

    float4 breg[2];
    float  areg[AR];
    const int nk = (K + PG_BK - 1) / PG_BK;
    const bool a_kcontig = (A.sk == 1);

    auto gload = [&](int kt) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int idx = tid + i * 256;
            int kk = idx >> 5, q = idx & 31;
            int k = kt * PG_BK + kk;
            breg[i] = (k < K) ? bl.load4(k, p0 + q * 4, N) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < AR; ++i) {
            int idx = tid + i * 256;
            float v = 0.f;
            if (idx < BM * PG_BK) {
                int m, kk;
                if (a_kcontig) { m = idx / PG_BK; kk = idx % PG_BK; }
                else           { kk = idx / BM;   m = idx % BM; }
                int k = kt * PG_BK + kk;
                if (k < K && m0 + m < M) v = __ldg(A.ptr + (long)(m0 + m) * A.sm + (long)k * A.sk);
            }
            areg[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int idx = tid + i * 256;
            int kk = idx >> 5, q = idx & 31;
            *reinterpret_cast<float4*>(&Bs[buf][kk][q * 4]) = breg[i];
        }
#pragma unroll
        for (int i = 0; i < AR; ++i) {
            int idx = tid + i * 256;
            if (idx < BM * PG_BK) {
                int m, kk;
                if (a_kcontig) { m = idx / PG_BK; kk = idx % PG_BK; }
                else           { kk = idx / BM;   m = idx % BM; }
                As[buf][kk][m] = areg[i];
            }
        }
    };

    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
        for (int kk = 0; kk < PG_BK; ++kk) {
            const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
            if constexpr (TM % 4 == 0) {
#pragma unroll
                for (int j = 0; j < TM / 4; ++j) {
                    const float4 a = *reinterpret_cast<const float4*>(&As[cur][kk][j * 32 + ty * 4]);
                    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc[j * 4 + u][0] = fmaf(av[u], b.x, acc[j * 4 + u][0]);
                        acc[j * 4 + u][1] = fmaf(av[u], b.y, acc[j * 4 + u][1]);
                        acc[j * 4 + u][2] = fmaf(av[u], b.z, acc[j * 4 + u][2]);
                        acc[j * 4 + u][3] = fmaf(av[u], b.w, acc[j * 4 + u][3]);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const float a = As[cur][kk][ty * TM + i];
                    acc[i][0] = fmaf(a, b.x, acc[i][0]);
                    acc[i][1] = fmaf(a, b.y, acc[i][1]);
                    acc[i][2] = fmaf(a, b.z, acc[i][2]);
                    acc[i][3] = fmaf(a, b.w, acc[i][3]);
                }
            }
        }
        if (kt + 1 < nk) {
            sstore(cur ^ 1);
            __syncthreads();
        }
    }
    epi.template run<TM>(acc, m0, M, ty, tx, p0 + tx * 4, N);
}

// ------------------------------------------------------------------------------------------------ B loaders
// load4(k, p, N): the 4 values of reduction row k at tile units p..p+3 (p % 4 == 0), zeros past N.

// k = 1: channel concatenation of up to three NCHW sources; optional reset gate on the last segment:
// value = src * sigmoid(G[gch0 + c] * scale[gch0 + c] + shift[gch0 + c]).
struct SegLoader {
    const float* src[3]; int cnt[3];        // cumulative ends in cend
    int cend[3];
    long plane;                             // elements per channel plane (= N)
    const float* gate_pre; const float* gate_scale; const float* gate_shift; int gate_ch0;  // nullptr => no gate
    bool vec;                               // plane % 4 == 0 and all bases 16-byte aligned: float4 path
    __device__ __forceinline__ void init(int, int) {}
    __device__ __forceinline__ float4 load4(int k, int p, int N) const {
        if (p >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
        int seg = (k < cend[0]) ? 0 : ((k < cend[1]) ? 1 : 2);
        int c = k - (seg == 0 ? 0 : cend[seg - 1]);
        const float* sp = src[seg] + (long)c * plane + p;
        float4 v;
        if (vec) {
            v = __ldg(reinterpret_cast<const float4*>(sp));
        } else {
            v.x = __ldg(sp);
            v.y = (p + 1 < N) ? __ldg(sp + 1) : 0.f;
            v.z = (p + 2 < N) ? __ldg(sp + 2) : 0.f;
            v.w = (p + 3 < N) ? __ldg(sp + 3) : 0.f;
        }
        bool last = (seg == 2) || (seg == 1 && cnt[2] == 0);
        if (gate_pre != nullptr && last) {
            int gc = gate_ch0 + c;
            const float* gp = gate_pre + (long)gc * plane + p;
            float4 g;
            if (vec) {
                g = __ldg(reinterpret_cast<const float4*>(gp));
            } else {
                g.x = __ldg(gp);
                g.y = (p + 1 < N) ? __ldg(gp + 1) : 0.f;
                g.z = (p + 2 < N) ? __ldg(gp + 2) : 0.f;
                g.w = (p + 3 < N) ? __ldg(gp + 3) : 0.f;
            }
            float sc = __ldg(gate_scale + gc), sh = __ldg(gate_shift + gc);
            v.x *= sigmoid_acc(fmaf(g.x, sc, sh));
            v.y *= sigmoid_acc(fmaf(g.y, sc, sh));
            v.z *= sigmoid_acc(fmaf(g.z, sc, sh));
            v.w *= sigmoid_acc(fmaf(g.w, sc, sh));
        }
        return v;
    }
};

// k x k taps (odd k, zero padding (k-1)/2): reduction index = (c_concat * k + dy) * k + dx, matching the
// nn.Conv2d weight layout (Cout, Cin, k, k).  Same segment/gate semantics as SegLoader.
struct TapLoader {
    const float* src[3]; int cnt[3]; int cend[3];
    long plane; int H, W, ks;
    const float* gate_pre; const float* gate_scale; const float* gate_shift; int gate_ch0;
    bool vec;   // unused (taps are gathered element-wise)
    __device__ __forceinline__ void init(int, int) {}
    __device__ __forceinline__ float4 load4(int k, int p, int N) const {
        float out[4] = {0.f, 0.f, 0.f, 0.f};
        if (p >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
        int kk = ks * ks;
        int cc = k / kk, tap = k % kk, dy = tap / ks - (ks - 1) / 2, dx = tap % ks - (ks - 1) / 2;
        int seg = (cc < cend[0]) ? 0 : ((cc < cend[1]) ? 1 : 2);
        int c = cc - (seg == 0 ? 0 : cend[seg - 1]);
        bool last = (seg == 2) || (seg == 1 && cnt[2] == 0);
        bool gated = gate_pre != nullptr && last;
        const float* base = src[seg] + (long)c * plane;
        float sc = 0.f, sh = 0.f; const float* gbase = nullptr;
        if (gated) {
            int gc = gate_ch0 + c;
            sc = __ldg(gate_scale + gc); sh = __ldg(gate_shift + gc);
            gbase = gate_pre + (long)gc * plane;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int q = p + u;
            if (q >= N) break;
            int yy = q / W + dy, xx = q % W + dx;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                long o = (long)yy * W + xx;
                float v = __ldg(base + o);
                if (gated) v *= sigmoid_acc(fmaf(__ldg(gbase + o), sc, sh));
                out[u] = v;
            }
        }
        return make_float4(out[0], out[1], out[2], out[3]);
    }
};

// 2x2 pooling gather: tile unit q = 4*Q + s addresses sub-pixel s of the 2x2 quad Q of a (C,H,W) source,
// so that a thread's 4 "pixels" are exactly one pooling window.  N = 4 * (H/2) * (W/2).
struct QuadLoader {
    const float* src; long plane; int W, W2;
    __device__ __forceinline__ void init(int, int) {}
    __device__ __forceinline__ float4 load4(int k, int p, int N) const {
        if (p >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
        int Q = p >> 2;
        int qy = Q / W2, qx = Q % W2;
        const float* b = src + (long)k * plane + (long)(2 * qy) * W + 2 * qx;
        float2 t = __ldg(reinterpret_cast<const float2*>(b));
        float2 u = __ldg(reinterpret_cast<const float2*>(b + W));
        return make_float4(t.x, t.y, u.x, u.y);
    }
};

// ------------------------------------------------------------------------------------------------ epilogues
// out[ch][pixel] = acc + bias[ch]; per-32-channel-group (sum, sumsq) partials -> StatSink (GroupNorm).
struct GnStatsEpilogue {
    static constexpr bool kAllowSmallTM = false;
    const float* bias; float* out; long plane; bool vec;
    StatSink sink; AffineOut aff;
    template <int TM>
    __device__ __forceinline__ void run(float (&acc)[TM][4], int m0, int M, int ty, int tx, int p, int N) {
        static_assert(TM % 4 == 0, "GroupNorm epilogue needs 32-channel blocks");
        constexpr int NG = TM / 4;
        __shared__ float red[2][NG][8];
        float s[NG], ss[NG];
#pragma unroll
        for (int j = 0; j < NG; ++j) { s[j] = 0.f; ss[j] = 0.f; }
        const int nvalid = (p >= N) ? 0 : ((N - p >= 4) ? 4 : (N - p));
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int ch = m0 + pg_channel<TM>(ty, i);
            float b = __ldg(bias + ch);
            float v[4] = {acc[i][0] + b, acc[i][1] + b, acc[i][2] + b, acc[i][3] + b};
            float* o = out + (long)ch * plane + p;
            if (vec && nvalid == 4) {
                *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) if (u < nvalid) o[u] = v[u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (u < nvalid) { s[i >> 2] += v[u]; ss[i >> 2] = fmaf(v[u], v[u], ss[i >> 2]); }
        }
#pragma unroll
        for (int j = 0; j < NG; ++j) {
            float a = warp_sum(s[j]), b = warp_sum(ss[j]);
            if (tx == 0) { red[0][j][ty] = a; red[1][j][ty] = b; }
        }
        __syncthreads();
        if (threadIdx.x < NG) {
            int j = threadIdx.x;
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { a += red[0][j][w]; b += red[1][j][w]; }
            int set = blockIdx.y * NG + j;
            sink.partial[(size_t)set * sink.stride + blockIdx.x] = make_float2(a, b);
        }
        stats_finalize_last_cta(sink, gridDim.x, gridDim.x * gridDim.y, &aff);
    }
};

// y = lrelu(acc + bias), stored NCHW.
struct LreluEpilogue {
    static constexpr bool kAllowSmallTM = true;
    const float* bias; float* out; long plane; float slope; bool vec;
    template <int TM>
    __device__ __forceinline__ void run(float (&acc)[TM][4], int m0, int M, int ty, int tx, int p, int N) {
        if (p >= N) return;
        const int nvalid = (N - p >= 4) ? 4 : (N - p);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int ch = m0 + pg_channel<TM>(ty, i);
            if (ch >= M) continue;
            float b = __ldg(bias + ch);
            float v[4] = {lrelu(acc[i][0] + b, slope), lrelu(acc[i][1] + b, slope),
                          lrelu(acc[i][2] + b, slope), lrelu(acc[i][3] + b, slope)};
            float* o = out + (long)ch * plane + p;
            if (vec && nvalid == 4) {
                *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) if (u < nvalid) o[u] = v[u];
            }
        }
    }
};

// y[ch][Q] = mean over the quad of lrelu(acc + bias)   (used with QuadLoader; plane = (H/2)*(W/2)).
struct LreluPoolEpilogue {
    static constexpr bool kAllowSmallTM = true;
    const float* bias; float* out; long plane; float slope;
    template <int TM>
    __device__ __forceinline__ void run(float (&acc)[TM][4], int m0, int M, int ty, int tx, int p, int N) {
        if (p >= N) return;
        int Q = p >> 2;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int ch = m0 + pg_channel<TM>(ty, i);
            if (ch >= M) continue;
            float b = __ldg(bias + ch);
            float v = (lrelu(acc[i][0] + b, slope) + lrelu(acc[i][1] + b, slope)) +
                      (lrelu(acc[i][2] + b, slope) + lrelu(acc[i][3] + b, slope));
            out[(long)ch * plane + Q] = 0.25f * v;
        }
    }
};

// ConvTranspose2d(k=2,s=2) scatter: GEMM row m = co*4 + dy*2 + dx; y[co][2y+dy][2x+dx] = lrelu(acc + bias[co]).
struct DeconvEpilogue {
    static constexpr bool kAllowSmallTM = false;
    const float* bias; float* out; int W; float slope;   // W = input width; output plane = 4*H*W
    long oplane; bool vec;                               // vec: W % 4 == 0 (a thread's 4 pixels share a row)
    template <int TM>
    __device__ __forceinline__ void run(float (&acc)[TM][4], int m0, int M, int ty, int tx, int p, int N) {
        static_assert(TM % 4 == 0, "deconv epilogue needs 4 consecutive rows per thread");
        if (p >= N) return;
        const int nvalid = (N - p >= 4) ? 4 : (N - p);
        int y = p / W, x0 = p % W;
#pragma unroll
        for (int j = 0; j < TM / 4; ++j) {
            int m = m0 + pg_channel<TM>(ty, j * 4);
            if (m >= M) continue;
            int co = m >> 2;
            float b = __ldg(bias + co);
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const float* e = acc[j * 4 + dy * 2];       // dx = 0
                const float* f = acc[j * 4 + dy * 2 + 1];   // dx = 1
                if (vec) {
                    float* o = out + (long)co * oplane + (long)(2 * y + dy) * (2 * W) + 2 * x0;
                    *reinterpret_cast<float4*>(o) = make_float4(lrelu(e[0] + b, slope), lrelu(f[0] + b, slope),
                                                                lrelu(e[1] + b, slope), lrelu(f[1] + b, slope));
                    *reinterpret_cast<float4*>(o + 4) = make_float4(lrelu(e[2] + b, slope), lrelu(f[2] + b, slope),
                                                                    lrelu(e[3] + b, slope), lrelu(f[3] + b, slope));
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (u >= nvalid) break;
                        int q = p + u, yy = q / W, xx = q % W;
                        float* o = out + (long)co * oplane + (long)(2 * yy + dy) * (2 * W) + 2 * xx;
                        *reinterpret_cast<float2*>(o) = make_float2(lrelu(e[u] + b, slope), lrelu(f[u] + b, slope));
                    }
                }
            }
        }
    }
};

// out[row][pixel] = acc, rows routed by ranges to up to three destination maps (store / accumulate / skip): the
// transposed-weight products of the backward pass (d input = W^T d output) land directly in dx / de / dh.
struct RouteEpilogue {
    static constexpr bool kAllowSmallTM = true;
    float* dst[3]; int mend[3]; int mode[3];   // mode: 0 skip, 1 store, 2 accumulate
    long plane;
    template <int TM>
    __device__ __forceinline__ void run(float (&acc)[TM][4], int m0, int M, int ty, int tx, int p, int N) {
        if (p >= N) return;
        const int nvalid = (N - p >= 4) ? 4 : (N - p);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int ch = m0 + pg_channel<TM>(ty, i);
            if (ch >= M) continue;
            int sg = (ch < mend[0]) ? 0 : ((ch < mend[1]) ? 1 : 2);
            if (mode[sg] == 0) continue;
            int c = ch - (sg == 0 ? 0 : mend[sg - 1]);
            float* o = dst[sg] + (long)c * plane + p;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (u < nvalid) { if (mode[sg] == 1) o[u] = acc[i][u]; else o[u] += acc[i][u]; }
        }
    }
};

// Backward of the stems: pre = acc + bias is the recomputed pre-activation; out = dL/dpre = dy(gathered) * LeakyReLU'(pre).
// kind 0: plain conv (dy[ch][p]); 1: conv + AvgPool2 (dy[ch][quad(p)] / 4); 2: 2x2 transposed conv, GEMM row
// m = co*4 + dy*2 + dx of input pixel (y,x) reads dy_out[co][2y+dy][2x+dx].
struct DpreEpilogue {
    static constexpr bool kAllowSmallTM = true;
    const float* bias; int bias_shift; const float* dy; float* out; long plane; int kind; int W; long dy_plane; float slope;
    template <int TM>
    __device__ __forceinline__ void run(float (&acc)[TM][4], int m0, int M, int ty, int tx, int p, int N) {
        if (p >= N) return;
        const int nvalid = (N - p >= 4) ? 4 : (N - p);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int m = m0 + pg_channel<TM>(ty, i);
            if (m >= M) continue;
            float b = __ldg(bias + (m >> bias_shift));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u >= nvalid) break;
                int q = p + u, y = q / W, x = q % W;
                float g;
                if (kind == 0)      g = __ldg(dy + (long)m * dy_plane + q);
                else if (kind == 1) g = 0.25f * __ldg(dy + (long)m * dy_plane + (long)(y >> 1) * (W >> 1) + (x >> 1));
                else                g = __ldg(dy + (long)(m >> 2) * dy_plane + (long)(2 * y + ((m >> 1) & 1)) * (2 * W) + 2 * x + (m & 1));
                float pre = acc[i][u] + b;
                out[(long)m * plane + q] = (pre > 0.f) ? g : g * slope;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ launch helpers
template <class BL, class EP>
static inline int launch_pixgemm(AView A, int M, int K, int N, const BL& bl, const EP& ep, bool need_gn, cudaStream_t st) {
    int tm;
    if (need_gn) {
        if (M % 128 == 0) tm = 16; else if (M % 96 == 0) tm = 12; else if (M % 64 == 0) tm = 8; else tm = 4;
        if (M % (8 * tm) != 0) { set_error("pixgemm: M=%d is not a multiple of 32", M); return URNN_E_INVALID; }
    } else {
        tm = (M <= 16) ? 2 : (M <= 32) ? 4 : (M <= 64) ? 8 : (M <= 96 || M % 96 == 0) ? 12 : 16;
    }
    dim3 grid((N + PG_BN - 1) / PG_BN, (M + 8 * tm - 1) / (8 * tm));
    switch (tm) {
        case 16: pixgemm_kernel<16, BL, EP><<<grid, 256, 0, st>>>(A, M, K, N, bl, ep); break;
        case 12: pixgemm_kernel<12, BL, EP><<<grid, 256, 0, st>>>(A, M, K, N, bl, ep); break;
        case 8:  pixgemm_kernel<8, BL, EP><<<grid, 256, 0, st>>>(A, M, K, N, bl, ep); break;
        case 4:  pixgemm_kernel<4, BL, EP><<<grid, 256, 0, st>>>(A, M, K, N, bl, ep); break;
        default:
            if constexpr (EP::kAllowSmallTM) { pixgemm_kernel<2, BL, EP><<<grid, 256, 0, st>>>(A, M, K, N, bl, ep); break; }
            else { set_error("pixgemm: unsupported tile"); return URNN_E_INVALID; }
    }
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}


// ------------------------------------------------------------------------------------------------ weight gradients
// dW[m][k] += sum_p A[m][p] * B[k][p]   (A: a plain [M][N] map of output gradients, B: any pixel-GEMM loader)
// db[m >> db_shift] += sum_p A[m][p].   Pixels are split over blockIdx.x; partial tiles are combined with atomics
// (fp32 atomics: summation order is not fixed, like cuDNN's weight-gradient kernels).
template <class BLoader>
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ A, long a_plane, int M, int K, int N, BLoader bl, float* dW, long w_sm, long w_sk,
             float* db, int db_shift, int pix_per_cta) {
    __shared__ __align__(16) float As[32][64 + 4];
    __shared__ __align__(16) float Bs[32][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, k0 = blockIdx.z * 64;
    const int pbeg = blockIdx.x * pix_per_cta;
    const int pend = (pbeg + pix_per_cta < N) ? (pbeg + pix_per_cta) : N;
    float acc[4][4];
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    bl.init(pbeg, N);
    for (int ps = pbeg; ps < pend; ps += 32) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int idx = tid + i * 256;               // 512 float4 slots: 64 rows x 8 pixel quads
            int row = idx >> 3, q = idx & 7, p = ps + q * 4;
            float a4[4] = {0.f, 0.f, 0.f, 0.f};
            if (m0 + row < M) {
                const float* src = A + (long)(m0 + row) * a_plane + p;
#pragma unroll
                for (int u = 0; u < 4; ++u) if (p + u < pend) a4[u] = __ldg(src + u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) As[q * 4 + u][row] = a4[u];
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + row < K && p < pend) b4 = bl.load4(k0 + row, p, N);
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) Bs[q * 4 + u][row] = (p + u < pend) ? bv[u] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int px = 0; px < 32; ++px) {
            const float4 a = *reinterpret_cast<const float4*>(&As[px][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[px][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                rs[i] += av[i];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k < K) atomicAdd(dW + (long)m * w_sm + (long)k * w_sk, acc[i][j]);
        }
        if (db != nullptr && blockIdx.z == 0 && tx == 0) atomicAdd(db + (m >> db_shift), rs[i]);
    }
}

template <class BL>
static inline int launch_wgrad(const float* A, long a_plane, int M, int K, int N, const BL& bl, float* dW, long w_sm, long w_sk,
                               float* db, int db_shift, cudaStream_t st) {
    const int pix = 2048;
    dim3 grid((N + pix - 1) / pix, (M + 63) / 64, (K + 63) / 64);
    wgrad_kernel<BL><<<grid, 256, 0, st>>>(A, a_plane, M, K, N, bl, dW, w_sm, w_sk, db, db_shift, pix);
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}

// single-map loader helper
static inline SegLoader single_map_loader(const float* x, int C, long plane) {
    SegLoader L;
    L.src[0] = L.src[1] = L.src[2] = x; L.cnt[0] = C; L.cnt[1] = L.cnt[2] = 0;
    L.cend[0] = L.cend[1] = L.cend[2] = C; L.plane = plane;
    L.gate_pre = nullptr; L.gate_scale = nullptr; L.gate_shift = nullptr; L.gate_ch0 = 0;
    L.vec = (plane % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    return L;
}

}  // namespace urnn
