// urnn_fp32.cu -- URNN_MATH_FP32 forward path: ConvGRU cell passes, stage stems, head, on the fp32 pixel-GEMM.
// Host-side orchestration only enqueues kernels on the caller's stream (graph-capturable).
#include "pixgemm.cuh"
#include <stdlib.h>
#include "head.cuh"
#include "urnn_internal.h"

namespace urnn {

// ------------------------------------------------------------------------------------------------ cell, fp32
// h_out = (1-z)*h + z*tanh(GN2(C)),  z = sigmoid(GN1(G)[:F])           (ConvRNN.py:160-162,180,185,189)
__global__ void __launch_bounds__(256)
cgru_blend_kernel(const float* __restrict__ G, const float* __restrict__ C, const float* __restrict__ h,
                  const float* __restrict__ sc1, const float* __restrict__ sh1,
                  const float* __restrict__ sc2, const float* __restrict__ sh2,
                  float* __restrict__ h_out, long N, long nquad) {
    // flat over (channel, pixel) in quads; F*N is always a multiple of 4 (F % 32 == 0) but a quad may
    // straddle two channels when N % 4 != 0
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nquad) return;
    float4 g = __ldg(reinterpret_cast<const float4*>(G) + idx);
    float4 cc = __ldg(reinterpret_cast<const float4*>(C) + idx);
    float4 hv = __ldg(reinterpret_cast<const float4*>(h) + idx);
    const float gv[4] = {g.x, g.y, g.z, g.w}, cv[4] = {cc.x, cc.y, cc.z, cc.w}, hh[4] = {hv.x, hv.y, hv.z, hv.w};
    float o[4];
    int c0 = (int)((idx * 4) / N), c3 = (int)((idx * 4 + 3) / N);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int c = (c0 == c3) ? c0 : (int)((idx * 4 + u) / N);
        float a1 = __ldg(sc1 + c), b1 = __ldg(sh1 + c), a2 = __ldg(sc2 + c), b2 = __ldg(sh2 + c);
        float z = sigmoid_acc(fmaf(gv[u], a1, b1)), t = tanhf(fmaf(cv[u], a2, b2));
        o[u] = (1.f - z) * hh[u] + z * t;
    }
    reinterpret_cast<float4*>(h_out)[idx] = make_float4(o[0], o[1], o[2], o[3]);
}

typedef CellWsView CellWs;

static size_t cell_ws_layout(const urnn_cell_desc* d, void* ws, size_t ws_bytes, CellWs* out) {
    long N = (long)d->H * d->W;
    int F = d->F, gx = (int)((N + PG_BN - 1) / PG_BN);
    Arena a(ws, ws_bytes);
    CellWs w;
    w.gx = gx;
    w.counter = a.take<unsigned>(64);
    w.total1 = a.take<double2>(2 * F / 32);
    w.total2 = a.take<double2>(F / 32);
    w.scale1 = a.take<float>(2 * F); w.shift1 = a.take<float>(2 * F);
    w.scale2 = a.take<float>(F);     w.shift2 = a.take<float>(F);
    w.partial1 = a.take<float2>((size_t)(2 * F / 32) * gx);
    w.partial2 = a.take<float2>((size_t)(F / 32) * gx);
    w.G = a.take<float>((size_t)2 * F * N);
    w.C = a.take<float>((size_t)F * N);
    if (out) *out = w;
    return align_up(a.off, 256);
}

size_t cgru_fwd_fp32_workspace(const urnn_cell_desc* d) { return cell_ws_layout(d, nullptr, 0, nullptr); }
size_t cell_ws_view(const urnn_cell_desc* d, void* ws, size_t ws_bytes, CellWsView* out) { return cell_ws_layout(d, ws, ws_bytes, out); }

int cgru_blend_launch(const CellWsView& w, const float* h, float* h_out, int F, long N, cudaStream_t st) {
    long nquad = (long)F * N / 4;
    cgru_blend_kernel<<<(unsigned)((nquad + 255) / 256), 256, 0, st>>>(w.G, w.C, h, w.scale1, w.shift1, w.scale2,
                                                                       w.shift2, h_out, N, nquad);
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <class Loader>
static void fill_segments(Loader& L, const urnn_cell_desc* d, const float* x, const float* e, const float* h) {
    int n = 0, acc = 0;
    const float* srcs[3] = {nullptr, nullptr, nullptr}; int cnts[3] = {0, 0, 0};
    if (x != nullptr) { srcs[n] = x; cnts[n] = d->Cx; ++n; }
    if (d->variant == URNN_CELL_DECODER) { srcs[n] = e; cnts[n] = d->F; ++n; }
    srcs[n] = h; cnts[n] = d->F; ++n;
    // the gated (last) segment must be index 2, or index 1 when only two segments exist
    if (n == 1) { srcs[1] = srcs[0]; cnts[1] = cnts[0]; srcs[0] = h; cnts[0] = 0; n = 2; }
    for (int i = 0; i < 3; ++i) {
        L.src[i] = (i < n) ? srcs[i] : srcs[n - 1];
        L.cnt[i] = (i < n) ? cnts[i] : 0;
        acc += L.cnt[i];
        L.cend[i] = acc;
    }
    L.plane = (long)d->H * d->W;
    L.gate_pre = nullptr; L.gate_scale = nullptr; L.gate_shift = nullptr; L.gate_ch0 = 0;
    L.vec = (L.plane % 4 == 0) && aligned16(x) && aligned16(e) && aligned16(h);
}

// sweeps A and B only (pre-GN maps G, C and the folded GroupNorm affines stay in the workspace view `w`)
int cgru_fwd_fp32_passes(const urnn_cell_desc* d, const urnn_cell_params* p, const float* x, const float* e,
                         const float* h, CellWsView* wout, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int F = d->F, ks = d->ksize, kk = ks * ks;
    const long N = (long)d->H * d->W;
    const int Ch = (d->variant == URNN_CELL_DECODER) ? 2 * F : F;
    const int Ktot = d->Cx + Ch;                       // weight columns (x columns present even if x == NULL)
    const int Keff = (x ? d->Cx : 0) + Ch;
    const long aoff = x ? 0 : (long)d->Cx * kk;        // skip the zero-input columns (ConvRNN.py:143-146)
    CellWs w;
    size_t need = cell_ws_layout(d, ws, ws_bytes, &w);
    if (need > ws_bytes) { set_error("cgru_fwd: workspace %zu < %zu bytes", ws_bytes, need); return URNN_E_WORKSPACE; }
    URNN_CUDA(cudaMemsetAsync(w.counter, 0, 64 * sizeof(unsigned), st));

    GnStatsEpilogue ep1;
    ep1.bias = p->b1; ep1.out = w.G; ep1.plane = N; ep1.vec = (N % 4 == 0);
    CommDev comm; current_comm(&comm);
    if (comm.world > 1 && 2 * F / 32 > COMM_MAX_SETS) { set_error("cgru_fwd: num_features=%d > 128 cannot be sharded (%d GroupNorm groups per statistics exchange)", F, COMM_MAX_SETS); return URNN_E_UNSUPPORTED; }
    const double gcount = 32.0 * (double)N * (double)(comm.world > 1 ? comm.world : 1);
    ep1.sink = StatSink{w.partial1, w.total1, w.counter, 2 * F / 32, w.gx, comm};
    ep1.aff = AffineOut{w.scale1, w.shift1, p->gn1_w, p->gn1_b, 2 * F, 32, gcount, d->eps};
    GnStatsEpilogue ep2;
    ep2.bias = p->b2; ep2.out = w.C; ep2.plane = N; ep2.vec = (N % 4 == 0);
    ep2.sink = StatSink{w.partial2, w.total2, w.counter + 1, F / 32, w.gx, comm};
    ep2.aff = AffineOut{w.scale2, w.shift2, p->gn2_w, p->gn2_b, F, 32, gcount, d->eps};
    AView A1{p->w1 + aoff, (long)Ktot * kk, 1};
    AView A2{p->w2 + aoff, (long)Ktot * kk, 1};

    if (ks == 1) {
        SegLoader L; fill_segments(L, d, x, e, h);
        URNN_TRY(launch_pixgemm(A1, 2 * F, Keff, (int)N, L, ep1, true, st));
        L.gate_pre = w.G; L.gate_scale = w.scale1; L.gate_shift = w.shift1; L.gate_ch0 = F;
        URNN_TRY(launch_pixgemm(A2, F, Keff, (int)N, L, ep2, true, st));
    } else {
        TapLoader L; fill_segments(L, d, x, e, h);
        L.H = d->H; L.W = d->W; L.ks = ks;
        URNN_TRY(launch_pixgemm(A1, 2 * F, Keff * kk, (int)N, L, ep1, true, st));
        L.gate_pre = w.G; L.gate_scale = w.scale1; L.gate_shift = w.shift1; L.gate_ch0 = F;
        URNN_TRY(launch_pixgemm(A2, F, Keff * kk, (int)N, L, ep2, true, st));
    }
    *wout = w;
    return URNN_OK;
}

int cgru_fwd_fp32(const urnn_cell_desc* d, const urnn_cell_params* p, const float* x, const float* e,
                  const float* h, float* h_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    CellWsView w;
    URNN_TRY(cgru_fwd_fp32_passes(d, p, x, e, h, &w, ws, ws_bytes, st));
    return cgru_blend_launch(w, h, h_out, d->F, (long)d->H * d->W, st);
}

// ------------------------------------------------------------------------------------------------ stems
int conv1x1_lrelu_fwd_fp32(int Cin, int Cout, int H, int W, int pool, float slope, const float* x,
                           const float* w, long w_ld, const float* b, float* y, cudaStream_t st) {
    AView A{w, w_ld, 1};
    if (pool == 1) {
        SegLoader L;
        L.src[0] = L.src[1] = L.src[2] = x; L.cnt[0] = Cin; L.cnt[1] = L.cnt[2] = 0;
        L.cend[0] = L.cend[1] = L.cend[2] = Cin; L.plane = (long)H * W;
        L.gate_pre = nullptr; L.gate_scale = nullptr; L.gate_shift = nullptr; L.gate_ch0 = 0;
        L.vec = (L.plane % 4 == 0) && aligned16(x);
        LreluEpilogue ep{b, y, (long)H * W, slope, (L.plane % 4 == 0) && aligned16(y)};
        return launch_pixgemm(A, Cout, Cin, H * W, L, ep, false, st);
    }
    QuadLoader L{x, (long)H * W, W, W / 2};
    LreluPoolEpilogue ep{b, y, (long)(H / 2) * (W / 2), slope};
    return launch_pixgemm(A, Cout, Cin, 4 * (H / 2) * (W / 2), L, ep, false, st);
}

int deconv2x2_lrelu_fwd_fp32(int Cin, int Cout, int H, int W, float slope, const float* x, const float* w,
                             const float* b, float* y, cudaStream_t st) {
    AView A{w, 1, (long)Cout * 4};     // A(m,k) = w[k][m], m = co*4 + dy*2 + dx
    SegLoader L;
    L.src[0] = L.src[1] = L.src[2] = x; L.cnt[0] = Cin; L.cnt[1] = L.cnt[2] = 0;
    L.cend[0] = L.cend[1] = L.cend[2] = Cin; L.plane = (long)H * W;
    L.gate_pre = nullptr; L.gate_scale = nullptr; L.gate_shift = nullptr; L.gate_ch0 = 0;
    L.vec = (L.plane % 4 == 0) && aligned16(x);
    DeconvEpilogue ep{b, y, W, slope, (long)4 * H * W, (W % 4 == 0) && aligned16(y)};
    return launch_pixgemm(A, 4 * Cout, Cin, H * W, L, ep, false, st);
}

// ------------------------------------------------------------------------------------------------ head
struct HeadWs { float2* partial; double2* total; unsigned* counter; int gx; float *bufA, *bufB; };
static size_t head_ws_layout(int H, int W, void* ws, size_t ws_bytes, HeadWs* out) {
    long N = (long)H * W;
    int gx = (int)((N + 127) / 128);
    Arena a(ws, ws_bytes);
    HeadWs w; w.gx = gx;
    w.counter = a.take<unsigned>(64);
    w.total = a.take<double2>(8);
    w.partial = a.take<float2>((size_t)5 * gx);
    w.bufA = a.take<float>((size_t)16 * N);              // un-normalised cls / reg branch maps between the staged sweeps
    w.bufB = a.take<float>((size_t)16 * N);
    if (out) *out = w;
    return align_up(a.off, 256);
}
size_t head_fwd_fp32_workspace(int H, int W) { return head_ws_layout(H, W, nullptr, 0, nullptr); }

int head_fwd_fp32(int H, int W, float cls_thred, float ln_eps, float slope, const urnn_head_params* p,
                  const float* feat, float* out, void* ws, size_t ws_bytes, cudaStream_t st, int comm_lane) {
    HeadWs w;
    size_t need = head_ws_layout(H, W, ws, ws_bytes, &w);
    if (need > ws_bytes) { set_error("head_fwd: workspace %zu < %zu bytes", ws_bytes, need); return URNN_E_WORKSPACE; }
    URNN_CUDA(cudaMemsetAsync(w.counter, 0, 64 * sizeof(unsigned), st));
    long N = (long)H * W;
    HeadDev hd;
    hd.p = *p; hd.cls_thred = cls_thred; hd.eps = ln_eps; hd.slope = slope; hd.plane = N;
    CommDev comm; current_comm(&comm, comm_lane);
    hd.count = 16.0 * (double)N * (double)(comm.world > 1 ? comm.world : 1);
    hd.sink = StatSink{w.partial, w.total, w.counter, 5, w.gx, comm};
    head_kernel<0><<<w.gx, 128, 0, st>>>(hd, feat, out, (int)N); URNN_LAUNCH_CHECK();
    static const bool staged = !(getenv("URNN_HEAD_STAGED") && getenv("URNN_HEAD_STAGED")[0] == '0');
    if (staged) {
        // one LayerNorm level per sweep, starting from the previous level's stored pre-norm maps (head.cuh)
        static const int px = (getenv("URNN_HEAD_PX") && getenv("URNN_HEAD_PX")[0] == '1') ? 1 : 2;   // pixels per thread
        const int gx2 = (int)((N + 128 * px - 1) / (128 * px));
        if (px == 2) {
            head_stage_kernel<1, 2><<<gx2, 128, 0, st>>>(hd, feat, w.bufA, w.bufB, out, (int)N); URNN_LAUNCH_CHECK();
            head_stage_kernel<2, 2><<<gx2, 128, 0, st>>>(hd, feat, w.bufA, w.bufB, out, (int)N); URNN_LAUNCH_CHECK();
            head_stage_kernel<3, 2><<<gx2, 128, 0, st>>>(hd, feat, w.bufA, w.bufB, out, (int)N); URNN_LAUNCH_CHECK();
        } else {
            head_stage_kernel<1, 1><<<gx2, 128, 0, st>>>(hd, feat, w.bufA, w.bufB, out, (int)N); URNN_LAUNCH_CHECK();
            head_stage_kernel<2, 1><<<gx2, 128, 0, st>>>(hd, feat, w.bufA, w.bufB, out, (int)N); URNN_LAUNCH_CHECK();
            head_stage_kernel<3, 1><<<gx2, 128, 0, st>>>(hd, feat, w.bufA, w.bufB, out, (int)N); URNN_LAUNCH_CHECK();
        }
        return URNN_OK;
    }
    head_kernel<1><<<w.gx, 128, 0, st>>>(hd, feat, out, (int)N); URNN_LAUNCH_CHECK();
    head_kernel<2><<<w.gx, 128, 0, st>>>(hd, feat, out, (int)N); URNN_LAUNCH_CHECK();
    head_kernel<3><<<w.gx, 128, 0, st>>>(hd, feat, out, (int)N); URNN_LAUNCH_CHECK();
    return URNN_OK;
}

}  // namespace urnn
