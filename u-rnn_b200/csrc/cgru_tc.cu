// cgru_tc.cu -- URNN_MATH_BF16 path of the (Skip-)ConvGRU cell: the two gate contractions run on tcgen05
// (tc_pixgemm.cuh), GroupNorm statistics are fused into their epilogues, the gate/blend sweep is shared with
// the fp32 path.  k = 1 only (the production configuration); other filter sizes use the fp32 path.
#include "tc_pixgemm.cuh"
#include "urnn_internal.h"
#include <stdlib.h>
#define TRACE(...) do { if (getenv("URNN_TRACE")) { fprintf(stderr, __VA_ARGS__); fflush(stderr); } } while (0)

namespace urnn {

static int g_num_sms = 0;
static bool g_attr_set[2] = {false, false};

static int tc_launch(tc::GemmParams& P, cudaStream_t st) {
    if (g_num_sms == 0) {
        int dev = 0;
        URNN_CUDA(cudaGetDevice(&dev));
        URNN_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const bool gated = P.seg.gate_seg >= 0;
    if (!g_attr_set[gated]) {
        if (gated) URNN_CUDA(cudaFuncSetAttribute(tc::gemm_gn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_CAP));
        else       URNN_CUDA(cudaFuncSetAttribute(tc::gemm_gn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_CAP));
        g_attr_set[gated] = true;
    }
    TRACE("tc_launch: NOUT=%d K=%d N=%d sms=%d\n", P.NOUT, P.K, P.N, g_num_sms);
    int nstage = 0;
    size_t smem = tc::gemm_smem_bytes(P.NOUT, P.K, &nstage);
    if (smem == 0) { set_error("tc gemm: weights %dx%d do not fit in shared memory", P.NOUT, P.K); return URNN_E_UNSUPPORTED; }
    P.nstage = nstage;
    int cols = 32;
    while (cols < 2 * P.NOUT) cols <<= 1;
    P.tmem_cols = cols;
    int ntiles = (P.N + tc::TILE_M - 1) / tc::TILE_M;
    int grid = ntiles < g_num_sms ? ntiles : g_num_sms;
    TRACE("tc_launch: grid=%d smem=%zu nstage=%d cols=%d\n", grid, smem, nstage, cols);
    if (gated) tc::gemm_gn_kernel<true><<<grid, tc::NTHREADS, smem, st>>>(P);
    else       tc::gemm_gn_kernel<false><<<grid, tc::NTHREADS, smem, st>>>(P);
    URNN_LAUNCH_CHECK();
    TRACE("tc_launch: launched\n");
    return URNN_OK;
}

int cgru_fwd_bf16(const urnn_cell_desc* d, const urnn_cell_params* p, const float* x, const float* e,
                  const float* h, float* h_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (d->ksize != 1) return cgru_fwd_fp32(d, p, x, e, h, h_out, ws, ws_bytes, st);   // k>1: fp32 taps path
    const int F = d->F;
    const long N = (long)d->H * d->W;
    const int Ch = (d->variant == URNN_CELL_DECODER) ? 2 * F : F;
    const int Ktot = d->Cx + Ch;
    const int Cx_eff = x ? d->Cx : 0;
    const int Keff = Cx_eff + Ch;
    const long aoff = x ? 0 : d->Cx;                    // skip the zero-input weight columns (ConvRNN.py:143-146)
    if (2 * F > 256) { set_error("cgru_fwd(bf16): num_features=%d > 128 not supported by the tcgen05 tile", F); return URNN_E_UNSUPPORTED; }
    CellWsView w;
    size_t need = cell_ws_view(d, ws, ws_bytes, &w);
    if (need > ws_bytes) { set_error("cgru_fwd: workspace %zu < %zu bytes", ws_bytes, need); return URNN_E_WORKSPACE; }
    TRACE("cgru_fwd_bf16: F=%d N=%ld Keff=%d\n", F, N, Keff);
    URNN_CUDA(cudaMemsetAsync(w.counter, 0, 64 * sizeof(unsigned), st));
    TRACE("cgru_fwd_bf16: memset ok\n");

    tc::GemmParams P;
    P.dbg = nullptr;
    // segments [x | e | h]; missing ones get zero width
    int n = 0; const float* srcs[3] = {h, h, h}; int cnt[3] = {0, 0, 0};
    if (x) { srcs[n] = x; cnt[n] = d->Cx; ++n; }
    if (d->variant == URNN_CELL_DECODER) { srcs[n] = e; cnt[n] = F; ++n; }
    srcs[n] = h; cnt[n] = F; const int hseg = n; ++n;
    int acc = 0;
    for (int i = 0; i < 3; ++i) { P.seg.src[i] = srcs[i]; acc += cnt[i]; P.seg.cend[i] = acc; }
    P.seg.plane = N;
    P.seg.vec = (N % 4 == 0) && ((((uintptr_t)x | (uintptr_t)e | (uintptr_t)h | (uintptr_t)w.G) & 15) == 0);
    P.seg.gate_seg = -1; P.seg.gate_ch0 = 0; P.seg.gate_pre = nullptr; P.seg.gate_scale = nullptr; P.seg.gate_shift = nullptr;
    P.N = (int)N; P.K = Keff; P.w_ld = Ktot;

    // pass A: G = W1 [x|e|h] + b1, GroupNorm-1 statistics
    P.W = p->w1 + aoff; P.bias = p->b1; P.NOUT = 2 * F; P.out = w.G; P.out_plane = N;
    P.sink = StatSink{w.partial1, w.total1, w.counter, 2 * F / 32, w.gx};
    P.aff = AffineOut{w.scale1, w.shift1, p->gn1_w, p->gn1_b, 2 * F, 32, 32.0 * (double)N, d->eps};
    URNN_TRY(tc_launch(P, st));
    // pass B: C = W2 [x|e|r*h] + b2 with r = sigmoid(GN1(G)[F:]), GroupNorm-2 statistics
    P.seg.gate_seg = hseg; P.seg.gate_ch0 = F; P.seg.gate_pre = w.G; P.seg.gate_scale = w.scale1; P.seg.gate_shift = w.shift1;
    P.W = p->w2 + aoff; P.bias = p->b2; P.NOUT = F; P.out = w.C;
    P.sink = StatSink{w.partial2, w.total2, w.counter + 1, F / 32, w.gx};
    P.aff = AffineOut{w.scale2, w.shift2, p->gn2_w, p->gn2_b, F, 32, 32.0 * (double)N, d->eps};
    URNN_TRY(tc_launch(P, st));
    // pass C: gates + blend (shared with the fp32 path)
    return cgru_blend_launch(w, h, h_out, F, N, st);
}

}  // namespace urnn
