// cgru_tc.cu -- URNN_MATH_BF16 path of the (Skip-)ConvGRU cell (k = 1; other filter sizes use the fp32 path).
//
// Three sweeps separated by the two GroupNorm statistics (ConvRNN.py:94-104 make a one-pass cell impossible):
//   A  tcgen05 GEMM  [x|e|h] -> G = W1.u + b1 (2F ch, GN-1 statistics fused)  and, when 3F <= 256,
//                               Pc = W2[:, x|e].[x|e] (F ch): the part of the candidate that does not depend on r
//   B  tcgen05 GEMM  r = sigmoid(GN1(G)[F:]) applied to h on the way into shared memory;
//                    C = Pc + W2[:, h].(r*h) + b2   (K = F only)          [3F <= 256]
//                    C = W2.[x|e|r*h] + b2          (K = Cx+Ch)           [otherwise]      (GN-2 statistics fused)
//   C  elementwise   z = sigmoid(GN1(G)[:F]); h' = (1-z) h + z tanh(GN2(C))
// G, Pc, C live in the workspace as bf16 maps (they are re-read once each); states stay fp32.
#include <stdlib.h>
#include "tc_pixgemm.cuh"
#include "urnn_internal.h"

namespace urnn {

// (no cached device properties: a process may drive several devices; the attribute query is a cached driver lookup)
// Sweep direction: consecutive kernels walk the grid in opposite directions so that a kernel starts on the data its
// predecessor touched last (still in the 126 MB L2).  Entry points reset it, so a given call sequence is reproducible.
static thread_local int g_dir = 0;
void tc_reset_direction() { g_dir = 0; }
static thread_local bool g_counters_clean = false;
// step driver: clear the cell workspace's ticket counters now and skip the per-cell memset until tc_counters_end()
int tc_counters_begin(const urnn_cell_desc* d, void* cell_ws, size_t ws_bytes, cudaStream_t st);
void tc_counters_end() { g_counters_clean = false; }
static inline long pad_plane(long n) { return (n + tc::TILE_M - 1) / tc::TILE_M * tc::TILE_M; }
long tc_pad_plane(long n) { return pad_plane(n); }

template <bool GATED, int EPI, bool BULK>
static int tc_launch_tb(tc::GemmParams& P, int grid, size_t smem, cudaStream_t st) {
    // per launch: the attribute is per device, and a process may use several (cheap driver call)
    URNN_CUDA(cudaFuncSetAttribute(tc::gemm_gn_kernel<GATED, EPI, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_CAP));
    // Programmatic dependent launch: the next kernel's CTAs may start their prologue (barrier init, TMEM allocation,
    // weight conversion) while the previous kernel drains; they touch its outputs only after griddepcontrol.wait.
    static const bool pdl = !(getenv("URNN_PDL") && getenv("URNN_PDL")[0] == '0');
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc::NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    URNN_CUDA(cudaLaunchKernelEx(&cfg, tc::gemm_gn_kernel<GATED, EPI, BULK>, (const tc::GemmParams)P));
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}
template <bool GATED, int EPI>
static int tc_launch_t(tc::GemmParams& P, int grid, size_t smem, cudaStream_t st) {
    return P.bulk ? tc_launch_tb<GATED, EPI, true>(P, grid, smem, st) : tc_launch_tb<GATED, EPI, false>(P, grid, smem, st);
}

// ---- weight images of a whole time step (see tc::wimg_kernel): the step driver runs its launch sequence twice, first in
// RECORD mode (GEMM launches only note their weights and get an arena offset; nothing else runs), converts all images
// with one kernel, then in REPLAY mode (each GEMM finds its image by position).
enum { WIMG_OFF = 0, WIMG_RECORD = 1, WIMG_REPLAY = 2 };
struct WImgState { int mode, idx; size_t cap, used; tc::WImgBatch batch; };
static thread_local WImgState g_wimg = {WIMG_OFF, 0, 0, 0, {}};
bool tc_recording() { return g_wimg.mode == WIMG_RECORD; }
void tc_wimg_begin_record(void* arena, size_t cap) {
    g_wimg.mode = WIMG_RECORD; g_wimg.idx = 0; g_wimg.cap = cap; g_wimg.used = 0; g_wimg.batch.n = 0; g_wimg.batch.base = (char*)arena;
}
void tc_wimg_off() { g_wimg.mode = WIMG_OFF; }
// ends RECORD: converts the recorded weights (one launch) and switches to REPLAY; on overflow images stay off
int tc_wimg_convert(cudaStream_t st) {
    static const bool env_on = !(getenv("URNN_WIMG") && getenv("URNN_WIMG")[0] == '0');
    static const bool bias_mma_env = getenv("URNN_BIAS_MMA") && getenv("URNN_BIAS_MMA")[0] == '1';
    if (!env_on || bias_mma_env || g_wimg.batch.n == 0 || g_wimg.batch.n > tc::WIMG_MAX || g_wimg.used > g_wimg.cap) { g_wimg.mode = WIMG_OFF; return URNN_OK; }
    tc::wimg_kernel<<<dim3(24, g_wimg.batch.n), 256, 0, st>>>(g_wimg.batch);
    URNN_LAUNCH_CHECK();
    g_wimg.mode = WIMG_REPLAY; g_wimg.idx = 0;
    return URNN_OK;
}

int tc_launch(tc::GemmParams& P, int epi, cudaStream_t st) {
    P.wimg = nullptr;
    if (g_wimg.mode == WIMG_RECORD) {
        if (g_wimg.batch.n < tc::WIMG_MAX) {
            tc::WImgSpec& S = g_wimg.batch.s[g_wimg.batch.n];
            S.W = P.W; S.w_ld = P.w_ld; S.w_ks = P.w_ks; S.nrow1 = P.nrow1; S.W2 = P.W2; S.w2_ld = P.w2_ld; S.k2 = P.k2;
            S.NOUT = P.NOUT; S.K = P.K; S.nout_store = P.nout_store; S.off = g_wimg.used;
            g_wimg.used += align_up(tc::wimg_bytes(P.NOUT, P.K), 256);
        }
        ++g_wimg.batch.n;
        return URNN_OK;
    }
    if (g_wimg.mode == WIMG_REPLAY) {
        if (g_wimg.idx < g_wimg.batch.n) {
            const tc::WImgSpec& S = g_wimg.batch.s[g_wimg.idx];
            if (S.W == P.W && S.W2 == P.W2 && S.NOUT == P.NOUT && S.K == P.K && S.nrow1 == P.nrow1 && S.k2 == P.k2 &&
                S.w_ld == P.w_ld && S.w_ks == P.w_ks && S.nout_store == P.nout_store)
                P.wimg = g_wimg.batch.base + S.off;
        }
        ++g_wimg.idx;
    }
    int g_num_sms = 0;
    {
        int dev = 0;
        URNN_CUDA(cudaGetDevice(&dev));
        URNN_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const bool gated = P.seg.gate_seg >= 0;
    static const bool bulk_env = !(getenv("URNN_BULK") && getenv("URNN_BULK")[0] == '0');
    static const bool bias_mma_env = getenv("URNN_BIAS_MMA") && getenv("URNN_BIAS_MMA")[0] == '1';
    const size_t smem = tc::plan_launch(P, epi, bulk_env, bias_mma_env);
    if (smem == 0) { set_error("tc gemm: weights %dx%d do not fit in shared memory", P.NOUT, P.K); return URNN_E_UNSUPPORTED; }
    // alternating sweep directions (L2 reuse between consecutive kernels) measured no gain at 500 x 500: off by default
    static const bool rev_env = getenv("URNN_REVERSE") && getenv("URNN_REVERSE")[0] == '1';
    P.reverse = rev_env ? g_dir : 0; g_dir ^= 1;
    P.dbg = nullptr;
    int ntiles = (P.N + tc::TILE_M - 1) / tc::TILE_M;
    int grid = ntiles < g_num_sms ? ntiles : g_num_sms;
    switch (epi) {
        case tc::EPI_GN:     return gated ? tc_launch_t<true, tc::EPI_GN>(P, grid, smem, st) : tc_launch_t<false, tc::EPI_GN>(P, grid, smem, st);
        case tc::EPI_LRELU:  return tc_launch_t<false, tc::EPI_LRELU>(P, grid, smem, st);
        case tc::EPI_POOL:   return tc_launch_t<false, tc::EPI_POOL>(P, grid, smem, st);
        case tc::EPI_DECONV: return tc_launch_t<false, tc::EPI_DECONV>(P, grid, smem, st);
    }
    set_error("tc gemm: bad epilogue %d", epi);
    return URNN_E_INVALID;
}

void tc_params_defaults(tc::GemmParams& P) {
    P.seg.gate_seg = -1; P.seg.gate_ch0 = 0; P.seg.gate_pre = nullptr; P.seg.gate_scale = nullptr; P.seg.gate_shift = nullptr;
    for (int i = 0; i < 3; ++i) { P.seg.src[i] = nullptr; P.seg.cend[i] = 0; P.seg.kind[i] = 0; P.seg.plane[i] = 0; }
    P.seg.gate_plane = 0; P.bulk = 0; P.nraw = 0; P.na = 0; P.reverse = 0; P.nstage = 0; P.bias_mma = 0; P.out_vec = 0; P.wimg = nullptr;
    P.w_ks = 1; P.W2 = nullptr; P.w2_ld = 0; P.k2 = 0; P.bias = nullptr; P.nbias = 0;
    P.out = nullptr; P.out_f32 = nullptr; P.addend = nullptr; P.nstat = 0; P.slope = 0.f; P.img_w = 0; P.n_base = 0;
    P.sink = StatSink{nullptr, nullptr, nullptr, 0, 0, CommDev{1, 0, {nullptr}, {nullptr}, nullptr}};
    P.aff = AffineOut{nullptr, nullptr, nullptr, nullptr, 0, 32, 1.0, 0.f};
    P.dbg = nullptr;
}

// ------------------------------------------------------------------------------------------------ stems on tcgen05
// y = [AvgPool2](LeakyReLU(conv1x1(x) + b)); x fp32 (kind 0) or bf16 (kind 1); exactly one of y_bf16 / y_f32 is set
int conv1x1_lrelu_fwd_tc(int Cin, int Cout, int H, int W, int pool, float slope, const void* x, int xkind,
                         const float* w, long w_ld, const float* b, __nv_bfloat16* y_bf16, float* y_f32, cudaStream_t st,
                         long x_plane, long y_plane) {
    if (x_plane == 0) x_plane = (long)H * W;
    if (Cout > 256) { set_error("conv1x1(bf16): Cout=%d > 256", Cout); return URNN_E_UNSUPPORTED; }
    if (pool == 2 && xkind != 0) { set_error("conv1x1(bf16): pooled stem needs an fp32 source"); return URNN_E_UNSUPPORTED; }
    tc::GemmParams P; tc_params_defaults(P);
    P.seg.src[0] = P.seg.src[1] = P.seg.src[2] = x; P.seg.kind[0] = P.seg.kind[1] = P.seg.kind[2] = xkind;
    P.seg.cend[0] = P.seg.cend[1] = P.seg.cend[2] = Cin; P.seg.plane[0] = P.seg.plane[1] = P.seg.plane[2] = x_plane;
    P.W = w; P.w_ld = w_ld; P.w_ks = 1; P.nrow1 = 1 << 30;
    P.bias = b; P.nbias = Cout; P.NOUT = (Cout + 31) & ~31; P.nout_store = Cout; P.K = Cin;
    P.out = y_bf16; P.out_f32 = y_f32; P.slope = slope;
    if (pool == 1) { P.N = H * W; P.out_plane = y_plane ? y_plane : (long)H * W; return tc_launch(P, tc::EPI_LRELU, st); }
    P.N = 4 * (H / 2) * (W / 2); P.out_plane = y_plane ? y_plane : (long)(H / 2) * (W / 2); P.img_w = W;
    return tc_launch(P, tc::EPI_POOL, st);
}

// y = LeakyReLU(ConvTranspose2d(k2,s2)(x) + b): GEMM rows n = co*4 + dy*2 + dx, split into launches of <= 256 rows
int deconv2x2_lrelu_fwd_tc(int Cin, int Cout, int H, int W, float slope, const void* x, int xkind, const float* w,
                           const float* b, __nv_bfloat16* y_bf16, float* y_f32, cudaStream_t st, long y_plane) {
    const int M = 4 * Cout;
    int per = 256;
    if (M <= 256) per = (M + 31) & ~31; else if (M % 192 == 0) per = 192; else if (M % 128 == 0) per = 128;
    for (int n0 = 0; n0 < M; n0 += per) {
        tc::GemmParams P; tc_params_defaults(P);
        P.seg.src[0] = P.seg.src[1] = P.seg.src[2] = x; P.seg.kind[0] = P.seg.kind[1] = P.seg.kind[2] = xkind;
        P.seg.cend[0] = P.seg.cend[1] = P.seg.cend[2] = Cin; P.seg.plane[0] = P.seg.plane[1] = P.seg.plane[2] = (long)H * W;
        P.W = w + n0; P.w_ld = 1; P.w_ks = M; P.nrow1 = 1 << 30;
        P.bias = b; P.NOUT = per; P.nout_store = (M - n0 < per) ? (M - n0) : per; P.K = Cin; P.N = H * W;
        P.out = y_bf16; P.out_f32 = y_f32; P.out_plane = y_plane ? y_plane : (long)4 * H * W; P.slope = slope; P.img_w = W; P.n_base = n0;
        URNN_TRY(tc_launch(P, tc::EPI_DECONV, st));
    }
    return URNN_OK;
}

// h' = (1-z)*h + z*tanh(GN2(C)),  z = sigmoid(GN1(G)[:F]); G and C are bf16 maps with padded planes (gplane).
// Channel-major streaming: thread = VEC consecutive pixels of one channel, four sequential streams per channel.
template <int VEC>
__global__ void __launch_bounds__(256)
cgru_blend_bf16_kernel(const __nv_bfloat16* __restrict__ G, const __nv_bfloat16* __restrict__ C,
                       const float* __restrict__ h, const float* __restrict__ sc1, const float* __restrict__ sh1,
                       const float* __restrict__ sc2, const float* __restrict__ sh2, float* __restrict__ h_out,
                       long N, long gplane, unsigned nq, long total) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const unsigned c = (unsigned)(idx / nq);
    const long p = (idx - (long)c * nq) * VEC;
    float gv[VEC], cv[VEC], hv[VEC], o[VEC];
    if constexpr (VEC == 4) {
        const uint2 gw = __ldg(reinterpret_cast<const uint2*>(G + (long)c * gplane + p));
        const uint2 cw = __ldg(reinterpret_cast<const uint2*>(C + (long)c * gplane + p));
        const float4 hh = __ldg(reinterpret_cast<const float4*>(h + (long)c * N + p));
        tc::unpack_bf16x2(gw.x, gv[0], gv[1]); tc::unpack_bf16x2(gw.y, gv[2], gv[3]);
        tc::unpack_bf16x2(cw.x, cv[0], cv[1]); tc::unpack_bf16x2(cw.y, cv[2], cv[3]);
        hv[0] = hh.x; hv[1] = hh.y; hv[2] = hh.z; hv[3] = hh.w;
    } else {
        gv[0] = __bfloat162float(G[(long)c * gplane + p]);
        cv[0] = __bfloat162float(C[(long)c * gplane + p]);
        hv[0] = __ldg(h + (long)c * N + p);
    }
    const float a1 = __ldg(sc1 + c), b1 = __ldg(sh1 + c), a2 = __ldg(sc2 + c), b2 = __ldg(sh2 + c);
#pragma unroll
    for (int u = 0; u < VEC; ++u) {
        const float z = sigmoid_fast(fmaf(gv[u], a1, b1)), t = tanh_fast(fmaf(cv[u], a2, b2));
        o[u] = (1.f - z) * hv[u] + z * t;
    }
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(h_out + (long)c * N + p) = make_float4(o[0], o[1], o[2], o[3]);
    else h_out[(long)c * N + p] = o[0];
}

struct CellWsBf16 {
    unsigned* counter; double2 *total1, *total2; float *scale1, *shift1, *scale2, *shift2;
    float2 *partial1, *partial2; __nv_bfloat16 *GP, *C; int gx; char* wimg;
};
constexpr size_t CELL_WIMG_BYTES = 256 << 10;      // weight images of a stand-alone cell call (sweep A + sweep B)

static size_t cell_ws_bf16(const urnn_cell_desc* d, void* ws, size_t ws_bytes, CellWsBf16* out) {
    const long N = (long)d->H * d->W;
    const int F = d->F, gx = (int)((N + tc::TILE_M - 1) / tc::TILE_M);
    Arena a(ws, ws_bytes);
    CellWsBf16 w;
    w.gx = gx;
    w.counter = a.take<unsigned>(64);
    w.total1 = a.take<double2>(2 * F / 32); w.total2 = a.take<double2>(F / 32);
    w.scale1 = a.take<float>(2 * F); w.shift1 = a.take<float>(2 * F);
    w.scale2 = a.take<float>(F);     w.shift2 = a.take<float>(F);
    w.partial1 = a.take<float2>((size_t)(2 * F / 32) * gx);
    w.partial2 = a.take<float2>((size_t)(F / 32) * gx);
    const size_t Np = (size_t)pad_plane(N);            // bf16 planes are padded to whole 128-pixel tiles (bulk copies)
    w.GP = a.take<__nv_bfloat16>((size_t)3 * F * Np);
    w.C = a.take<__nv_bfloat16>((size_t)F * Np);
    w.wimg = a.take<char>(CELL_WIMG_BYTES);
    if (out) *out = w;
    return align_up(a.off, 256);
}

size_t cgru_fwd_bf16_workspace(const urnn_cell_desc* d) { return cell_ws_bf16(d, nullptr, 0, nullptr); }

int tc_counters_begin(const urnn_cell_desc* d, void* cell_ws, size_t ws_bytes, cudaStream_t st) {
    CellWsBf16 w;
    cell_ws_bf16(d, cell_ws, ws_bytes, &w);          // the counters sit at the same offset for every cell geometry
    URNN_CUDA(cudaMemsetAsync(w.counter, 0, 64 * sizeof(unsigned), st));
    g_counters_clean = true;
    return URNN_OK;
}

int cgru_fwd_bf16(const urnn_cell_desc* d, const urnn_cell_params* p, const void* x, int xkind, const float* e,
                  const float* h, float* h_out, void* ws, size_t ws_bytes, cudaStream_t st, long x_plane) {
    if (d->ksize != 1) {                                                                 // k>1: fp32 taps path
        if (xkind != 0) { set_error("cgru_fwd: bf16 x needs k=1"); return URNN_E_UNSUPPORTED; }
        return cgru_fwd_fp32(d, p, (const float*)x, e, h, h_out, ws, ws_bytes, st);
    }
    const int F = d->F;
    const long N = (long)d->H * d->W, Np = pad_plane(N);
    if (x_plane == 0) x_plane = N;
    const int Ch = (d->variant == URNN_CELL_DECODER) ? 2 * F : F;
    const int Ktot = d->Cx + Ch;
    const int Cx_eff = x ? d->Cx : 0;
    const int Keff = Cx_eff + Ch;
    const int Kxe = Keff - F;                           // channels that do not pass through the reset gate
    const long aoff = x ? 0 : d->Cx;                    // skip the zero-input weight columns (ConvRNN.py:143-146)
    if (2 * F > 256) { set_error("cgru_fwd(bf16): num_features=%d > 128 not supported by the tcgen05 tile", F); return URNN_E_UNSUPPORTED; }
    const bool split = (3 * F <= 256) && Kxe > 0;       // compute the r-independent part of the candidate in sweep A
    CellWsBf16 w;
    size_t need = cell_ws_bf16(d, ws, ws_bytes, &w);
    if (need > ws_bytes) { set_error("cgru_fwd: workspace %zu < %zu bytes", ws_bytes, need); return URNN_E_WORKSPACE; }
    // the ticket counters reset themselves (last CTA); the step driver clears them once per step instead of once per cell
    if (!g_counters_clean) URNN_CUDA(cudaMemsetAsync(w.counter, 0, 64 * sizeof(unsigned), st));

    tc::GemmParams P; tc_params_defaults(P);
    // segments [x | e | h]; missing ones get zero width
    int n = 0; const void* srcs[3] = {h, h, h}; int cnt[3] = {0, 0, 0}; int kinds[3] = {0, 0, 0}; long planes[3] = {N, N, N};
    if (x) { srcs[n] = x; cnt[n] = d->Cx; kinds[n] = xkind; planes[n] = x_plane; ++n; }
    if (d->variant == URNN_CELL_DECODER) { srcs[n] = e; cnt[n] = F; ++n; }
    srcs[n] = h; cnt[n] = F; const int hseg = n; ++n;
    int acc = 0;
    for (int i = 0; i < 3; ++i) { P.seg.src[i] = srcs[i]; P.seg.kind[i] = kinds[i]; P.seg.plane[i] = planes[i]; acc += cnt[i]; P.seg.cend[i] = acc; }
    P.seg.gate_plane = Np;
    P.seg.gate_seg = -1; P.seg.gate_ch0 = 0; P.seg.gate_pre = nullptr; P.seg.gate_scale = nullptr; P.seg.gate_shift = nullptr;
    P.N = (int)N; P.K = Keff; P.out_plane = Np; P.addend = nullptr;

    // ---- sweep A
    P.W = p->w1 + aoff; P.w_ld = Ktot; P.nrow1 = 2 * F;
    P.W2 = p->w2 + aoff; P.w2_ld = Ktot; P.k2 = Kxe;
    P.bias = p->b1; P.nbias = 2 * F;
    P.NOUT = split ? 3 * F : 2 * F; P.nout_store = P.NOUT;
    P.out = w.GP; P.nstat = 2 * F / 32;
    CommDev comm; current_comm(&comm);
    const double gcount = 32.0 * (double)N * (double)(comm.world > 1 ? comm.world : 1);
    P.sink = StatSink{w.partial1, w.total1, w.counter, 2 * F / 32, w.gx, comm};
    P.aff = AffineOut{w.scale1, w.shift1, p->gn1_w, p->gn1_b, 2 * F, 32, gcount, d->eps};
    URNN_TRY(tc_launch(P, tc::EPI_GN, st));

    // ---- sweep B
    P.seg.gate_ch0 = F; P.seg.gate_pre = w.GP; P.seg.gate_scale = w.scale1; P.seg.gate_shift = w.shift1;
    if (split) {
        P.seg.src[0] = P.seg.src[1] = P.seg.src[2] = h; P.seg.kind[0] = P.seg.kind[1] = P.seg.kind[2] = 0;
        P.seg.plane[0] = P.seg.plane[1] = P.seg.plane[2] = N;
        P.seg.cend[0] = P.seg.cend[1] = P.seg.cend[2] = F;
        P.seg.gate_seg = 0;
        P.K = F; P.W = p->w2 + aoff + Kxe;
        P.addend = w.GP + (size_t)2 * F * Np;
    } else {
        P.seg.gate_seg = hseg;
        P.W = p->w2 + aoff;
    }
    P.w_ld = Ktot; P.nrow1 = F; P.W2 = nullptr; P.w2_ld = 0; P.k2 = 0;
    P.bias = p->b2; P.nbias = F; P.NOUT = F; P.nout_store = F; P.out = w.C; P.nstat = F / 32;
    P.sink = StatSink{w.partial2, w.total2, w.counter + 1, F / 32, w.gx, comm};
    P.aff = AffineOut{w.scale2, w.shift2, p->gn2_w, p->gn2_b, F, 32, gcount, d->eps};
    URNN_TRY(tc_launch(P, tc::EPI_GN, st));

    // ---- sweep C
    if (tc_recording()) return URNN_OK;
    {
        static const bool pdl = !(getenv("URNN_PDL") && getenv("URNN_PDL")[0] == '0');
        const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(h_out)) & 15) == 0;
        const long nq = vec ? N / 4 : N, total = nq * F;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((total + 255) / 256));
        cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        g_dir ^= 1;
        if (vec)
            URNN_CUDA(cudaLaunchKernelEx(&cfg, cgru_blend_bf16_kernel<4>, (const __nv_bfloat16*)w.GP, (const __nv_bfloat16*)w.C, h,
                                         (const float*)w.scale1, (const float*)w.shift1, (const float*)w.scale2,
                                         (const float*)w.shift2, h_out, N, Np, (unsigned)nq, total));
        else
            URNN_CUDA(cudaLaunchKernelEx(&cfg, cgru_blend_bf16_kernel<1>, (const __nv_bfloat16*)w.GP, (const __nv_bfloat16*)w.C, h,
                                         (const float*)w.scale1, (const float*)w.shift1, (const float*)w.scale2,
                                         (const float*)w.shift2, h_out, N, Np, (unsigned)nq, total));
    }
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}

// Stand-alone cell call (urnn_cgru_fwd): same record -> convert -> replay of the weight images as the step driver.
int cgru_fwd_bf16_standalone(const urnn_cell_desc* d, const urnn_cell_params* p, const void* x, int xkind, const float* e,
                             const float* h, float* h_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (d->ksize != 1 || g_wimg.mode != WIMG_OFF) return cgru_fwd_bf16(d, p, x, xkind, e, h, h_out, ws, ws_bytes, st);
    CellWsBf16 w;
    size_t need = cell_ws_bf16(d, ws, ws_bytes, &w);
    if (need > ws_bytes) { set_error("cgru_fwd: workspace %zu < %zu bytes", ws_bytes, need); return URNN_E_WORKSPACE; }
    URNN_CUDA(cudaMemsetAsync(w.counter, 0, 64 * sizeof(unsigned), st));
    struct Guard { ~Guard() { g_counters_clean = false; tc_wimg_off(); } } guard;
    g_counters_clean = true;
    tc_wimg_begin_record(w.wimg, CELL_WIMG_BYTES);
    tc_reset_direction();
    URNN_TRY(cgru_fwd_bf16(d, p, x, xkind, e, h, h_out, ws, ws_bytes, st));
    URNN_TRY(tc_wimg_convert(st));
    tc_reset_direction();
    return cgru_fwd_bf16(d, p, x, xkind, e, h, h_out, ws, ws_bytes, st);
}

}  // namespace urnn
