// urnn_bwd.cu -- backward of the stage stems and of the (Skip-)ConvGRU cell (fp32 arithmetic).
//
// Recompute-in-backward, like the reference's reentrant checkpointing (ConvRNN.py:154-158, encoder.py:146-149): only
// the op's inputs are saved; the forward sweeps are re-run to rebuild the pre-normalisation maps and statistics.
// The channel-mixing products reuse the fp32 pixel-GEMM: d(input) = W^T d(output) is the same kernel with a transposed
// weight view and a routing epilogue that drops the rows straight into dx / de / dh; weight gradients are a
// pixel-split outer-product GEMM (wgrad_kernel).  GroupNorm backward adds one statistic level per normalisation:
//   dx = rstd * (gamma*dy - mean_g(gamma*dy) - xhat * mean_g(gamma*dy*xhat)).
// Gradients of a bf16-mode forward are computed by this fp32 path as well (recompute in fp32).
#include "pixgemm.cuh"
#include "head_bwd.cuh"
#include "urnn_internal.h"

namespace urnn {

// ------------------------------------------------------------------------------------------------ stems
size_t conv1x1_lrelu_bwd_workspace(int Cin, int Cout, int H, int W, int pool) {
    (void)Cin; (void)pool;
    return align_up((size_t)Cout * H * W * sizeof(float), 256) + 256;
}

int conv1x1_lrelu_bwd_fp32(int Cin, int Cout, int H, int W, int pool, float slope, const float* x, const float* w,
                           const float* b, const float* dy, float* dx, float* dw, float* db, void* ws, size_t ws_bytes,
                           cudaStream_t st) {
    const long N = (long)H * W;
    if (conv1x1_lrelu_bwd_workspace(Cin, Cout, H, W, pool) > ws_bytes) { set_error("conv1x1_lrelu_bwd: workspace too small"); return URNN_E_WORKSPACE; }
    float* dpre = (float*)ws;
    SegLoader Lx = single_map_loader(x, Cin, N);
    // 1. dpre = dy (gathered through the pooling) * LeakyReLU'(W x + b)
    DpreEpilogue ep{b, 0, dy, dpre, N, pool == 2 ? 1 : 0, W, pool == 2 ? (long)(H / 2) * (W / 2) : N, slope};
    URNN_TRY(launch_pixgemm(AView{w, (long)Cin, 1}, Cout, Cin, (int)N, Lx, ep, false, st));
    // 2. dx = W^T dpre
    if (dx != nullptr) {
        SegLoader Ld = single_map_loader(dpre, Cout, N);
        RouteEpilogue re{{dx, dx, dx}, {Cin, Cin, Cin}, {1, 1, 1}, N};
        URNN_TRY(launch_pixgemm(AView{w, 1, (long)Cin}, Cin, Cout, (int)N, Ld, re, false, st));
    }
    // 3. dW += dpre x^T, db += sum dpre
    if (dw != nullptr) URNN_TRY(launch_wgrad(dpre, N, Cout, Cin, (int)N, Lx, dw, (long)Cin, 1, db, 0, st));
    return URNN_OK;
}

size_t deconv2x2_lrelu_bwd_workspace(int Cin, int Cout, int H, int W) {
    (void)Cin;
    return align_up((size_t)4 * Cout * H * W * sizeof(float), 256) + 256;
}

int deconv2x2_lrelu_bwd_fp32(int Cin, int Cout, int H, int W, float slope, const float* x, const float* w, const float* b,
                             const float* dy, float* dx, float* dw, float* db, void* ws, size_t ws_bytes, cudaStream_t st) {
    const long N = (long)H * W;
    const int M = 4 * Cout;
    if (deconv2x2_lrelu_bwd_workspace(Cin, Cout, H, W) > ws_bytes) { set_error("deconv2x2_lrelu_bwd: workspace too small"); return URNN_E_WORKSPACE; }
    float* dpre = (float*)ws;                        // [M][N], row m = co*4 + dy*2 + dx
    SegLoader Lx = single_map_loader(x, Cin, N);
    DpreEpilogue ep{b, 2, dy, dpre, N, 2, W, 4 * N, slope};
    URNN_TRY(launch_pixgemm(AView{w, 1, (long)M}, M, Cin, (int)N, Lx, ep, false, st));
    if (dx != nullptr) {
        SegLoader Ld = single_map_loader(dpre, M, N);
        RouteEpilogue re{{dx, dx, dx}, {Cin, Cin, Cin}, {1, 1, 1}, N};
        URNN_TRY(launch_pixgemm(AView{w, (long)M, 1}, Cin, M, (int)N, Ld, re, false, st));
    }
    if (dw != nullptr) URNN_TRY(launch_wgrad(dpre, N, M, Cin, (int)N, Lx, dw, 1, (long)M, db, 2, st));
    return URNN_OK;
}

// ------------------------------------------------------------------------------------------------ cell
struct GnStat { float mu, rstd; };

__global__ void gn_mu_rstd_kernel(const double2* __restrict__ total, double count, float eps, GnStat* __restrict__ out, int ngroups) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    double2 t = total[g];
    double mean = t.x / count, var = t.y / count - mean * mean;
    if (var < 0.0) var = 0.0;
    out[g].mu = (float)mean;
    out[g].rstd = (float)(1.0 / sqrt(var + (double)eps));
}

__device__ __forceinline__ void block_add2(double a, double b, double* dst) {
    __shared__ double sh[2][8];
    a = warp_sum(a); b = warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += sh[0][w]; tb += sh[1][w]; }
        atomicAdd(dst, ta); atomicAdd(dst + 1, tb);
    }
    __syncthreads();
}

// E1: gates / blend backward for channel c = blockIdx.y over a pixel chunk.
//   z = sigmoid(GN1(G)[c]); ct = tanh(GN2(C)[c]); h' = (1-z) h + z ct
//   dh = dh' (1-z) (stored); dyc = dh' z (1-ct^2) (stored in DC); dyz = dh' (ct-h) z (1-z) (stored in DY1 z-half)
//   per-channel sums for the two GroupNorm backward passes.
__global__ void __launch_bounds__(256)
cell_bwd_e1_kernel(const float* __restrict__ G, const float* __restrict__ C, const float* __restrict__ h,
                   const float* __restrict__ dho, const float* __restrict__ sc1, const float* __restrict__ sh1,
                   const float* __restrict__ sc2, const float* __restrict__ sh2, const GnStat* __restrict__ st1,
                   const GnStat* __restrict__ st2, float* __restrict__ dh, float* __restrict__ DC, float* __restrict__ DY1,
                   double* __restrict__ sums2, double* __restrict__ sums1, long N) {
    const int c = blockIdx.y;
    const float a1 = sc1[c], b1 = sh1[c], a2 = sc2[c], b2 = sh2[c];
    const GnStat g1 = st1[c >> 5], g2 = st2[c >> 5];
    double s2a = 0.0, s2b = 0.0, s1a = 0.0, s1b = 0.0;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long)gridDim.x * blockDim.x) {
        const long o = (long)c * N + p;
        const float gz = G[o], cc = C[o], hv = h[o], d = dho[o];
        const float z = sigmoid_acc(fmaf(gz, a1, b1)), ct = tanhf(fmaf(cc, a2, b2));
        dh[o] = d * (1.f - z);
        const float dyc = d * z * (1.f - ct * ct);
        const float dyz = d * (ct - hv) * z * (1.f - z);
        DC[o] = dyc; DY1[o] = dyz;
        s2a += dyc; s2b += dyc * ((cc - g2.mu) * g2.rstd);
        s1a += dyz; s1b += dyz * ((gz - g1.mu) * g1.rstd);
    }
    block_add2(s2a, s2b, sums2 + 2 * c);
    block_add2(s1a, s1b, sums1 + 2 * c);
}

// per-channel sums -> parameter gradients and per-group GroupNorm-backward coefficients (one block)
//   dgamma[c] += B_c, dbeta[c] += A_c;  m1[g] = sum_c gamma_c A_c / n,  m2[g] = sum_c gamma_c B_c / n
// Spatial sharding: the per-group sums run over the whole grid, so they are all-gathered over NVLink (ll_allgather) and
// added in rank order; `count` is the global element count.  dgamma / dbeta stay rank-local partial sums (the caller
// all-reduces the gradients of replicated parameters once per window).
__global__ void gn_bwd_coef_kernel(const double* __restrict__ sums, const float* __restrict__ gamma, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta, float2* __restrict__ coef, int channels, double count, CommDev comm) {
    __shared__ double s1[8], s2[8];
    __shared__ unsigned xin[8 * 4], xout[COMM_MAX_WORLD * 8 * 4];
    const int c = threadIdx.x;
    const int ngroups = channels >> 5;
    if (c < 8) { s1[c] = 0.0; s2[c] = 0.0; }
    __syncthreads();
    if (c < channels) {
        const double A = sums[2 * c], B = sums[2 * c + 1];
        if (dgamma) dgamma[c] += (float)B;
        if (dbeta) dbeta[c] += (float)A;
        atomicAdd(&s1[c >> 5], (double)gamma[c] * A);
        atomicAdd(&s2[c >> 5], (double)gamma[c] * B);
    }
    __syncthreads();
    if (comm.world > 1) {
        if (c < ngroups) {
            const unsigned long long a = (unsigned long long)__double_as_longlong(s1[c]), b = (unsigned long long)__double_as_longlong(s2[c]);
            xin[4 * c] = (unsigned)a; xin[4 * c + 1] = (unsigned)(a >> 32); xin[4 * c + 2] = (unsigned)b; xin[4 * c + 3] = (unsigned)(b >> 32);
        }
        __syncthreads();
        ll_allgather(comm, xin, ngroups, 4, xout);
        if (c < ngroups) {
            double a = 0.0, b = 0.0;
            for (int r = 0; r < comm.world; ++r) {
                const unsigned* w = xout + (r * ngroups + c) * 4;
                a += __longlong_as_double((long long)((unsigned long long)w[0] | ((unsigned long long)w[1] << 32)));
                b += __longlong_as_double((long long)((unsigned long long)w[2] | ((unsigned long long)w[3] << 32)));
            }
            s1[c] = a; s2[c] = b;
        }
        __syncthreads();
    }
    if (c < ngroups) coef[c] = make_float2((float)(s1[c] / count), (float)(s2[c] / count));
}

// in-place sum over the ranks of `nsets` groups of two doubles (LayerNorm backward sums of the head); one block
__global__ void comm_sum2_kernel(double* __restrict__ v, int nsets, CommDev comm) {
    __shared__ unsigned xin[8 * 4], xout[COMM_MAX_WORLD * 8 * 4];
    const int c = threadIdx.x;
    if (c < nsets) {
        const unsigned long long a = (unsigned long long)__double_as_longlong(v[2 * c]), b = (unsigned long long)__double_as_longlong(v[2 * c + 1]);
        xin[4 * c] = (unsigned)a; xin[4 * c + 1] = (unsigned)(a >> 32); xin[4 * c + 2] = (unsigned)b; xin[4 * c + 3] = (unsigned)(b >> 32);
    }
    __syncthreads();
    ll_allgather(comm, xin, nsets, 4, xout);
    if (c < nsets) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < comm.world; ++r) {
            const unsigned* w = xout + (r * nsets + c) * 4;
            a += __longlong_as_double((long long)((unsigned long long)w[0] | ((unsigned long long)w[1] << 32)));
            b += __longlong_as_double((long long)((unsigned long long)w[2] | ((unsigned long long)w[3] << 32)));
        }
        v[2 * c] = a; v[2 * c + 1] = b;
    }
}

// dX = rstd * (gamma*dy - m1 - xhat*m2), in place over the dy map; X is the pre-normalisation map
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(float* __restrict__ DY, const float* __restrict__ X, const float* __restrict__ gamma,
                    const GnStat* __restrict__ st, const float2* __restrict__ coef, long N) {
    const int c = blockIdx.y;
    const GnStat g = st[c >> 5];
    const float2 m = coef[c >> 5];
    const float ga = gamma[c];
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long)gridDim.x * blockDim.x) {
        const long o = (long)c * N + p;
        const float xh = (X[o] - g.mu) * g.rstd;
        DY[o] = g.rstd * (ga * DY[o] - m.x - xh * m.y);
    }
}

// E3: reset-gate backward for channel c: r = sigmoid(GN1(G)[F+c]); d(r*h) = DRH
//   dh += DRH r;  dyr = DRH h r (1-r) -> DY1[F+c];  per-channel sums for GroupNorm-1 backward (r half)
__global__ void __launch_bounds__(256)
cell_bwd_e3_kernel(const float* __restrict__ G, const float* __restrict__ h, const float* __restrict__ DRH,
                   const float* __restrict__ sc1, const float* __restrict__ sh1, const GnStat* __restrict__ st1,
                   float* __restrict__ dh, float* __restrict__ DY1, double* __restrict__ sums1, int F, long N) {
    const int c = blockIdx.y, cg = F + c;
    const float a1 = sc1[cg], b1 = sh1[cg];
    const GnStat g1 = st1[cg >> 5];
    double sa = 0.0, sb = 0.0;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long)gridDim.x * blockDim.x) {
        const long o = (long)c * N + p, og = (long)cg * N + p;
        const float gr = G[og], d = DRH[o], hv = h[o];
        const float r = sigmoid_acc(fmaf(gr, a1, b1));
        dh[o] += d * r;
        const float dyr = d * hv * r * (1.f - r);
        DY1[og] = dyr;
        sa += dyr; sb += dyr * ((gr - g1.mu) * g1.rstd);
    }
    block_add2(sa, sb, sums1 + 2 * cg);
}

struct CellBwdWs { float *DC, *DRH, *DY1, *dh_scratch; GnStat *st1, *st2; double *sums1, *sums2; float2 *coef1, *coef2; void* fwd_ws; size_t fwd_bytes; };

static size_t cell_bwd_layout(const urnn_cell_desc* d, void* ws, size_t ws_bytes, CellBwdWs* out) {
    const long N = (long)d->H * d->W;
    const int F = d->F;
    Arena a(ws, ws_bytes);
    CellBwdWs w;
    w.sums1 = a.take<double>(4 * F); w.sums2 = a.take<double>(2 * F);
    w.st1 = a.take<GnStat>(2 * F / 32); w.st2 = a.take<GnStat>(F / 32);
    w.coef1 = a.take<float2>(2 * F / 32); w.coef2 = a.take<float2>(F / 32);
    w.DC = a.take<float>((size_t)F * N); w.DRH = a.take<float>((size_t)F * N);
    w.DY1 = a.take<float>((size_t)2 * F * N); w.dh_scratch = a.take<float>((size_t)F * N);
    w.fwd_bytes = cgru_fwd_fp32_workspace(d);
    w.fwd_ws = a.take<char>(w.fwd_bytes);
    if (out) *out = w;
    return align_up(a.off, 256);
}

size_t cgru_bwd_workspace(const urnn_cell_desc* d) { return cell_bwd_layout(d, nullptr, 0, nullptr); }

int cgru_bwd_fp32(const urnn_cell_desc* d, const urnn_cell_params* p, const float* x, const float* e, const float* h,
                  const float* dh_out, float* dx, float* de, float* dh, const urnn_cell_grads* gr, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
    if (d->ksize != 1) { set_error("cgru_bwd: only 1x1 gates are implemented (the encoder-decoder never uses k>1)"); return URNN_E_UNSUPPORTED; }
    const int F = d->F;
    const long N = (long)d->H * d->W;
    const int Ch = (d->variant == URNN_CELL_DECODER) ? 2 * F : F;
    const int Ktot = d->Cx + Ch;
    const int Cx_eff = x ? d->Cx : 0;
    const int Keff = Cx_eff + Ch;
    const long aoff = x ? 0 : d->Cx;
    CommDev comm; current_comm(&comm);
    if (F > 128) { set_error("cgru_bwd: num_features=%d > 128 is not supported (8 GroupNorm groups per statistics exchange)", F); return URNN_E_UNSUPPORTED; }
    CellBwdWs w;
    size_t need = cell_bwd_layout(d, ws, ws_bytes, &w);
    if (need > ws_bytes) { set_error("cgru_bwd: workspace %zu < %zu bytes", ws_bytes, need); return URNN_E_WORKSPACE; }
    if (dh == nullptr) dh = w.dh_scratch;
    const double count = 32.0 * (double)N * (double)(comm.world > 1 ? comm.world : 1);      // GroupNorm spans the whole (sharded) grid

    // ---- recompute the forward sweeps: G, C, folded affines, statistics
    CellWsView f;
    URNN_TRY(cgru_fwd_fp32_passes(d, p, x, e, h, &f, w.fwd_ws, w.fwd_bytes, st));
    gn_mu_rstd_kernel<<<1, 32, 0, st>>>(f.total1, count, d->eps, w.st1, 2 * F / 32); URNN_LAUNCH_CHECK();
    gn_mu_rstd_kernel<<<1, 32, 0, st>>>(f.total2, count, d->eps, w.st2, F / 32); URNN_LAUNCH_CHECK();
    URNN_CUDA(cudaMemsetAsync(w.sums1, 0, sizeof(double) * 4 * F, st));
    URNN_CUDA(cudaMemsetAsync(w.sums2, 0, sizeof(double) * 2 * F, st));

    int chunks = (int)((N + 256 * 8 - 1) / (256 * 8)); if (chunks > 296) chunks = 296; if (chunks < 1) chunks = 1;
    dim3 gridF(chunks, F), grid2F(chunks, 2 * F);
    // ---- E1: blend / tanh / update-gate backward
    cell_bwd_e1_kernel<<<gridF, 256, 0, st>>>(f.G, f.C, h, dh_out, f.scale1, f.shift1, f.scale2, f.shift2, w.st1, w.st2, dh,
                                              w.DC, w.DY1, w.sums2, w.sums1, N);
    URNN_LAUNCH_CHECK();
    // ---- GroupNorm-2 backward -> dC
    gn_bwd_coef_kernel<<<1, 256, 0, st>>>(w.sums2, p->gn2_w, gr ? gr->gn2_w : nullptr, gr ? gr->gn2_b : nullptr, w.coef2, F, count, comm);
    URNN_LAUNCH_CHECK();
    gn_bwd_apply_kernel<<<gridF, 256, 0, st>>>(w.DC, f.C, p->gn2_w, w.st2, w.coef2, N); URNN_LAUNCH_CHECK();
    // ---- candidate conv backward: dW2 += dC [x|e|r*h]^T ; d[x|e|r*h] = W2^T dC
    SegLoader L2;
    {
        int n = 0, acc = 0; const float* srcs[3] = {h, h, h}; int cnt[3] = {0, 0, 0};
        if (x) { srcs[n] = x; cnt[n] = d->Cx; ++n; }
        if (d->variant == URNN_CELL_DECODER) { srcs[n] = e; cnt[n] = F; ++n; }
        srcs[n] = h; cnt[n] = F; ++n;
        if (n == 1) { srcs[1] = srcs[0]; cnt[1] = cnt[0]; cnt[0] = 0; n = 2; }
        for (int i = 0; i < 3; ++i) { L2.src[i] = (i < n) ? srcs[i] : srcs[n - 1]; L2.cnt[i] = (i < n) ? cnt[i] : 0; acc += L2.cnt[i]; L2.cend[i] = acc; }
        L2.plane = N; L2.vec = false;
        L2.gate_pre = f.G; L2.gate_scale = f.scale1; L2.gate_shift = f.shift1; L2.gate_ch0 = F;
    }
    SegLoader L1 = L2; L1.gate_pre = nullptr; L1.gate_scale = nullptr; L1.gate_shift = nullptr;
    if (gr) URNN_TRY(launch_wgrad(w.DC, N, F, Keff, (int)N, L2, gr->w2 + aoff, (long)Ktot, 1, gr->b2, 0, st));
    {
        SegLoader Ld = single_map_loader(w.DC, F, N);
        RouteEpilogue re;
        re.plane = N;
        // rows of W2^T dC: [x (Cx_eff) | e (F, decoder) | h (F)] -> dx (store), de (store), DRH (store)
        int ends[3]; float* dsts[3]; int modes[3]; int n = 0, acc = 0;
        if (x) { acc += d->Cx; ends[n] = acc; dsts[n] = dx; modes[n] = dx ? 1 : 0; ++n; }
        if (d->variant == URNN_CELL_DECODER) { acc += F; ends[n] = acc; dsts[n] = de; modes[n] = de ? 1 : 0; ++n; }
        acc += F; ends[n] = acc; dsts[n] = w.DRH; modes[n] = 1; ++n;
        for (int i = 0; i < 3; ++i) { int j = i < n ? i : n - 1; re.mend[i] = ends[j]; re.dst[i] = dsts[j]; re.mode[i] = modes[j]; }
        URNN_TRY(launch_pixgemm(AView{p->w2 + aoff, 1, (long)Ktot}, Keff, F, (int)N, Ld, re, false, st));
    }
    // ---- E3: reset-gate backward; GroupNorm-1 backward -> dG
    cell_bwd_e3_kernel<<<gridF, 256, 0, st>>>(f.G, h, w.DRH, f.scale1, f.shift1, w.st1, dh, w.DY1, w.sums1, F, N);
    URNN_LAUNCH_CHECK();
    gn_bwd_coef_kernel<<<1, 256, 0, st>>>(w.sums1, p->gn1_w, gr ? gr->gn1_w : nullptr, gr ? gr->gn1_b : nullptr, w.coef1, 2 * F, count, comm);
    URNN_LAUNCH_CHECK();
    gn_bwd_apply_kernel<<<grid2F, 256, 0, st>>>(w.DY1, f.G, p->gn1_w, w.st1, w.coef1, N); URNN_LAUNCH_CHECK();
    // ---- gate conv backward: dW1 += dG [x|e|h]^T ; d[x|e|h] += W1^T dG
    if (gr) URNN_TRY(launch_wgrad(w.DY1, N, 2 * F, Keff, (int)N, L1, gr->w1 + aoff, (long)Ktot, 1, gr->b1, 0, st));
    {
        SegLoader Ld = single_map_loader(w.DY1, 2 * F, N);
        RouteEpilogue re;
        re.plane = N;
        int ends[3]; float* dsts[3]; int modes[3]; int n = 0, acc = 0;
        if (x) { acc += d->Cx; ends[n] = acc; dsts[n] = dx; modes[n] = dx ? 2 : 0; ++n; }
        if (d->variant == URNN_CELL_DECODER) { acc += F; ends[n] = acc; dsts[n] = de; modes[n] = de ? 2 : 0; ++n; }
        acc += F; ends[n] = acc; dsts[n] = dh; modes[n] = 2; ++n;
        for (int i = 0; i < 3; ++i) { int j = i < n ? i : n - 1; re.mend[i] = ends[j]; re.dst[i] = dsts[j]; re.mode[i] = modes[j]; }
        URNN_TRY(launch_pixgemm(AView{p->w1 + aoff, 1, (long)Ktot}, Keff, 2 * F, (int)N, Ld, re, false, st));
    }
    return URNN_OK;
}

// ------------------------------------------------------------------------------------------------ head
struct HeadBwdWs { float2* partial; double2* total; unsigned* counter; double *bsum, *psum; float* m[4]; int gx; };
static size_t head_bwd_layout(int H, int W, void* ws, size_t ws_bytes, HeadBwdWs* out) {
    const long N = (long)H * W;
    Arena a(ws, ws_bytes);
    HeadBwdWs w;
    w.gx = (int)((N + 127) / 128);
    w.counter = a.take<unsigned>(64);
    w.total = a.take<double2>(8);
    w.bsum = a.take<double>(16);
    w.psum = a.take<double>(40);
    w.partial = a.take<float2>((size_t)5 * w.gx);
    for (int i = 0; i < 4; ++i) w.m[i] = a.take<float>((size_t)16 * N);
    if (out) *out = w;
    return align_up(a.off, 256);
}
size_t head_bwd_workspace(int H, int W) { return head_bwd_layout(H, W, nullptr, 0, nullptr); }

int head_bwd_fp32(int H, int W, float cls_thred, float ln_eps, float slope, const urnn_head_params* p, const float* feat,
                  const float* dout, float* dfeat, const urnn_head_grads* g, void* ws, size_t ws_bytes, cudaStream_t st) {
    CommDev comm; current_comm(&comm);
    HeadBwdWs w;
    size_t need = head_bwd_layout(H, W, ws, ws_bytes, &w);
    if (need > ws_bytes) { set_error("head_bwd: workspace %zu < %zu bytes", ws_bytes, need); return URNN_E_WORKSPACE; }
    const long N = (long)H * W;
    URNN_CUDA(cudaMemsetAsync(w.counter, 0, 64 * sizeof(unsigned), st));
    URNN_CUDA(cudaMemsetAsync(w.bsum, 0, 16 * sizeof(double), st));
    URNN_CUDA(cudaMemsetAsync(w.psum, 0, 40 * sizeof(double), st));
    HeadBwdDev hb;
    hb.fwd.p = *p; hb.fwd.cls_thred = cls_thred; hb.fwd.eps = ln_eps; hb.fwd.slope = slope; hb.fwd.plane = N;
    hb.fwd.count = 16.0 * (double)N * (double)(comm.world > 1 ? comm.world : 1);     // LayerNorm([16,H,W]) over the whole (sharded) grid
    hb.fwd.sink = StatSink{w.partial, w.total, w.counter, 5, w.gx, comm};
    hb.g = *g; hb.bsum = w.bsum; hb.psum = w.psum; hb.dout = dout; hb.dfeat = dfeat;
    hb.m0 = w.m[0]; hb.m1 = w.m[1]; hb.m2 = w.m[2]; hb.m3 = w.m[3];
    // forward statistics (three LayerNorm levels)
    head_kernel<0><<<w.gx, 128, 0, st>>>(hb.fwd, feat, nullptr, (int)N); URNN_LAUNCH_CHECK();
    head_kernel<1><<<w.gx, 128, 0, st>>>(hb.fwd, feat, nullptr, (int)N); URNN_LAUNCH_CHECK();
    head_kernel<2><<<w.gx, 128, 0, st>>>(hb.fwd, feat, nullptr, (int)N); URNN_LAUNCH_CHECK();
    // backward sweeps
    head_bwd_kernel<1><<<w.gx, 128, 0, st>>>(hb, feat, (int)N); URNN_LAUNCH_CHECK();
    if (comm.world > 1) { comm_sum2_kernel<<<1, 64, 0, st>>>(w.bsum + 2 * 3, 2, comm); URNN_LAUNCH_CHECK(); }      // LayerNorm sets 3, 4
    head_pred_grads_kernel<<<1, 32, 0, st>>>(w.psum, *g); URNN_LAUNCH_CHECK();
    head_bwd_kernel<2><<<w.gx, 128, 0, st>>>(hb, feat, (int)N); URNN_LAUNCH_CHECK();
    if (comm.world > 1) { comm_sum2_kernel<<<1, 64, 0, st>>>(w.bsum + 2 * 1, 2, comm); URNN_LAUNCH_CHECK(); }      // sets 1, 2
    URNN_TRY(launch_wgrad(w.m[0], N, 16, 16, (int)N, single_map_loader(w.m[1], 16, N), g->conv_w[2], 16, 1, nullptr, 0, st));   // cls_convs.1
    URNN_TRY(launch_wgrad(w.m[2], N, 16, 16, (int)N, single_map_loader(w.m[3], 16, N), g->conv_w[4], 16, 1, nullptr, 0, st));   // reg_convs.1
    head_bwd_kernel<3><<<w.gx, 128, 0, st>>>(hb, feat, (int)N); URNN_LAUNCH_CHECK();
    if (comm.world > 1) { comm_sum2_kernel<<<1, 64, 0, st>>>(w.bsum, 1, comm); URNN_LAUNCH_CHECK(); }               // set 0
    URNN_TRY(launch_wgrad(w.m[0], N, 16, 16, (int)N, single_map_loader(w.m[1], 16, N), g->conv_w[1], 16, 1, nullptr, 0, st));   // cls_convs.0
    URNN_TRY(launch_wgrad(w.m[2], N, 16, 16, (int)N, single_map_loader(w.m[1], 16, N), g->conv_w[3], 16, 1, nullptr, 0, st));   // reg_convs.0
    head_bwd_kernel<4><<<w.gx, 128, 0, st>>>(hb, feat, (int)N); URNN_LAUNCH_CHECK();
    URNN_TRY(launch_wgrad(w.m[0], N, 16, 16, (int)N, single_map_loader(feat, 16, N), g->conv_w[0], 16, 1, nullptr, 0, st));     // stems
    return URNN_OK;
}

}  // namespace urnn
