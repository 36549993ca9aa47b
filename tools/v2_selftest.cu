// v2_selftest.cu -- bring-up harness for u-rnn_b200/csrc/gemm_v2.cuh (no torch, no Python): random split maps, every launch
// shape the step uses, results compared with a double-precision host computation on the same hi+lo operand values.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/bin/v2_selftest tools/v2_selftest.cu
// Run (GPU box): timeout 120 tools/bin/v2_selftest
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../u-rnn_b200/csrc/v2_host.cuh"

namespace urnn {
static char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
void count_launch(int) {}
}  // namespace urnn
using namespace urnn;
using namespace urnn::v2;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define OK(x) do { int r_ = (x); if (r_ != 0) { printf("error %d: %s (%s:%d)\n", r_, urnn::g_err, __FILE__, __LINE__); exit(2); } } while (0)

static uint32_t rng = 12345;
static float frand() { rng = rng * 1664525u + 1013904223u; return ((rng >> 8) & 0xFFFF) / 65536.0f * 2.f - 1.f; }
// the 16-bit element format of the split maps (fp16 unless URNN_SPLIT_BF16), emulated on the host
#ifdef URNN_SPLIT_BF16
static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16; return (uint16_t)u; }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
#else
static uint16_t f2bf(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float bf2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
#endif

struct HostMap {          // split map on the host + its exact values (hi + lo)
    int C; long long ntot; std::vector<uint16_t> raw; std::vector<double> val;
    sp16* dev = nullptr;
    void init(int C_, long long n_, long long blk_stride, long long blk_valid, float scale) {
        C = C_; ntot = n_; raw.assign((size_t)2 * C * ntot, 0); val.assign((size_t)C * ntot, 0.0);
        for (int c = 0; c < C; ++c)
            for (long long p = 0; p < ntot; ++p) {
                if ((p % blk_stride) >= blk_valid) continue;          // padding stays zero
                float v = frand() * scale + 0.3f * scale;
                uint16_t h = f2bf(v), l = f2bf(v - bf2f(h));
                raw[(size_t)c * ntot + p] = h; raw[(size_t)(C + c) * ntot + p] = l;
                val[(size_t)c * ntot + p] = (double)bf2f(h) + (double)bf2f(l);
            }
        CK(cudaMalloc(&dev, raw.size() * 2));
        CK(cudaMemcpy(dev, raw.data(), raw.size() * 2, cudaMemcpyHostToDevice));
    }
    SplitMap sm() const { return SplitMap{dev, C, ntot}; }
};

static double wval(float w) { uint16_t h = f2bf(w), l = f2bf(w - bf2f(h)); return (double)bf2f(h) + (double)bf2f(l); }

struct Case {
    const char* name;
    int nseg; int segc[3];
    int gate_ch;              // gated extra segment
    int N, nacc, acc_mode, epi;
    long long ntot, blk_stride, blk_valid;
    int nstat;
};

static int run_case(const Case& cs, int num_sms) {
    printf("== %s: N=%d nacc=%d ntot=%lld\n", cs.name, cs.N, cs.nacc, cs.ntot); fflush(stdout);
    const bool pool = cs.acc_mode == ACC_POOL, deconv = cs.acc_mode == ACC_DECONV;
    // source maps; pooled sources are 4x larger (four phase blocks)
    const long long src_ntot = pool ? 4 * cs.ntot : cs.ntot;
    HostMap seg[3];
    int Kt = 0;
    for (int i = 0; i < cs.nseg; ++i) { seg[i].init(cs.segc[i], src_ntot, cs.blk_stride, cs.blk_valid, 1.0f); Kt += cs.segc[i]; }
    HostMap hmap; std::vector<float> gpre, gsc, gsh;
    float *d_gpre = nullptr, *d_gsc = nullptr, *d_gsh = nullptr;
    if (cs.gate_ch) {
        hmap.init(cs.gate_ch, cs.ntot, cs.blk_stride, cs.blk_valid, 1.0f);
        gpre.resize((size_t)cs.gate_ch * cs.ntot); gsc.resize(cs.gate_ch); gsh.resize(cs.gate_ch);
        for (auto& v : gpre) v = frand() * 2.f;
        for (int c = 0; c < cs.gate_ch; ++c) { gsc[c] = 0.5f + 0.5f * fabsf(frand()); gsh[c] = 0.2f * frand(); }
        CK(cudaMalloc(&d_gpre, gpre.size() * 4)); CK(cudaMemcpy(d_gpre, gpre.data(), gpre.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_gsc, cs.gate_ch * 4)); CK(cudaMemcpy(d_gsc, gsc.data(), cs.gate_ch * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_gsh, cs.gate_ch * 4)); CK(cudaMemcpy(d_gsh, gsh.data(), cs.gate_ch * 4, cudaMemcpyHostToDevice));
    }
    const int K = Kt + cs.gate_ch;
    const int nrows = deconv ? cs.nacc * cs.N : cs.N;
    // weights: plain [nrows][K]
    std::vector<float> W((size_t)nrows * K), bias(nrows);
    for (auto& v : W) v = frand() * 0.2f;
    for (auto& v : bias) v = frand() * 0.5f;
    float *d_W, *d_b; char* d_img;
    CK(cudaMalloc(&d_W, W.size() * 4)); CK(cudaMemcpy(d_W, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_b, bias.size() * 4)); CK(cudaMemcpy(d_b, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_img, wimg_bytes(nrows, K)));
    WImgBatch B; memset(&B, 0, sizeof(B));
    B.n = 1; B.base = d_img;
    B.s[0] = WImgSpec{d_W, K, 1, 0, 1 << 30, nrows, nrows, K, 0, 0};
    wimg_kernel<<<dim3(8, 1), 256>>>(B);
    CK(cudaGetLastError());

    GemmLaunch L; memset(&L, 0, sizeof(L)); params_defaults(L.P);
    GemmParams& P = L.P;
    P.acc_mode = cs.acc_mode; P.nacc = cs.nacc;
    int k = 0;
    if (pool) {
        for (int a = 0; a < 4; ++a) { k = 0; for (int i = 0; i < cs.nseg; ++i) k = add_segment_steps(L, i, cs.segc[i], k, a, a * cs.ntot); }
    } else {
        for (int i = 0; i < cs.nseg; ++i) k = add_segment_steps(L, i, cs.segc[i], k, 0, 0);
    }
    for (int i = 0; i < 3; ++i) {
        const HostMap& m = seg[i < cs.nseg ? i : 0];
        OK(make_split_tmap(&L.maps[i], m.sm(), (m.C % 32 == 0) ? 32 : 16, 2));
    }
    if (cs.gate_ch) {
        P.gate_ch = cs.gate_ch; P.gate_k0 = Kt; P.gate_h = hmap.dev; P.gate_h_plane = cs.ntot; P.gate_h_lo = (long long)cs.gate_ch * cs.ntot;
        P.gate_pre = d_gpre; P.gate_pre_plane = cs.ntot; P.gate_scale = d_gsc; P.gate_shift = d_gsh;
    }
    P.wimg = d_img; P.nkb = (K + 63) / 64; P.nrows = nrows; P.N = cs.N; P.nmma = 3;
    P.ntot = cs.ntot; P.blk_stride = cs.blk_stride; P.blk_valid = cs.blk_valid;
    P.epi = cs.epi; P.slope = 0.2f; P.bias = d_b; P.nbias = nrows; P.bias_mod = 1 << 30;
    const long long out_ntot = deconv ? cs.nacc * cs.ntot : cs.ntot;
    float* d_out = nullptr; sp16* d_split = nullptr;
    const int outC = cs.N;
    if (cs.epi == EPI_LRELU_SPLIT) {
        CK(cudaMalloc(&d_split, SplitMap::bytes(outC, out_ntot))); CK(cudaMemset(d_split, 0, SplitMap::bytes(outC, out_ntot)));
        P.out_hi = d_split; P.out_lo = (long long)outC * out_ntot; P.out_plane = out_ntot; P.out_acc_stride = cs.ntot;
    } else {
        CK(cudaMalloc(&d_out, (size_t)outC * out_ntot * 4)); CK(cudaMemset(d_out, 0, (size_t)outC * out_ntot * 4));
        P.out_f32 = d_out; P.out_plane = out_ntot; P.store_c0 = 0; P.store_c1 = getenv("V2_NOSTORE") ? 0 : cs.N;
    }
    // statistics
    float4* d_part = nullptr; double* d_tot = nullptr; unsigned* d_cnt = nullptr; float *d_scale = nullptr, *d_shift = nullptr, *d_gamma = nullptr, *d_beta = nullptr;
    std::vector<float> gamma(cs.N, 1.f), beta(cs.N, 0.f);
    if (cs.nstat) {
        for (int c = 0; c < cs.N; ++c) { gamma[c] = 1.f + 0.1f * frand(); beta[c] = 0.1f * frand(); }
        CK(cudaMalloc(&d_part, (size_t)cs.nstat * NWARP_EPI * num_sms * sizeof(float4)));
        CK(cudaMalloc(&d_tot, cs.nstat * 4 * sizeof(double))); CK(cudaMalloc(&d_cnt, 4)); CK(cudaMemset(d_cnt, 0, 4));
        CK(cudaMalloc(&d_scale, cs.N * 4)); CK(cudaMalloc(&d_shift, cs.N * 4));
        CK(cudaMalloc(&d_gamma, cs.N * 4)); CK(cudaMalloc(&d_beta, cs.N * 4));
        CK(cudaMemcpy(d_gamma, gamma.data(), cs.N * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_beta, beta.data(), cs.N * 4, cudaMemcpyHostToDevice));
        P.nstat = cs.nstat;
        P.sink.partial = d_part; P.sink.total = d_tot; P.sink.counter = d_cnt; P.sink.nsets = cs.nstat; P.sink.stride = NWARP_EPI * num_sms;
        P.aff.scale = d_scale; P.aff.shift = d_shift; P.aff.gamma = d_gamma; P.aff.beta = d_beta; P.aff.channels = cs.nstat * 32; P.aff.ch_per_set = 32; P.aff.eps = 1e-5f;
    }
    OK(plan_gemm(L, num_sms));
    long long* d_dbg = nullptr;
    if (getenv("V2_TRACE")) { CK(cudaMalloc(&d_dbg, (9 * 32 + 768) * 8)); CK(cudaMemset(d_dbg, 0, (9 * 32 + 768) * 8)); }
    printf("   grid %d smem %zu nslots %d gdepth %d tmem %d stages %d nsteps %d\n", L.grid, L.smem, P.nslots, P.gdepth, P.tmem_cols, P.acc_stages, P.nsteps); fflush(stdout);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    OK(launch_gemm(L, 0, false));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int it = 0; it < 5; ++it) OK(launch_gemm(L, 0, false));
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double in_bytes = 0; for (int i = 0; i < cs.nseg; ++i) in_bytes += 4.0 * cs.segc[i] * src_ntot;
    if (cs.gate_ch) in_bytes += 8.0 * cs.gate_ch * cs.ntot;
    const double out_bytes = 4.0 * outC * out_ntot;
    printf("   %.1f us per launch, %.0f GB/s (in %.1f MB + out %.1f MB)\n", ms * 200.0, (in_bytes + out_bytes) / (ms / 5 * 1e-3) * 1e-9, in_bytes * 1e-6, out_bytes * 1e-6);

    if (d_dbg) {
        P.dbg = d_dbg; OK(launch_gemm(L, 0, false)); CK(cudaDeviceSynchronize()); P.dbg = nullptr;
        long long h[9 * 32 + 768]; CK(cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
        {
            long long tmin = 1LL << 62; for (int b = 0; b < L.grid; ++b) if (h[9 * 32 + b] < tmin) tmin = h[9 * 32 + b];
            printf("   per CTA (start us, end us, smid):");
            for (int b = 0; b < L.grid; b += 7) printf(" [%d: %.1f %.1f sm%lld]", b, (h[9 * 32 + b] - tmin) * 1e-3, (h[9 * 32 + 256 + b] - tmin) * 1e-3, h[9 * 32 + 512 + b]);
            double emax = 0, emin = 1e30; for (int b = 0; b < L.grid; ++b) { double e = (h[9 * 32 + 256 + b] - tmin) * 1e-3; if (e > emax) emax = e; if (e < emin) emin = e; }
            printf("\n   CTA end times: min %.1f us max %.1f us\n", emin, emax);
            printf("   first CTA entry %.1f us | last CTA past its partials %.1f us | last CTA exit (after finalize) %.1f us\n",
                   ((long long)~(unsigned long long)h[8 * 32 + 1] - tmin) * 1e-3, (h[8 * 32 + 2] - tmin) * 1e-3, (h[8 * 32 + 3] - tmin) * 1e-3);
        }
        const long long t0 = h[8 * 32];
        printf("   timeline of CTA 0 (cycles since setup done): tile: tma_first tma_last | mma_tempty mma_full0 mma_fullN mma_commit | epi_wake epi_done\n");
        for (int t = 0; t < 14; ++t) {
            printf("   %2d:", t);
            for (int e = 0; e < 8; ++e) printf(" %7lld%s", h[e * 32 + t] ? h[e * 32 + t] - t0 : -1, (e == 1 || e == 5) ? " |" : "");
            printf("\n");
        }
    }
    // ---- host reference on a sample of pixels
    std::vector<float> out_h; std::vector<uint16_t> split_h;
    if (d_out) { out_h.resize((size_t)outC * out_ntot); CK(cudaMemcpy(out_h.data(), d_out, out_h.size() * 4, cudaMemcpyDeviceToHost)); }
    if (d_split) { split_h.resize((size_t)2 * outC * out_ntot); CK(cudaMemcpy(split_h.data(), d_split, split_h.size() * 2, cudaMemcpyDeviceToHost)); }
    auto got = [&](int c, long long p) -> double {
        if (d_out) return out_h[(size_t)c * out_ntot + p];
        return (double)bf2f(split_h[(size_t)c * out_ntot + p]) + (double)bf2f(split_h[(size_t)(outC + c) * out_ntot + p]);
    };
    auto acc_of = [&](int a, int n, long long p) -> double {       // pre-activation of accumulator a, column n, pixel p
        double s = bias[deconv ? a * cs.N + n : n];
        const int row = deconv ? a * cs.N + n : n;
        int kk = 0;
        for (int i = 0; i < cs.nseg; ++i)
            for (int c = 0; c < cs.segc[i]; ++c, ++kk) s += wval(W[(size_t)row * K + kk]) * seg[i].val[(size_t)c * src_ntot + p + (pool ? a * cs.ntot : 0)];
        for (int c = 0; c < cs.gate_ch; ++c, ++kk) {
            const double r = 1.0 / (1.0 + exp(-((double)gpre[(size_t)c * cs.ntot + p] * gsc[c] + gsh[c])));
            float gv = (float)(hmap.val[(size_t)c * cs.ntot + p] * r);
            uint16_t h = f2bf(gv), l = f2bf(gv - bf2f(h));
            s += wval(W[(size_t)row * K + kk]) * ((double)bf2f(h) + (double)bf2f(l));
        }
        return s;
    };
    double maxerr = 0, maxref = 0; long long nbad = 0, nchk = 0, npadbad = 0;
    const long long stepp = cs.ntot > 100000 ? 997 : (cs.ntot > 4096 ? 37 : 1);
    for (long long p = 0; p < cs.ntot; p += stepp) {
        const bool valid = (p % cs.blk_stride) < cs.blk_valid;
        for (int n = 0; n < cs.N; ++n) {
            for (int a = 0; a < (deconv ? cs.nacc : 1); ++a) {
                double ref;
                if (cs.epi == EPI_STATS_F32) ref = acc_of(0, n, p);
                else if (pool) { ref = 0; for (int q = 0; q < 4; ++q) { double v = acc_of(q, n, p); ref += 0.25 * (v >= 0 ? v : 0.2 * v); } }
                else { double v = acc_of(a, n, p); ref = v >= 0 ? v : 0.2 * v; }
                const double g = got(n, p + (deconv ? a * cs.ntot : 0));
                if (!valid) { if (g != 0.0) ++npadbad; continue; }
                const double e = fabs(g - ref);
                ++nchk;
                if (e > maxerr) maxerr = e;
                if (fabs(ref) > maxref) maxref = fabs(ref);
                if (e > 2e-5 * (1.0 + fabs(ref))) { if (nbad < 5) printf("   MISMATCH p=%lld n=%d a=%d got %.7f ref %.7f\n", p, n, a, g, ref); ++nbad; }
            }
        }
    }
    printf("   checked %lld values: max |err| %.3e (max |ref| %.3f), bad %lld, nonzero padding %lld\n", nchk, maxerr, maxref, nbad, npadbad);
    int fail = (nbad > 0 || npadbad > 0);
    if (cs.nstat) {
        std::vector<float> sc(cs.N), sh(cs.N);
        CK(cudaMemcpy(sc.data(), d_scale, cs.N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(sh.data(), d_shift, cs.N * 4, cudaMemcpyDeviceToHost));
        // reference statistics from the device's own stored fp32 values (all valid pixels)
        double worst = 0;
        for (int g = 0; g < cs.nstat; ++g) {
            double s = 0, q = 0, n = 0;
            for (int c = g * 32; c < g * 32 + 32; ++c)
                for (long long p = 0; p < cs.ntot; ++p) if ((p % cs.blk_stride) < cs.blk_valid) { double v = out_h[(size_t)c * out_ntot + p]; s += v; q += v * v; n += 1; }
            const double mean = s / n, var = q / n - mean * mean, rstd = 1.0 / sqrt(var + 1e-5);
            for (int c = g * 32; c < g * 32 + 32; ++c) {
                const double esc = gamma[c] * rstd, esh = beta[c] - mean * esc;
                worst = fmax(worst, fabs(sc[c] - esc) / fabs(esc)); worst = fmax(worst, fabs(sh[c] - esh));
            }
        }
        printf("   GroupNorm affine: worst deviation %.3e\n", worst);
        if (!(worst < 1e-4)) fail = 1;
    }
    printf("   %s\n", fail ? "FAIL" : "ok"); fflush(stdout);
    return fail;
}

int main(int argc, char** argv) {
    int dev = 0, num_sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    printf("SMs: %d\n", num_sms);
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    const long long big = 148LL * 3 * 128 + 5 * 128;      // several tiles per CTA, uneven
    Case cases[] = {
        {"enc1 sweep A (x16 | h64 -> 128, stats)", 2, {16, 64, 0}, 0, 128, 1, ACC_SINGLE, EPI_STATS_F32, 6 * 128, 384, 300, 4},
        {"dec1 sweep A (x96 | e64 | h64 -> 128, stats), many tiles", 3, {96, 64, 64}, 0, 128, 1, ACC_SINGLE, EPI_STATS_F32, big, big, big - 77, 4},
        {"dec1 sweep B gated (x96 | e64 | r*h64 -> 64, stats)", 2, {96, 64, 0}, 64, 64, 1, ACC_SINGLE, EPI_STATS_F32, big, big, big - 77, 2},
        {"enc3 sweep A (x96 | h96 -> 192, stats)", 2, {96, 96, 0}, 0, 192, 1, ACC_SINGLE, EPI_STATS_F32, 123 * 128, 123 * 128, 15625, 6},
        {"dec2 sweep B gated (x96 | e96 | r*h96 -> 96)", 2, {96, 96, 0}, 96, 96, 1, ACC_SINGLE, EPI_STATS_F32, 50 * 128, 25 * 128, 3100, 3},
        {"pooled stem 64 -> 64 (4 phases)", 1, {64, 0, 0}, 0, 64, 4, ACC_POOL, EPI_LRELU_SPLIT, 40 * 128, 20 * 128, 2501, 0},
        {"pooled stem 96 -> 96 (4 phases)", 1, {96, 0, 0}, 0, 96, 4, ACC_POOL, EPI_LRELU_SPLIT, 123 * 128, 123 * 128, 15625, 0},
        {"deconv half 96 -> 2 x 96", 1, {96, 0, 0}, 0, 96, 2, ACC_DECONV, EPI_LRELU_SPLIT, 123 * 128, 123 * 128, 15625, 0},
        {"final stem 64 -> 16 fp32", 1, {64, 0, 0}, 0, 16, 1, ACC_SINGLE, EPI_LRELU_F32, big, big, big - 3, 0},
        {"plain stem 96 -> 96 split", 1, {96, 0, 0}, 0, 96, 1, ACC_SINGLE, EPI_LRELU_SPLIT, 30 * 128, 30 * 128, 30 * 128 - 1, 0},
    };
    const long long full = 16LL * 123 * 128;               // location1 grid in the phase-separated layout
    Case perf[] = {
        {"PERF enc1 sweep A", 2, {16, 64, 0}, 0, 128, 1, ACC_SINGLE, EPI_STATS_F32, full, 123 * 128, 15625, 4},
        {"PERF enc1 sweep B gated", 1, {16, 0, 0}, 64, 64, 1, ACC_SINGLE, EPI_STATS_F32, full, 123 * 128, 15625, 2},
        {"PERF dec1 sweep A", 3, {96, 64, 64}, 0, 128, 1, ACC_SINGLE, EPI_STATS_F32, full, 123 * 128, 15625, 4},
        {"PERF dec1 sweep B gated", 2, {96, 64, 0}, 64, 64, 1, ACC_SINGLE, EPI_STATS_F32, full, 123 * 128, 15625, 2},
        {"PERF stage-2 pooled stem 64 -> 64", 1, {64, 0, 0}, 0, 64, 4, ACC_POOL, EPI_LRELU_SPLIT, full / 4, 123 * 128, 15625, 0},
        {"PERF deconv half 96 -> 2 x 96 (half res -> full res)", 1, {96, 0, 0}, 0, 96, 2, ACC_DECONV, EPI_LRELU_SPLIT, full / 4, 123 * 128, 15625, 0},
        {"PERF final stem 64 -> 16", 1, {64, 0, 0}, 0, 16, 1, ACC_SINGLE, EPI_LRELU_F32, full, 123 * 128, 15625, 0},
    };
    if (argc > 1 && !strcmp(argv[1], "perf")) {
        int f = 0;
        int pi = 0;
        for (const Case& c : perf) { if (argc < 3 || atoi(argv[2]) == pi) f += run_case(c, num_sms); ++pi; }
        return f ? 1 : 0;
    }
    int fails = 0, i = 0;
    for (const Case& c : cases) { if (only < 0 || only == i) fails += run_case(c, num_sms); ++i; }
    printf("%s (%d failing cases)\n", fails ? "SELFTEST FAILED" : "SELFTEST PASSED", fails);
    return fails ? 1 : 0;
}
