// urnn_v2.cu -- URNN_MATH_F16X3: the encoder-decoder time step on the second-generation tcgen05 pixel GEMM (gemm_v2.cuh).
//
// Arithmetic: every GEMM operand is an fp16 hi + lo pair (22 significant bits on both sides, three MMAs per K=16 step),
// accumulation / GroupNorm / LayerNorm statistics / gates / blend in fp32, pre-norm maps stored as fp32.  This is the mode
// that stays inside the config-3 tolerance at T = 180 (tests/test_gpu_x3.py); single-pass bf16 does not, and bf16 hi+lo pairs miss the max |d state| bound on the full grid.
//
// Internal data layout ("phase-separated"): an H x W map (H, W multiples of 4) is stored as 16
// blocks (full resolution), 4 blocks (half) or 1 block (quarter) of n4p = roundup((H/4)*(W/4), 128) pixels:
//     full-res (y, x)  ->  block ((y&1)*2 + (x&1))*4 + (((y>>1)&1)*2 + ((x>>1)&1)),  position (y>>2)*(W/4) + (x>>2)
// so that the 2x2 children of a coarse pixel sit at the SAME position of four consecutive block groups:
// AvgPool2 (utils.py:92-94) and ConvTranspose2d(k2,s2) (utils.py:95-100) become pure channel operations on whole tiles
// (four accumulators in tensor memory), with plain TMA boxes and plain coalesced stores.  Activations that feed GEMMs
// are "split maps" [2 (hi|lo)][C][ntot] fp16.  The reference's NCHW fp32 tensors exist only at the API boundary.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <memory>
#include <mutex>

#include "v2_host.cuh"
#include "urnn_internal.h"

namespace urnn {
namespace v2 {

// ------------------------------------------------------------------------------------------------ layout
struct Layout {
    int H, W, phased;            // phased = 0: plain layout (stand-alone cell): pixel y*W + x, one block
    long long n4, n4p;           // quarter-resolution pixels, padded to whole tiles
    long long ntot[3];           // pixels per plane at full / half / quarter resolution (padded)
    long long blk_stride, blk_valid;
};
static Layout make_layout(int H, int W, bool phased) {
    Layout l; l.H = H; l.W = W; l.phased = phased ? 1 : 0;
    if (phased) {
        l.n4 = (long long)(H / 4) * (W / 4); l.n4p = (l.n4 + 127) / 128 * 128;
        l.ntot[0] = 16 * l.n4p; l.ntot[1] = 4 * l.n4p; l.ntot[2] = l.n4p;
        l.blk_stride = l.n4p; l.blk_valid = l.n4;
    } else {
        l.n4 = (long long)H * W; l.n4p = (l.n4 + 127) / 128 * 128;
        l.ntot[0] = l.ntot[1] = l.ntot[2] = l.n4p;
        l.blk_stride = l.n4p; l.blk_valid = l.n4;
    }
    return l;
}
// position of pixel (y, x) of a level-`level` map (h = H >> level rows) in the internal order
__host__ __device__ static inline long long ix_internal(int phased, int level, int y, int x, int w_level, int w4, long long n4p) {
    if (!phased) return (long long)y * w_level + x;
    if (level == 0) return (long long)((((y & 1) * 2 + (x & 1)) * 4) + (((y >> 1) & 1) * 2 + ((x >> 1) & 1))) * n4p + (long long)(y >> 2) * w4 + (x >> 2);
    if (level == 1) return (long long)((y & 1) * 2 + (x & 1)) * n4p + (long long)(y >> 1) * w4 + (x >> 1);
    return (long long)y * w4 + x;
}

// ------------------------------------------------------------------------------------------------ elementwise kernels
// NCHW fp32 -> split map (API boundary, once per sequence / per stand-alone call)
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, sp16* __restrict__ hi, long long lo_off,
                                                   int C, int h, int w, int phased, int level, int w4, long long n4p, long long ntot) {
    const long long total = (long long)C * h * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w); const long long r = i / w; const int y = (int)(r % h); const int c = (int)(r / h);
        const uint32_t s = split16(__ldg(src + i));
        const long long o = (long long)c * ntot + ix_internal(phased, level, y, x, w, w4, n4p);
        reinterpret_cast<unsigned short*>(hi)[o] = (unsigned short)(s & 0xFFFFu);
        reinterpret_cast<unsigned short*>(hi)[lo_off + o] = (unsigned short)(s >> 16);
    }
}
// split map -> NCHW fp32
__global__ void __launch_bounds__(256) unpack_kernel(const sp16* __restrict__ hi, long long lo_off, float* __restrict__ dst,
                                                     int C, int h, int w, int phased, int level, int w4, long long n4p, long long ntot) {
    const long long total = (long long)C * h * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w); const long long r = i / w; const int y = (int)(r % h); const int c = (int)(r / h);
        const long long o = (long long)c * ntot + ix_internal(phased, level, y, x, w, w4, n4p);
        dst[i] = lo16_to_f32(reinterpret_cast<const unsigned short*>(hi)[o]) + lo16_to_f32(reinterpret_cast<const unsigned short*>(hi)[lo_off + o]);
    }
}

// h' = (1-z) h + z tanh(GN2(C)),  z = sigmoid(GN1(G_z))  (ConvRNN.py:160-162,180,185,189).  G_z, C: fp32 [F][ntot];
// h, h': split maps.  One thread = 8 consecutive pixels of one channel (16/32-byte vectors).
struct BlendArgs {
    const float* Gz; const float* C; const sp16* h; long long h_lo; sp16* ho; long long ho_lo;
    const float *sc1, *sh1, *sc2, *sh2; int F; long long ntot;
};
__global__ void __launch_bounds__(256) blend_kernel(const BlendArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const long long nvec = a.ntot >> 3, total = nvec * a.F;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx / nvec);
    const long long e = (long long)c * a.ntot + (idx - (long long)c * nvec) * 8;
    const float4 g0 = __ldcs(reinterpret_cast<const float4*>(a.Gz + e)), g1 = __ldcs(reinterpret_cast<const float4*>(a.Gz + e) + 1);
    const float4 c0 = __ldcs(reinterpret_cast<const float4*>(a.C + e)), c1 = __ldcs(reinterpret_cast<const float4*>(a.C + e) + 1);
    const uint4 hh = __ldcs(reinterpret_cast<const uint4*>(a.h + e)), hl = __ldcs(reinterpret_cast<const uint4*>(a.h + a.h_lo + e));
    const float a1 = __ldg(a.sc1 + c), b1 = __ldg(a.sh1 + c), a2 = __ldg(a.sc2 + c), b2 = __ldg(a.sh2 + c);
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const uint32_t hw[4] = {hh.x, hh.y, hh.z, hh.w}, lw[4] = {hl.x, hl.y, hl.z, hl.w};
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float h0 = lo16_to_f32(hw[u]) + lo16_to_f32(lw[u]), h1 = hi16_to_f32(hw[u]) + hi16_to_f32(lw[u]);
        const float z0 = sigmoid_fast(fmaf(gv[2 * u], a1, b1)), z1 = sigmoid_fast(fmaf(gv[2 * u + 1], a1, b1));
        const float t0 = tanh_fast(fmaf(cv[2 * u], a2, b2)), t1 = tanh_fast(fmaf(cv[2 * u + 1], a2, b2));
        split16x2(fmaf(z0, t0 - h0, h0), fmaf(z1, t1 - h1, h1), oh[u], ol[u]);
    }
    *reinterpret_cast<uint4*>(a.ho + e) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4*>(a.ho + a.ho_lo + e) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
}

// Encoder stage 1 (encoder.py:142-157): y = LeakyReLU(W x + b), x = the step's NCHW fp32 input, y -> split map in the
// internal layout.  fp32 FFMA (C_in = 63 or 3 channels: HBM-bound).  One thread = one image row of a 4x4 patch (4 pixels,
// one 16-byte load per channel), up to 16 output channels per pass.
struct Stem1Args {
    const float* x; int cin; const float* w; long long w_ld; const float* b; int cout; float slope;
    int H, W, w4, phased; long long n4p, ntot;
    sp16* out; long long out_lo;
};
__global__ void __launch_bounds__(64) stem1_kernel(const Stem1Args a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(16) float sw[];           // weights transposed [cin][16 * npass] (one 64-byte row per channel and pass), then bias
    const int npass = (a.cout + 15) / 16, ldw = 16 * npass;
    for (int i = threadIdx.x; i < a.cin * ldw; i += blockDim.x) {
        const int c = i / ldw, o = i % ldw;
        sw[i] = o < a.cout ? __ldg(a.w + (long long)o * a.w_ld + c) : 0.f;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");    // the per-step bias of the event mode is written by an earlier kernel
    for (int i = threadIdx.x; i < ldw; i += blockDim.x) sw[a.cin * ldw + i] = i < a.cout ? __ldg(a.b + i) : 0.f;
    __syncthreads();
    const long long nthr = (long long)a.H * a.w4;         // threads: (y, qx)
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthr) return;
    const int qx = (int)(t % a.w4), y = (int)(t / a.w4);
    const long long plane = (long long)a.H * a.W;
    const float* xp = a.x + (long long)y * a.W + 4 * qx;
    unsigned short* outp = reinterpret_cast<unsigned short*>(a.out);
    for (int ps = 0; ps < npass; ++ps) {
        float acc[16][4];
        const float* bias = sw + a.cin * ldw + 16 * ps;
#pragma unroll
        for (int o = 0; o < 16; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = bias[o];
        const float* wrow = sw + 16 * ps;
        int c = 0;
        for (; c + 8 <= a.cin; c += 8) {                  // eight independent 16-byte loads in flight per thread
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(xp + (long long)(c + u) * plane));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float4* w4p = reinterpret_cast<const float4*>(wrow + (c + u) * ldw);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = w4p[q];
                    const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int o = 4 * q + r;
                        acc[o][0] = fmaf(ww[r], v[u].x, acc[o][0]); acc[o][1] = fmaf(ww[r], v[u].y, acc[o][1]);
                        acc[o][2] = fmaf(ww[r], v[u].z, acc[o][2]); acc[o][3] = fmaf(ww[r], v[u].w, acc[o][3]);
                    }
                }
            }
        }
        for (; c < a.cin; ++c) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(xp + (long long)c * plane));
            const float4* w4p = reinterpret_cast<const float4*>(wrow + c * ldw);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 wv = w4p[q];
                const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int o = 4 * q + r;
                    acc[o][0] = fmaf(ww[r], v.x, acc[o][0]); acc[o][1] = fmaf(ww[r], v.y, acc[o][1]);
                    acc[o][2] = fmaf(ww[r], v.z, acc[o][2]); acc[o][3] = fmaf(ww[r], v.w, acc[o][3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long pi = ix_internal(a.phased, 0, y, 4 * qx + j, a.W, a.w4, a.n4p);
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                if (16 * ps + o < a.cout) {
                    const uint32_t sv = split16(lrelu(acc[o][j], a.slope));
                    const long long e = (long long)(16 * ps + o) * a.ntot + pi;
                    outp[e] = (unsigned short)(sv & 0xFFFFu);
                    outp[a.out_lo + e] = (unsigned short)(sv >> 16);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ plan
struct Op { int kind; GemmLaunch g; BlendArgs b; ChainLaunch c; char name[24]; };     // kind 0: GEMM, 1: blend, 2: chain (recompute) sweep
static inline const void*& op_wimg(Op& o) { return o.kind == 2 ? o.c.P.wimg : o.g.P.wimg; }
struct StatBufs { float4* partial; double* total; unsigned* counter; };

struct Plan {
    urnn_ed_desc d; urnn_ed_params p;
    Layout lay; int num_sms; int device;
    SplitMap st[2][6];                         // recurrent states, ping-pong (reference order: e1 e2 e3 d(1/4) d(1/2) d(1x))
    SplitMap s[3], up3, up2;                   // stem outputs, deconv outputs
    float *G, *C, *feat;                       // pre-GN maps (largest cell), decoder features (NCHW fp32)
    float *G2, *C2;                            // the decoder cells' own pre-GN maps (encoder of step t+1 overlaps decoder of step t)
    size_t n_enc[2];                           // ops[q][0 .. n_enc[q]) = encoder, the rest = decoder
    // two-stream software pipeline of a sequence call (single GPU): encoder(t+1) || decoder + head (t)
    bool want_pipe = false;                    // decided before build_plan (sharded + pipelined launches leave two SMs free)
    bool sharded = false;                      // a communicator with world > 1 was active when the plan was built
    bool pipe = false; cudaStream_t sE = nullptr, sD = nullptr; cudaEvent_t ev_call = nullptr, evE[2] = {nullptr, nullptr}, evD[2] = {nullptr, nullptr};
    ~Plan() {
        if (sE) cudaStreamDestroy(sE);
        if (sD) cudaStreamDestroy(sD);
        if (ev_call) cudaEventDestroy(ev_call);
        for (int i = 0; i < 2; ++i) { if (evE[i]) cudaEventDestroy(evE[i]); if (evD[i]) cudaEventDestroy(evD[i]); }
    }
    void* head_ws; size_t head_ws_bytes;
    char* wimg; size_t wimg_cap, wimg_used;
    WImgBatch wb;
    char* stat_base; size_t stat_bytes;
    std::vector<Op> ops[2];                    // the step after stage 1, before the head, per state parity
    size_t total;
    int lvl_of_state[6]; int ch_of_state[6];
    bool weights_ready;
};

static inline int state_level(int k) { const int sc[6] = {0, 1, 2, 2, 1, 0}; return sc[k]; }

struct ArenaW {                                // bump allocator (sizes only when base == nullptr)
    char* base; size_t off;
    template <class T> T* take(size_t n) { off = (off + 255) / 256 * 256; T* r = base ? (T*)(base + off) : nullptr; off += n * sizeof(T); return r; }
};

static SplitMap take_map(ArenaW& a, int C, long long ntot) { SplitMap m; m.C = C; m.ntot = ntot; m.hi = a.take<sp16>((size_t)2 * C * ntot); return m; }

static int layout_plan(Plan& pl, const urnn_ed_desc* d, void* ws) {
    const Layout& l = pl.lay;
    ArenaW a{(char*)ws, 0};
    pl.wimg_cap = (size_t)12 << 20; pl.wimg = a.take<char>(pl.wimg_cap);
    pl.stat_bytes = (size_t)1 << 20; pl.stat_base = a.take<char>(pl.stat_bytes);
    const int ch[6] = {d->enc_gru[0], d->enc_gru[1], d->enc_gru[2], d->dec_gru[0], d->dec_gru[1], d->dec_gru[2]};
    for (int k = 0; k < 6; ++k) { pl.lvl_of_state[k] = state_level(k); pl.ch_of_state[k] = ch[k]; }
    for (int q = 0; q < 2; ++q) for (int k = 0; k < 6; ++k) pl.st[q][k] = take_map(a, ch[k], l.ntot[state_level(k)]);
    for (int k = 0; k < 3; ++k) pl.s[k] = take_map(a, d->enc_conv[k], l.ntot[k]);
    pl.up3 = take_map(a, d->dec_conv[0], l.ntot[1]);
    pl.up2 = take_map(a, d->dec_conv[1], l.ntot[0]);
    size_t gmax = 0, cmax = 0;
    for (int k = 0; k < 6; ++k) { const size_t n = (size_t)l.ntot[state_level(k)]; gmax = std::max(gmax, 2 * (size_t)ch[k] * n); cmax = std::max(cmax, (size_t)ch[k] * n); }
    pl.G = a.take<float>(gmax); pl.C = a.take<float>(cmax);
    pl.G2 = a.take<float>(gmax); pl.C2 = a.take<float>(cmax);
    pl.feat = a.take<float>((size_t)d->dec_conv[2] * d->H * d->W);
    pl.head_ws_bytes = head_fwd_fp32_workspace(d->H, d->W);
    pl.head_ws = a.take<char>(pl.head_ws_bytes);
    pl.total = (a.off + 255) / 256 * 256;
    return URNN_OK;
}

struct StatAlloc { char* base; size_t off, cap; };
static char* sa_take(StatAlloc& sa, size_t bytes) { sa.off = (sa.off + 255) / 256 * 256; char* r = sa.base + sa.off; sa.off += bytes; return r; }
static bool take_stats(StatAlloc& sa, int nsets, int stride, StatBufs* sb) {
    sb->partial = (float4*)sa_take(sa, (size_t)nsets * stride * sizeof(float4));
    sb->total = (double*)sa_take(sa, (size_t)nsets * 4 * sizeof(double));
    sb->counter = (unsigned*)sa_take(sa, sizeof(unsigned));
    return sa.off <= sa.cap;
}
static bool take_affine(StatAlloc& sa, int nch, float** scale, float** shift) {
    *scale = (float*)sa_take(sa, (size_t)nch * sizeof(float)); *shift = (float*)sa_take(sa, (size_t)nch * sizeof(float));
    return sa.off <= sa.cap;
}

// registers the weight image of a launch; returns its device address
static const void* add_wimg(Plan& pl, const float* W, long long w_ld, long long w_ks, long long w_acc, int racc, int nrows, int nrows_valid, int K, int k_skip) {
    if (pl.wb.n >= WIMG_MAX) return nullptr;
    const size_t bytes = (wimg_bytes(nrows, K) + 255) / 256 * 256;
    if (pl.wimg_used + bytes > pl.wimg_cap) return nullptr;
    WImgSpec& S = pl.wb.s[pl.wb.n++];
    S.W = W; S.w_ld = w_ld; S.w_ks = w_ks; S.w_acc = w_acc; S.racc = racc; S.nrows = nrows; S.nrows_valid = nrows_valid; S.K = K; S.k_skip = k_skip;
    S.off = pl.wimg_used; pl.wimg_used += bytes;
    return pl.wimg + S.off;
}

static void fill_common(GemmParams& P, const Plan& pl, int level) {
    P.ntot = pl.lay.ntot[level]; P.blk_stride = pl.lay.blk_stride; P.blk_valid = pl.lay.blk_valid; P.nmma = 3; P.slope = pl.d.lrelu_slope;
}

static int build_cell_y(Plan& pl, std::vector<Op>& ops, StatAlloc& sa, const urnn_cell_desc& cd, const urnn_cell_params& cp, int level,
                        const SplitMap* x, const SplitMap* e, const SplitMap& h, const SplitMap& hout, bool first_parity, int lane);
static bool chain_fits(int F, int Keff);

// One (Skip-)ConvGRU cell step (ConvRNN.py:140-190) as sweep A (gates, GN-1 statistics), sweep B (candidate, GN-2
// statistics) and the blend.  x / e may be absent (decoder stage 3: x = None -> its weight columns are skipped; encoder: no e).
static int build_cell(Plan& pl, std::vector<Op>& ops, StatAlloc& sa, const urnn_cell_desc& cd, const urnn_cell_params& cp, int level,
                      const SplitMap* x, const SplitMap* e, const SplitMap& h, const SplitMap& hout, bool first_parity, float* Gmap, float* Cmap, int lane) {
    const int F = cd.F, Cx = cd.Cx;
    const int Ch = cd.variant == URNN_CELL_DECODER ? 2 * F : F;
    const int Ktot = Cx + Ch, kskip = x ? 0 : Cx, Keff = Ktot - kskip;
    const long long ntot = pl.lay.ntot[level];
    if (F % 32 || (x && x->C % 16) || Keff % 16) { set_error("cell(f16x3): channel counts must be multiples of 16 (Cx=%d F=%d)", Cx, F); return URNN_E_UNSUPPORTED; }
    if (chain_fits(F, Keff)) return build_cell_y(pl, ops, sa, cd, cp, level, x, e, h, hout, first_parity, lane);
    CommDev comm; current_comm(&comm, lane);
    // ---- sweep A: G = W1 u + b1 (2F channels), in chunks of output channels whose hi+lo weight image fits beside the ring
    const int nkb = (Keff + 63) / 64;
    int nchunk = 1;
    while ((size_t)2 * (2 * F / nchunk) * nkb * 128 > (size_t)150 * 1024) nchunk *= 2;      // N=2F: <= ~150 KB of weights
    if ((2 * F / nchunk) % 32) { set_error("cell(f16x3): cannot split %d gate channels", 2 * F); return URNN_E_UNSUPPORTED; }
    float *scale1 = nullptr, *shift1 = nullptr;
    if (!take_affine(sa, 2 * F, &scale1, &shift1)) { set_error("cell(f16x3): statistics arena too small"); return URNN_E_WORKSPACE; }
    for (int ci = 0; ci < nchunk; ++ci) {
        const int N = 2 * F / nchunk, n0 = ci * N;
        Op op; op.kind = 0; memset(&op.g, 0, sizeof(op.g)); params_defaults(op.g.P);
        GemmParams& P = op.g.P; fill_common(P, pl, level);
        int k = 0, m = 0;
        const SplitMap* segs[3] = {nullptr, nullptr, nullptr};
        if (x) { segs[m] = x; k = add_segment_steps(op.g, m, x->C, k, 0, 0); ++m; }
        if (e) { segs[m] = e; k = add_segment_steps(op.g, m, e->C, k, 0, 0); ++m; }
        segs[m] = &h; k = add_segment_steps(op.g, m, h.C, k, 0, 0); ++m;
        for (int i = 0; i < 3; ++i) { const SplitMap* sm = segs[i] ? segs[i] : &h; URNN_TRY(make_split_tmap(&op.g.maps[i], *sm, sm->C % 32 == 0 ? 32 : 16, 2)); }
        P.N = N; P.nrows = N; P.nkb = nkb;
        P.wimg = first_parity ? add_wimg(pl, cp.w1 + (long long)n0 * Ktot, Ktot, 1, 0, 1 << 30, N, N, Keff, kskip) : nullptr;
        P.bias = cp.b1 + n0; P.nbias = N;
        P.epi = EPI_STATS_F32; P.out_f32 = Gmap + (long long)n0 * ntot; P.out_plane = ntot; P.store_c0 = 0; P.store_c1 = N;
        StatBufs sb;
        if (!take_stats(sa, N / 32, pl.num_sms, &sb)) { set_error("cell(f16x3): statistics arena too small"); return URNN_E_WORKSPACE; }
        P.nstat = N / 32;
        P.sink.partial = sb.partial; P.sink.total = sb.total; P.sink.counter = sb.counter; P.sink.nsets = N / 32; P.sink.stride = pl.num_sms; P.sink.comm = comm;
        P.aff.scale = scale1 + n0; P.aff.shift = shift1 + n0; P.aff.gamma = cp.gn1_w + n0; P.aff.beta = cp.gn1_b + n0; P.aff.channels = N; P.aff.ch_per_set = 32; P.aff.eps = cd.eps;
        URNN_TRY(plan_gemm(op.g, pl.num_sms));
        ops.push_back(op);
    }
    // ---- sweep B: C = W2 [x | e | r*h] + b2,  r = sigmoid(GN1(G)[F:2F]) applied by the gate warps
    float *scale2 = nullptr, *shift2 = nullptr;
    {
        Op op; op.kind = 0; memset(&op.g, 0, sizeof(op.g)); params_defaults(op.g.P);
        GemmParams& P = op.g.P; fill_common(P, pl, level);
        int k = 0, m = 0;
        const SplitMap* segs[3] = {nullptr, nullptr, nullptr};
        if (x) { segs[m] = x; k = add_segment_steps(op.g, m, x->C, k, 0, 0); ++m; }
        if (e) { segs[m] = e; k = add_segment_steps(op.g, m, e->C, k, 0, 0); ++m; }
        for (int i = 0; i < 3; ++i) { const SplitMap* sm = segs[i] ? segs[i] : &h; URNN_TRY(make_split_tmap(&op.g.maps[i], *sm, sm->C % 32 == 0 ? 32 : 16, 2)); }
        P.gate_ch = F; P.gate_k0 = k; P.gate_h = h.hi; P.gate_h_plane = ntot; P.gate_h_lo = h.lo_off();
        P.gate_pre = Gmap + (long long)F * ntot; P.gate_pre_plane = ntot; P.gate_scale = scale1 + F; P.gate_shift = shift1 + F;
        P.N = F; P.nrows = F; P.nkb = nkb;
        P.wimg = first_parity ? add_wimg(pl, cp.w2, Ktot, 1, 0, 1 << 30, F, F, Keff, kskip) : nullptr;
        P.bias = cp.b2; P.nbias = F;
        P.epi = EPI_STATS_F32; P.out_f32 = Cmap; P.out_plane = ntot; P.store_c0 = 0; P.store_c1 = F;
        StatBufs sb;
        if (!take_stats(sa, F / 32, pl.num_sms, &sb) || !take_affine(sa, F, &scale2, &shift2)) { set_error("cell(f16x3): statistics arena too small"); return URNN_E_WORKSPACE; }
        P.nstat = F / 32;
        P.sink.partial = sb.partial; P.sink.total = sb.total; P.sink.counter = sb.counter; P.sink.nsets = F / 32; P.sink.stride = pl.num_sms; P.sink.comm = comm;
        P.aff.scale = scale2; P.aff.shift = shift2; P.aff.gamma = cp.gn2_w; P.aff.beta = cp.gn2_b; P.aff.channels = F; P.aff.ch_per_set = 32; P.aff.eps = cd.eps;
        URNN_TRY(plan_gemm(op.g, pl.num_sms));
        ops.push_back(op);
    }
    // ---- blend
    {
        Op op; op.kind = 1; memset(&op.g, 0, sizeof(op.g));
        op.b = BlendArgs{Gmap, Cmap, h.hi, h.lo_off(), hout.hi, hout.lo_off(), scale1, shift1, scale2, shift2, F, ntot};
        ops.push_back(op);
    }
    return URNN_OK;
}

// The same cell as three recompute sweeps (gemm_v2_chain.cuh): the pre-norm maps G and C stay in tensor memory.
//   A: G statistics (no store) -> GN-1 affine;  B': r*h from TMEM, C statistics -> GN-2 affine;  C': z, r, C again, blend, store h'.
// Used when [W1 ; W2] hi + lo (3F rows) fit in shared memory next to the gate buffers and a ring of >= 3 slots.
static bool chain_fits(int F, int Keff) {
    // Opt-in (URNN_V2_Y=1).  Measured on B200 at 500 x 500, encoder stage 1: A 39 + B' 60 + C' 91 = 190 us against 163 us for
    // the materialising schedule, although DRAM traffic drops from 672 MB to 276 MB: both chain sweeps are bound by SIMT issue
    // (23 / 45 instructions per element in the gate / blend stages at 0.6 IPC per scheduler), not by HBM.
    const char* e = getenv("URNN_V2_Y");
    if (!e || atoi(e) == 0) return false;
    if (F % 32 || 3 * F > 256) return false;
    const int nkb = (Keff + 63) / 64;
    return chain_slots(nkb, 3 * F, F, 2) >= 3 || chain_slots(nkb, 3 * F, F, 1) >= 3;
}

static int build_cell_y(Plan& pl, std::vector<Op>& ops, StatAlloc& sa, const urnn_cell_desc& cd, const urnn_cell_params& cp, int level,
                        const SplitMap* x, const SplitMap* e, const SplitMap& h, const SplitMap& hout, bool first_parity, int lane) {
    const int F = cd.F, Cx = cd.Cx;
    const int Ch = cd.variant == URNN_CELL_DECODER ? 2 * F : F;
    const int Ktot = Cx + Ch, kskip = x ? 0 : Cx, Keff = Ktot - kskip;
    const long long ntot = pl.lay.ntot[level];
    const int nkb = (Keff + 63) / 64;
    CommDev comm; current_comm(&comm, lane);
    float *scale1 = nullptr, *shift1 = nullptr, *scale2 = nullptr, *shift2 = nullptr;
    if (!take_affine(sa, 2 * F, &scale1, &shift1) || !take_affine(sa, F, &scale2, &shift2)) { set_error("cell(f16x3): statistics arena too small"); return URNN_E_WORKSPACE; }
    // ---- sweep A: statistics of G = W1 u + b1
    {
        Op op; op.kind = 0; memset(&op.g, 0, sizeof(op.g)); params_defaults(op.g.P);
        GemmParams& P = op.g.P; fill_common(P, pl, level);
        int k = 0, m = 0;
        const SplitMap* segs[3] = {nullptr, nullptr, nullptr};
        if (x) { segs[m] = x; k = add_segment_steps(op.g, m, x->C, k, 0, 0); ++m; }
        if (e) { segs[m] = e; k = add_segment_steps(op.g, m, e->C, k, 0, 0); ++m; }
        segs[m] = &h; k = add_segment_steps(op.g, m, h.C, k, 0, 0); ++m;
        for (int i = 0; i < 3; ++i) { const SplitMap* sm = segs[i] ? segs[i] : &h; URNN_TRY(make_split_tmap(&op.g.maps[i], *sm, sm->C % 32 == 0 ? 32 : 16, 2)); }
        const int N = 2 * F;
        P.N = N; P.nrows = N; P.nkb = nkb;
        P.wimg = first_parity ? add_wimg(pl, cp.w1, Ktot, 1, 0, 1 << 30, N, N, Keff, kskip) : nullptr;
        P.bias = cp.b1; P.nbias = N;
        P.epi = EPI_STATS_F32; P.out_f32 = nullptr; P.out_plane = ntot; P.store_c0 = 0; P.store_c1 = 0;      // nothing stored
        StatBufs sb;
        if (!take_stats(sa, N / 32, pl.num_sms, &sb)) { set_error("cell(f16x3): statistics arena too small"); return URNN_E_WORKSPACE; }
        P.nstat = N / 32;
        P.sink.partial = sb.partial; P.sink.total = sb.total; P.sink.counter = sb.counter; P.sink.nsets = N / 32; P.sink.stride = pl.num_sms; P.sink.comm = comm;
        P.aff.scale = scale1; P.aff.shift = shift1; P.aff.gamma = cp.gn1_w; P.aff.beta = cp.gn1_b; P.aff.channels = N; P.aff.ch_per_set = 32; P.aff.eps = cd.eps;
        URNN_TRY(plan_gemm(op.g, pl.num_sms));
        ops.push_back(op);
    }
    // ---- sweeps B' (final = 0: rows [W1_r ; W2]) and C' (final = 1: rows [W1_z ; W1_r ; W2])
    for (int fin = 0; fin < 2; ++fin) {
        Op op; op.kind = 2; memset(&op.g, 0, sizeof(op.g)); memset(&op.c, 0, sizeof(op.c));
        ChainParams& P = op.c.P;
        P.ntot = ntot; P.blk_stride = pl.lay.blk_stride; P.blk_valid = pl.lay.blk_valid;
        P.F = F; P.final = fin;
        P.N = fin ? 3 * F : 2 * F; P.nrows = P.N; P.nkb = nkb;
        P.col_z = fin ? 0 : -1; P.col_r = fin ? F : 0; P.col_c = fin ? 2 * F : F;
        int k = 0, m = 0;
        const SplitMap* segs[3] = {nullptr, nullptr, nullptr};
        if (x) { segs[m] = x; k = add_chain_steps(op.c, m, x->C, k, P.N); ++m; }
        if (e) { segs[m] = e; k = add_chain_steps(op.c, m, e->C, k, P.N); ++m; }
        P.gate_k0 = k;
        segs[m] = &h; k = add_chain_steps(op.c, m, h.C, k, P.col_c); ++m;
        if (k < 0) { set_error("cell(f16x3): too many operand units"); return URNN_E_UNSUPPORTED; }
        for (int i = 0; i < 3; ++i) { const SplitMap* sm = segs[i] ? segs[i] : &h; URNN_TRY(make_split_tmap(&op.c.maps[i], *sm, sm->C % 32 == 0 ? 32 : 16, 2)); }
        P.h = h.hi; P.h_plane = ntot; P.h_lo = h.lo_off();
        // image rows: block 0 = the W1 rows of this sweep, block 1 = W2 (element offset between the two matrices; both are device fp32)
        const float* w1rows = fin ? cp.w1 : cp.w1 + (long long)F * Ktot;
        const int r1 = fin ? 2 * F : F;
        P.wimg = first_parity ? add_wimg(pl, w1rows, Ktot, 1, (long long)(cp.w2 - w1rows), r1, P.N, P.N, Keff, kskip) : nullptr;
        P.bias_z = fin ? cp.b1 : nullptr; P.bias_r = cp.b1 + F; P.bias_c = cp.b2;
        P.scale_z = scale1; P.shift_z = shift1; P.scale_r = scale1 + F; P.shift_r = shift1 + F; P.scale_c = scale2; P.shift_c = shift2;
        if (fin) { P.out_hi = hout.hi; P.out_lo = hout.lo_off(); P.out_plane = ntot; }
        else {
            StatBufs sb;
            if (!take_stats(sa, F / 32, pl.num_sms, &sb)) { set_error("cell(f16x3): statistics arena too small"); return URNN_E_WORKSPACE; }
            P.nstat = F / 32;
            P.sink.partial = sb.partial; P.sink.total = sb.total; P.sink.counter = sb.counter; P.sink.nsets = F / 32; P.sink.stride = pl.num_sms; P.sink.comm = comm;
            P.aff.scale = scale2; P.aff.shift = shift2; P.aff.gamma = cp.gn2_w; P.aff.beta = cp.gn2_b; P.aff.channels = F; P.aff.ch_per_set = 32; P.aff.eps = cd.eps;
        }
        URNN_TRY(plan_chain(op.c, pl.num_sms));
        ops.push_back(op);
    }
    return URNN_OK;
}

// 1x1 conv + LeakyReLU + AvgPool2 (encoder stages 2, 3): four phase accumulators, averaged in the epilogue
static int build_pool_stem(Plan& pl, std::vector<Op>& ops, const SplitMap& src, int src_level, const SplitMap& dst, const float* w, const float* b, bool first_parity) {
    const int Cin = src.C, Cout = dst.C;
    if (Cin % 16 || Cout % 16 || Cout > 128) { set_error("pooled stem(f16x3): %d -> %d channels not supported", Cin, Cout); return URNN_E_UNSUPPORTED; }
    Op op; op.kind = 0; memset(&op.g, 0, sizeof(op.g)); params_defaults(op.g.P);
    GemmParams& P = op.g.P; fill_common(P, pl, src_level + 1);
    P.nacc = 4; P.acc_mode = ACC_POOL;
    const long long phase_stride = pl.lay.phased ? pl.lay.ntot[src_level + 1] : 0;
    for (int a = 0; a < 4; ++a) add_segment_steps(op.g, 0, Cin, 0, a, a * phase_stride);
    for (int i = 0; i < 3; ++i) URNN_TRY(make_split_tmap(&op.g.maps[i], src, Cin % 32 == 0 ? 32 : 16, 2));
    P.N = Cout; P.nrows = Cout; P.nkb = (Cin + 63) / 64;
    P.wimg = first_parity ? add_wimg(pl, w, Cin, 1, 0, 1 << 30, Cout, Cout, Cin, 0) : nullptr;
    P.bias = b; P.nbias = Cout;
    P.epi = EPI_LRELU_SPLIT; P.out_hi = dst.hi; P.out_lo = dst.lo_off(); P.out_plane = dst.ntot;
    URNN_TRY(plan_gemm(op.g, pl.num_sms));
    ops.push_back(op);
    return URNN_OK;
}

// ConvTranspose2d(k2,s2) + LeakyReLU (decoder stages 3, 2): phase a = dy*2 + dx of the fine map is an accumulator;
// launches cover as many phases as fit (weights hi+lo resident, <= 512 TMEM columns)
static int build_deconv(Plan& pl, std::vector<Op>& ops, const SplitMap& src, int src_level, const SplitMap& dst, const float* w, const float* b, bool first_parity) {
    const int Cin = src.C, Cout = dst.C;
    if (Cin % 16 || Cout % 16 || Cout > 128) { set_error("deconv(f16x3): %d -> %d channels not supported", Cin, Cout); return URNN_E_UNSUPPORTED; }
    const int nkb = (Cin + 63) / 64;
    int per = 4;
    while (per > 1 && (size_t)2 * per * Cout * nkb * 128 > (size_t)110 * 1024) per /= 2;
    for (int a0 = 0; a0 < 4; a0 += per) {
        Op op; op.kind = 0; memset(&op.g, 0, sizeof(op.g)); params_defaults(op.g.P);
        GemmParams& P = op.g.P; fill_common(P, pl, src_level);
        P.nacc = per; P.acc_mode = ACC_DECONV;
        add_segment_steps(op.g, 0, Cin, 0, 0, 0);
        for (int i = 0; i < 3; ++i) URNN_TRY(make_split_tmap(&op.g.maps[i], src, Cin % 32 == 0 ? 32 : 16, 2));
        P.N = Cout; P.nrows = per * Cout; P.nkb = nkb;
        // weight (Cin, Cout, 2, 2): row n = a*Cout + co of this launch <-> element [ci][co][a0 + a]
        P.wimg = first_parity ? add_wimg(pl, w + a0, 4, (long long)Cout * 4, 1, Cout, per * Cout, per * Cout, Cin, 0) : nullptr;
        P.bias = b; P.nbias = per * Cout; P.bias_mod = Cout;
        P.epi = EPI_LRELU_SPLIT;
        const long long blk = pl.lay.phased ? src.ntot : 0;       // fine-map phase a lives `blk` pixels after phase a-1
        P.out_hi = dst.hi + (long long)a0 * blk; P.out_lo = dst.lo_off(); P.out_plane = dst.ntot; P.out_acc_stride = blk;
        URNN_TRY(plan_gemm(op.g, pl.num_sms));
        ops.push_back(op);
    }
    return URNN_OK;
}

// decoder stage 1 stem: 1x1 conv + LeakyReLU -> decoder features as NCHW fp32 (the head's input, flood_head.py:131)
static int build_final_stem(Plan& pl, std::vector<Op>& ops, const SplitMap& src, const float* w, const float* b, int Cout, bool first_parity) {
    const int Cin = src.C;
    if (Cin % 16 || Cout % 16 || Cout > 256) { set_error("final stem(f16x3): %d -> %d channels not supported", Cin, Cout); return URNN_E_UNSUPPORTED; }
    Op op; op.kind = 0; memset(&op.g, 0, sizeof(op.g)); params_defaults(op.g.P);
    GemmParams& P = op.g.P; fill_common(P, pl, 0);
    add_segment_steps(op.g, 0, Cin, 0, 0, 0);
    for (int i = 0; i < 3; ++i) URNN_TRY(make_split_tmap(&op.g.maps[i], src, Cin % 32 == 0 ? 32 : 16, 2));
    P.N = Cout; P.nrows = Cout; P.nkb = (Cin + 63) / 64;
    P.wimg = first_parity ? add_wimg(pl, w, Cin, 1, 0, 1 << 30, Cout, Cout, Cin, 0) : nullptr;
    P.bias = b; P.nbias = Cout;
    P.epi = EPI_LRELU_F32; P.out_f32 = pl.feat; P.out_plane = (long long)pl.d.H * pl.d.W; P.store_c0 = 0; P.store_c1 = Cout;
    P.nchw = pl.lay.phased; P.nchw_w = pl.d.W; P.nchw_w4 = pl.d.W / 4; P.nchw_n4p = pl.lay.n4p;
    URNN_TRY(plan_gemm(op.g, pl.num_sms));
    ops.push_back(op);
    return URNN_OK;
}

size_t step_workspace_bytes(const urnn_ed_desc* d) {
    Plan pl; pl.d = *d; pl.lay = make_layout(d->H, d->W, true);
    layout_plan(pl, d, nullptr);
    return pl.total;
}

static int check_desc(const urnn_ed_desc* d) {
    URNN_CHECK_ARG(d && d->H > 0 && d->W > 0 && d->H % 4 == 0 && d->W % 4 == 0, "ed(f16x3): H, W must be positive multiples of 4");
    URNN_CHECK_ARG(d->ksize == 1, "ed(f16x3): filter_size must be 1");
    URNN_CHECK_ARG(d->dec_conv[2] == 16, "ed: decoder.conv_out_channels[-1]=%d must be 16 (head width, model.py:62-63)", d->dec_conv[2]);
    URNN_CHECK_ARG(d->dec_gru[0] == d->enc_gru[2] && d->dec_gru[1] == d->enc_gru[1] && d->dec_gru[2] == d->enc_gru[0],
                   "ed: decoder gru_channels must mirror the encoder's (skip concat, decoder.py:135)");
    URNN_CHECK_ARG((long long)d->H * d->W < (1LL << 30), "ed(f16x3): more than 2^30 cells per map are not supported");
    return URNN_OK;
}

// Builds the plan: workspace carving, tensor maps, launch planning for both state parities (host work only).
int build_plan(Plan& pl, const urnn_ed_desc* d, const urnn_ed_params* p, void* ws, size_t ws_bytes) {
    URNN_TRY(check_desc(d));
    pl.d = *d; pl.p = *p; pl.lay = make_layout(d->H, d->W, true);
    URNN_CUDA(cudaGetDevice(&pl.device));
    URNN_CUDA(cudaDeviceGetAttribute(&pl.num_sms, cudaDevAttrMultiProcessorCount, pl.device));
    layout_plan(pl, d, ws);
    if (pl.total > ws_bytes) { set_error("ed(f16x3): workspace %zu < %zu bytes", ws_bytes, pl.total); return URNN_E_WORKSPACE; }
    pl.wb.n = 0; pl.wb.base = pl.wimg; pl.wimg_used = 0; pl.weights_ready = false;
    StatAlloc sa{pl.stat_base, 0, pl.stat_bytes};
    urnn_cell_desc enc[3], dec[3];
    for (int k = 0; k < 3; ++k) {
        enc[k] = urnn_cell_desc{d->H >> k, d->W >> k, d->enc_conv[k], d->enc_gru[k], 1, URNN_CELL_ENCODER, d->math, d->gn_eps};
        const int sc = 2 - k;
        dec[k] = urnn_cell_desc{d->H >> sc, d->W >> sc, (k == 0) ? d->dec_conv[0] : d->dec_conv[k - 1], d->dec_gru[k], 1, URNN_CELL_DECODER, d->math, d->gn_eps};
    }
    for (int q = 0; q < 2; ++q) {
        std::vector<Op>& ops = pl.ops[q];
        ops.clear();
        const bool fp = q == 0;
        StatAlloc sq = sa;                         // both parities share the statistics buffers (they never overlap in time)
        SplitMap* in = pl.st[q]; SplitMap* out = pl.st[q ^ 1];
        size_t tagged = 0;
        auto tag = [&](const char* what) {          // names the launches appended since the last call (profiles, urnn_ed_profile_dev)
            int gi = 0;
            for (; tagged < ops.size(); ++tagged) {
                Op& o = ops[tagged];
                if (o.kind == 1) snprintf(o.name, sizeof(o.name), "%s.blend", what);
                else if (o.kind == 2) snprintf(o.name, sizeof(o.name), o.c.P.final ? "%s.C'" : "%s.B'", what);
                else if (o.g.P.gate_ch) snprintf(o.name, sizeof(o.name), "%s.B", what);
                else if (o.g.P.epi == EPI_STATS_F32) snprintf(o.name, sizeof(o.name), gi++ ? "%s.A%d" : "%s.A", what, gi);
                else snprintf(o.name, sizeof(o.name), gi++ ? "%s.%d" : "%s", what, gi);
            }
        };
        // encoder (encoder.py:187-215)
        URNN_TRY(build_cell(pl, ops, sq, enc[0], p->enc_cell[0], 0, &pl.s[0], nullptr, in[0], out[0], fp, pl.G, pl.C, 0)); tag("enc1");
        URNN_TRY(build_pool_stem(pl, ops, out[0], 0, pl.s[1], p->enc_stem_w[1], p->enc_stem_b[1], fp)); tag("stem2");
        URNN_TRY(build_cell(pl, ops, sq, enc[1], p->enc_cell[1], 1, &pl.s[1], nullptr, in[1], out[1], fp, pl.G, pl.C, 0)); tag("enc2");
        URNN_TRY(build_pool_stem(pl, ops, out[1], 1, pl.s[2], p->enc_stem_w[2], p->enc_stem_b[2], fp)); tag("stem3");
        URNN_TRY(build_cell(pl, ops, sq, enc[2], p->enc_cell[2], 2, &pl.s[2], nullptr, in[2], out[2], fp, pl.G, pl.C, 0)); tag("enc3");
        pl.n_enc[q] = ops.size();
        // decoder (decoder.py:173-217): deepest first; stage 3 has no x (ConvRNN.py:143-146)
        URNN_TRY(build_cell(pl, ops, sq, dec[0], p->dec_cell[0], 2, nullptr, &out[2], in[3], out[3], fp, pl.G2, pl.C2, 1)); tag("dec3");
        URNN_TRY(build_deconv(pl, ops, out[3], 2, pl.up3, p->dec_stem_w[0], p->dec_stem_b[0], fp)); tag("deconv3");
        URNN_TRY(build_cell(pl, ops, sq, dec[1], p->dec_cell[1], 1, &pl.up3, &out[1], in[4], out[4], fp, pl.G2, pl.C2, 1)); tag("dec2");
        URNN_TRY(build_deconv(pl, ops, out[4], 1, pl.up2, p->dec_stem_w[1], p->dec_stem_b[1], fp)); tag("deconv2");
        URNN_TRY(build_cell(pl, ops, sq, dec[2], p->dec_cell[2], 0, &pl.up2, &out[0], in[5], out[5], fp, pl.G2, pl.C2, 1)); tag("dec1");
        URNN_TRY(build_final_stem(pl, ops, out[5], p->dec_stem_w[2], p->dec_stem_b[2], d->dec_conv[2], fp)); tag("stem_out");
        {   // caps of the persistent grids of the encoder / decoder launches
            // Sharded + pipelined: the last CTA of a statistics launch spins for its peers and keeps its SM; the other
            // stream's persistent launch must not need that SM (a 148-CTA grid would run a second wave for one CTA:
            // 2 GPUs, 500^2 bands: 1.035 ms/step with full grids, 0.955 with 147 CTAs, 0.938 with 146).
            CommDev cm; current_comm(&cm);
            pl.sharded = cm.world > 1;
            const int spare = (pl.want_pipe && cm.world > 1) ? pl.num_sms - 2 : 0;
            const char* ce = getenv("URNN_V2_GRID_E"); const char* cd_ = getenv("URNN_V2_GRID_D");
            const int capE = ce ? atoi(ce) : spare, capD = cd_ ? atoi(cd_) : spare;
            for (size_t i = 0; i < ops.size(); ++i) {
                const int cap = i < pl.n_enc[q] ? capE : capD;
                if (cap <= 0) continue;
                if (ops[i].kind == 0 && ops[i].g.grid > cap) ops[i].g.grid = cap;
                if (ops[i].kind == 2 && ops[i].c.grid > cap) ops[i].c.grid = cap;
            }
        }
        if (q == 1) {                              // weight images are shared: copy the addresses recorded for parity 0
            size_t gi = 0;
            for (size_t i = 0; i < ops.size(); ++i) if (ops[i].kind != 1) { while (pl.ops[0][gi].kind == 1) ++gi; op_wimg(ops[i]) = op_wimg(pl.ops[0][gi]); ++gi; }
        }
    }
    for (Op& op : pl.ops[0]) if (op.kind != 1 && op_wimg(op) == nullptr) { set_error("ed(f16x3): weight image arena too small"); return URNN_E_WORKSPACE; }
    return URNN_OK;
}

// device work that has to precede the first step: counters, weight images
int prepare(Plan& pl, cudaStream_t st) {
    URNN_CUDA(cudaMemsetAsync(pl.stat_base, 0, pl.stat_bytes, st));
    wimg_kernel<<<dim3(16, pl.wb.n), 256, 0, st>>>(pl.wb);
    URNN_LAUNCH_CHECK();
    pl.weights_ready = true;
    return URNN_OK;
}

static int launch_pack(const float* src, const SplitMap& m, int C, int h, int w, const Layout& l, int level, cudaStream_t st) {
    const long long total = (long long)C * h * w;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    pack_kernel<<<grid, 256, 0, st>>>(src, m.hi, m.lo_off(), C, h, w, l.phased, level, l.W / 4, l.n4p, m.ntot);
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}
static int launch_unpack(const SplitMap& m, float* dst, int C, int h, int w, const Layout& l, int level, cudaStream_t st) {
    const long long total = (long long)C * h * w;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    unpack_kernel<<<grid, 256, 0, st>>>(m.hi, m.lo_off(), dst, C, h, w, l.phased, level, l.W / 4, l.n4p, m.ntot);
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}

int load_states(Plan& pl, int parity, const float* const* s_nchw, cudaStream_t st) {
    for (int k = 0; k < 6; ++k) {
        const int lv = pl.lvl_of_state[k];
        URNN_TRY(launch_pack(s_nchw[k], pl.st[parity][k], pl.ch_of_state[k], pl.d.H >> lv, pl.d.W >> lv, pl.lay, lv, st));
    }
    return URNN_OK;
}
int store_states(Plan& pl, int parity, float* const* s_nchw, cudaStream_t st) {
    for (int k = 0; k < 6; ++k) {
        const int lv = pl.lvl_of_state[k];
        URNN_TRY(launch_unpack(pl.st[parity][k], s_nchw[k], pl.ch_of_state[k], pl.d.H >> lv, pl.d.W >> lv, pl.lay, lv, st));
    }
    return URNN_OK;
}

static int launch_ew(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args, bool pdl = true) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    URNN_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}

// One time step: states of `parity` -> states of `parity ^ 1`, depth/probability maps -> out (2, H, W).
// ev (optional): cudaEvent_t pairs recorded around every launch (stage-1 stem, the program's ops, the head).
int step(Plan& pl, int parity, const float* x, int cin, const float* w, long long w_ld, const float* b, float* out, cudaStream_t st,
         cudaEvent_t* ev = nullptr) {
    const urnn_ed_desc& d = pl.d;
    int evi = 0;
#define V2_EV() do { if (ev) URNN_CUDA(cudaEventRecord(ev[evi++], st)); } while (0)
    V2_EV();
    {
        Stem1Args a{x, cin, w, w_ld, b, d.enc_conv[0], d.lrelu_slope, d.H, d.W, d.W / 4, pl.lay.phased, pl.lay.n4p, pl.s[0].ntot, pl.s[0].hi, pl.s[0].lo_off()};
        const long long nthr = (long long)d.H * (d.W / 4);
        const size_t smem = (size_t)(((d.enc_conv[0] + 15) / 16) * 16) * (cin + 1) * sizeof(float);
        if (smem > 48 * 1024) { set_error("stage-1 stem: %d x %d weights exceed 48 KB of shared memory", d.enc_conv[0], cin); return URNN_E_UNSUPPORTED; }
        void* args[1] = {(void*)&a};
        URNN_TRY(launch_ew((const void*)stem1_kernel, dim3((unsigned)((nthr + 63) / 64)), dim3(64), smem, st, args));
    }
    V2_EV();
    for (const Op& op : pl.ops[parity]) {
        V2_EV();
        if (op.kind == 0) URNN_TRY(launch_gemm(op.g, st, ev == nullptr));
        else if (op.kind == 2) URNN_TRY(launch_chain(op.c, st, ev == nullptr));
        else {
            const long long total = (op.b.ntot >> 3) * op.b.F;
            void* args[1] = {(void*)&op.b};
            URNN_TRY(launch_ew((const void*)blend_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, args));
        }
        V2_EV();
    }
    V2_EV();
    URNN_TRY(head_fwd_fp32(d.H, d.W, d.cls_thred, d.ln_eps, d.lrelu_slope, &pl.p.head, pl.feat, out, pl.head_ws, pl.head_ws_bytes, st, 1));
    V2_EV();
#undef V2_EV
    return URNN_OK;
}

// ---- two-stream software pipeline of a sequence call
// The encoder of step t+1 needs the encoder states of step t and the next input, not the decoder of step t; the decoder +
// head of step t need the encoder outputs of step t and the decoder states of step t-1.  Every GEMM launch is one
// persistent CTA per SM with a serial tail (CTAs finish up to 25 % apart, then the last CTA folds the statistics) and a
// prologue (weight image, TMEM): with the two halves on two streams the hardware fills one kernel's tail with the other
// stream's CTAs.  Buffers: the encoder states ping-pong, so encoder(t+2) overwrites what decoder(t) reads as skip input
// -> one event wait; the pre-norm scratch maps exist twice (G, C / G2, C2).  Sharded runs: the in-kernel statistic exchange
// needs every rank to issue its exchanges in the same order, so the two halves use two exchange lanes (current_comm).
static int launch_ops(Plan& pl, int parity, size_t i0, size_t i1, cudaStream_t st, bool pdl) {
    for (size_t i = i0; i < i1; ++i) {
        const Op& op = pl.ops[parity][i];
        if (op.kind == 0) URNN_TRY(launch_gemm(op.g, st, pdl));
        else if (op.kind == 2) URNN_TRY(launch_chain(op.c, st, pdl));
        else {
            const long long total = (op.b.ntot >> 3) * op.b.F;
            void* args[1] = {(void*)&op.b};
            URNN_TRY(launch_ew((const void*)blend_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, args, pdl));
        }
    }
    return URNN_OK;
}

// Two streams pay off while the serial head and tail of a launch matter (tens of microseconds against kernels of
// 20-80 us); on grids of several million cells per GPU the kernels run for milliseconds and the overlap buys nothing.
// SINGLE GPU ONLY.  In a sharded run the last CTA of every statistics launch spins for its peers; with two streams and
// programmatic dependent launch, the next launch of the same stream is already resident (blocked in griddepcontrol.wait)
// and holds the SMs the OTHER stream's launch needs on the peer the spinning CTA is waiting for -> circular wait.
// Observed: 2 GPUs ran (0.938 ms/step with two SMs left free, against 1.075 on one stream), 4 GPUs deadlocked.  The exchange
// lanes (current_comm) keep the protocol itself legal; what is missing is a launch order that cannot starve a peer.
// URNN_V2_PIPE=0 switches the pipeline off; URNN_V2_PIPE=2 forces it in sharded runs, then without programmatic dependent
// launch and with two SMs left free (the order the model in tools/sim_sharded_pipeline.py finds safe; NOT yet run on
// hardware -- the round's GPU time ended with the hang).
bool pipe_wanted(const urnn_ed_desc* d) {
    CommDev c; current_comm(&c);
    const char* e = getenv("URNN_V2_PIPE");
    const int mode = e ? atoi(e) : 1;
    if (mode == 0) return false;
    if (c.world > 1) return mode == 2;
    if (e) return true;
    return (long long)d->H * d->W <= (4LL << 20);
}

int pipe_enable(Plan& pl, bool want) {
    URNN_CUDA(cudaEventCreateWithFlags(&pl.ev_call, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        URNN_CUDA(cudaEventCreateWithFlags(&pl.evE[i], cudaEventDisableTiming));
        URNN_CUDA(cudaEventCreateWithFlags(&pl.evD[i], cudaEventDisableTiming));
    }
    if (!want) return URNN_OK;
    URNN_CUDA(cudaStreamCreateWithFlags(&pl.sE, cudaStreamNonBlocking));
    URNN_CUDA(cudaStreamCreateWithFlags(&pl.sD, cudaStreamNonBlocking));
    pl.pipe = true;
    return URNN_OK;
}

// One step of a sequence.  Work already queued on `st` is respected (the encoder waits for it); completion is NOT joined
// back into `st`: *in_done fires when the step's input has been consumed, *out_done when `out` (and the optional copies
// depth_dst / prob_dst of its two planes) is complete.  Without the pipeline both are recorded on `st`.
int seq_step(Plan& pl, int t, const float* x, int cin, const float* w, long long w_ld, const float* b, float* out,
             float* depth_dst, float* prob_dst, cudaStream_t st, cudaEvent_t* in_done, cudaEvent_t* out_done) {
    const urnn_ed_desc& d = pl.d;
    const int parity = t & 1;
    const size_t N = (size_t)d.H * d.W;
    cudaStream_t sE = pl.pipe ? pl.sE : st, sD = pl.pipe ? pl.sD : st;
    // forced sharded pipeline (URNN_V2_PIPE=2): no programmatic dependent launch -- a launch that is resident before its
    // predecessor's exchange completed is what lets two ranks starve each other (tools/sim_sharded_pipeline.py)
    const bool pdl = !(pl.pipe && pl.sharded);
    if (pl.pipe) {
        URNN_CUDA(cudaEventRecord(pl.ev_call, st));
        URNN_CUDA(cudaStreamWaitEvent(sE, pl.ev_call, 0));
        if (t >= 2) URNN_CUDA(cudaStreamWaitEvent(sE, pl.evD[parity], 0));      // decoder(t-2) has read the states encoder(t) overwrites
    }
    {
        Stem1Args a{x, cin, w, w_ld, b, d.enc_conv[0], d.lrelu_slope, d.H, d.W, d.W / 4, pl.lay.phased, pl.lay.n4p, pl.s[0].ntot, pl.s[0].hi, pl.s[0].lo_off()};
        const long long nthr = (long long)d.H * (d.W / 4);
        const size_t smem = (size_t)(((d.enc_conv[0] + 15) / 16) * 16) * (cin + 1) * sizeof(float);
        if (smem > 48 * 1024) { set_error("stage-1 stem: %d x %d weights exceed 48 KB of shared memory", d.enc_conv[0], cin); return URNN_E_UNSUPPORTED; }
        void* args[1] = {(void*)&a};
        URNN_TRY(launch_ew((const void*)stem1_kernel, dim3((unsigned)((nthr + 63) / 64)), dim3(64), smem, sE, args, pdl));
    }
    URNN_TRY(launch_ops(pl, parity, 0, pl.n_enc[parity], sE, pdl));
    if (pl.pipe) {
        URNN_CUDA(cudaEventRecord(pl.evE[parity], sE));
        URNN_CUDA(cudaStreamWaitEvent(sD, pl.evE[parity], 0));
    }
    URNN_TRY(launch_ops(pl, parity, pl.n_enc[parity], pl.ops[parity].size(), sD, pdl));
    URNN_TRY(head_fwd_fp32(d.H, d.W, d.cls_thred, d.ln_eps, d.lrelu_slope, &pl.p.head, pl.feat, out, pl.head_ws, pl.head_ws_bytes, sD, 1));
    if (depth_dst) URNN_CUDA(cudaMemcpyAsync(depth_dst, out, N * sizeof(float), cudaMemcpyDeviceToDevice, sD));
    if (prob_dst) URNN_CUDA(cudaMemcpyAsync(prob_dst, out + N, N * sizeof(float), cudaMemcpyDeviceToDevice, sD));
    URNN_CUDA(cudaEventRecord(pl.evD[parity], sD));
    if (!pl.pipe) URNN_CUDA(cudaEventRecord(pl.evE[parity], st));
    if (in_done) *in_done = pl.evE[parity];
    if (out_done) *out_done = pl.evD[parity];
    return URNN_OK;
}

// joins the pipeline back into `st` (everything issued so far is complete for work queued on `st` afterwards)
int seq_join(Plan& pl, cudaStream_t st) {
    if (!pl.pipe) return URNN_OK;
    cudaEvent_t e = pl.ev_call;
    URNN_CUDA(cudaEventRecord(e, pl.sE)); URNN_CUDA(cudaStreamWaitEvent(st, e, 0));
    URNN_CUDA(cudaEventRecord(e, pl.sD)); URNN_CUDA(cudaStreamWaitEvent(st, e, 0));
    return URNN_OK;
}

// T timed steps: mean milliseconds of every launch (CUDA events on the launching stream), names as "stem1", "enc1.A", ...
int profile(Plan& pl, int T, const float* inputs, size_t in_elems, int cin, const float* w, long long w_ld, const float* b, float* out,
            cudaStream_t st, float* op_ms, char* names, int max_ops, int* nops) {
    const int n = (int)pl.ops[0].size() + 2;
    if (n > max_ops) { set_error("profile: %d launches > max_ops=%d", n, max_ops); return URNN_E_INVALID; }
    std::vector<cudaEvent_t> ev(2 * n);
    for (auto& e : ev) URNN_CUDA(cudaEventCreate(&e));
    std::vector<double> acc(n, 0.0);
    int rc = URNN_OK;
    for (int t = 0; t < T && rc == URNN_OK; ++t) {
        rc = step(pl, t & 1, inputs + (size_t)t * in_elems, cin, w, w_ld, b, out, st, ev.data());
        if (rc != URNN_OK) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("profile: stream synchronisation failed"); rc = URNN_E_CUDA; break; }
        for (int i = 0; i < n; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]); acc[i] += ms; }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (rc != URNN_OK) return rc;
    for (int i = 0; i < n; ++i) {
        op_ms[i] = (float)(acc[i] / T);
        const char* nm = i == 0 ? "stem1" : (i == n - 1 ? "head" : pl.ops[0][i - 1].name);
        snprintf(names + (size_t)i * 24, 24, "%s", nm);
    }
    *nops = n;
    return URNN_OK;
}

// ------------------------------------------------------------------------------------------------ plan cache (per-step API)
// urnn_ed_step_fwd is called once per time step by the drop-in ED.forward; building a plan (20 tensor maps, launch
// planning) costs ~0.1 ms of host time, so the last few plans are kept, keyed by everything they depend on.
struct CacheKey { int device; urnn_ed_desc d; urnn_ed_params p; void* ws; size_t ws_bytes; int comm_world, comm_rank; void* comm_seq; };
static bool same_key(const CacheKey& a, const CacheKey& b) { return memcmp(&a, &b, sizeof(CacheKey)) == 0; }
struct CacheEnt { CacheKey key; std::unique_ptr<Plan> plan; unsigned long long stamp; };
static std::mutex g_cache_mu;
static std::vector<CacheEnt> g_cache;
static unsigned long long g_stamp = 0;

static int cached_plan(const urnn_ed_desc* d, const urnn_ed_params* p, void* ws, size_t ws_bytes, Plan** out) {
    CacheKey key; memset(&key, 0, sizeof(key));
    URNN_CUDA(cudaGetDevice(&key.device));
    key.d = *d; key.p = *p; key.ws = ws; key.ws_bytes = ws_bytes;
    { CommDev c; current_comm(&c); key.comm_world = c.world; key.comm_rank = c.rank; key.comm_seq = c.seq; }   // plans embed the communicator
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (CacheEnt& e : g_cache) if (same_key(e.key, key)) { e.stamp = ++g_stamp; *out = e.plan.get(); return URNN_OK; }
    std::unique_ptr<Plan> pl(new Plan());
    URNN_TRY(build_plan(*pl, d, p, ws, ws_bytes));
    if (g_cache.size() >= 8) {
        size_t old = 0;
        for (size_t i = 1; i < g_cache.size(); ++i) if (g_cache[i].stamp < g_cache[old].stamp) old = i;
        g_cache.erase(g_cache.begin() + old);
    }
    g_cache.push_back(CacheEnt{key, std::move(pl), ++g_stamp});
    *out = g_cache.back().plan.get();
    return URNN_OK;
}

// the reference's one-step contract with fp32 NCHW states in and out (model.py:65-121): convert, step, convert back
int step_fwd_nchw(const urnn_ed_desc* d, const urnn_ed_params* p, const float* x, int cin, const float* w, long long w_ld, const float* b,
                  const float* const* sin, float* const* sout, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    Plan* pl = nullptr;
    URNN_TRY(cached_plan(d, p, ws, ws_bytes, &pl));
    URNN_TRY(prepare(*pl, st));                    // weights may have changed between calls (training): rebuild the images
    URNN_TRY(load_states(*pl, 0, sin, st));
    URNN_TRY(step(*pl, 0, x, cin, w, w_ld, b, out, st));
    return store_states(*pl, 1, sout, st);
}

}  // namespace v2

// host-side view of the internal layout (tests, tools): position of pixel (y, x) of the level-`level` map; -1 if out of range
long long v2_layout_index(int H, int W, int level, int y, int x, long long* ntot) {
    if (H <= 0 || W <= 0 || H % 4 || W % 4 || level < 0 || level > 2) return -1;
    const v2::Layout l = v2::make_layout(H, W, true);
    if (ntot) *ntot = l.ntot[level];
    if (y < 0 || x < 0 || y >= (H >> level) || x >= (W >> level)) return -1;
    return v2::ix_internal(1, level, y, x, W >> level, W / 4, l.n4p);
}

// ---- entry points used by capi.cu
size_t v2_step_workspace_bytes(const urnn_ed_desc* d) { return v2::step_workspace_bytes(d); }
int v2_step_fwd_nchw(const urnn_ed_desc* d, const urnn_ed_params* p, const float* x, int cin, const float* w, long long w_ld, const float* b,
                     const float* const* sin, float* const* sout, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    return v2::step_fwd_nchw(d, p, x, cin, w, w_ld, b, sin, sout, out, ws, ws_bytes, st);
}
V2Seq* v2_seq_begin(const urnn_ed_desc* d, const urnn_ed_params* p, const float* const* states, void* ws, size_t ws_bytes, cudaStream_t st, int* rc,
                    bool pipelined) {
    std::unique_ptr<v2::Plan> pl(new v2::Plan());
    pl->want_pipe = pipelined && v2::pipe_wanted(d);
    *rc = v2::build_plan(*pl, d, p, ws, ws_bytes);
    if (*rc == URNN_OK) *rc = v2::prepare(*pl, st);
    if (*rc == URNN_OK) *rc = v2::load_states(*pl, 0, states, st);
    if (*rc == URNN_OK) *rc = v2::pipe_enable(*pl, pl->want_pipe);
    if (*rc != URNN_OK) return nullptr;
    return reinterpret_cast<V2Seq*>(pl.release());
}
int v2_seq_step(V2Seq* s, int t, const float* x, int cin, const float* w, long long w_ld, const float* b, float* out,
                float* depth_dst, float* prob_dst, cudaStream_t st, cudaEvent_t* in_done, cudaEvent_t* out_done) {
    return v2::seq_step(*reinterpret_cast<v2::Plan*>(s), t, x, cin, w, w_ld, b, out, depth_dst, prob_dst, st, in_done, out_done);
}
int v2_seq_profile(V2Seq* s, int T, const float* inputs, size_t in_elems, int cin, const float* w, long long w_ld, const float* b, float* out,
                   cudaStream_t st, float* op_ms, char* names, int max_ops, int* nops) {
    return v2::profile(*reinterpret_cast<v2::Plan*>(s), T, inputs, in_elems, cin, w, w_ld, b, out, st, op_ms, names, max_ops, nops);
}
int v2_seq_end(V2Seq* s, int T, float* const* states, cudaStream_t st) {
    std::unique_ptr<v2::Plan> pl(reinterpret_cast<v2::Plan*>(s));
    if (!states) { if (pl->pipe) { cudaStreamSynchronize(pl->sE); cudaStreamSynchronize(pl->sD); } return URNN_OK; }   // error path: drain before the streams die
    URNN_TRY(v2::seq_join(*pl, st));
    return v2::store_states(*pl, T & 1, states, st);
}

}  // namespace urnn
