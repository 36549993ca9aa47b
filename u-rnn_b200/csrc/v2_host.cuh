// v2_host.cuh -- host side of gemm_v2.cuh: split-map descriptors, TMA tensor maps, shared-memory / TMEM planning, launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>
#include <stdlib.h>
#include "gemm_v2_chain.cuh"

namespace urnn {
namespace v2 {

// A "split map": C channels x ntot pixels stored as 16-bit hi planes followed by 16-bit lo planes, [2][C][ntot] (fp16 pairs, see ptx_sm100.cuh).
// v = hi + lo.  ntot is a multiple of 128; pixels are grouped in blocks of blk_stride of which the first blk_valid are
// real (phase-separated layout of the encoder-decoder, see layout_v2 in urnn_v2.cu; a plain map is one block).
struct SplitMap {
    sp16* hi; int C; long long ntot;
    sp16* lo() const { return hi + (long long)C * ntot; }
    long long lo_off() const { return (long long)C * ntot; }
    static size_t bytes(int C, long long ntot) { return (size_t)4 * C * ntot; }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;           // the driver entry point is process-wide, not per device
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

static inline CUtensorMapL2promotion l2promo() {
    const char* e = getenv("URNN_V2_L2PROMO");
    const int v = e ? atoi(e) : 256;
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
}

// tensor map over a split map: dims (pixels, channels, hi|lo), box (64 pixels, unit_ch channels, nhl), SWIZZLE_128B
static inline int make_split_tmap(CUtensorMap* tm, const SplitMap& m, int unit_ch, int nhl) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return URNN_E_CUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)m.ntot, (cuuint64_t)m.C, 2};
    cuuint64_t strides[2] = {(cuuint64_t)m.ntot * 2, (cuuint64_t)m.C * m.ntot * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)unit_ch, (cuuint32_t)nhl};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, (SPLIT_FMT == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16), 3, m.hi, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, l2promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for map C=%d ntot=%lld unit=%d", (int)r, m.C, m.ntot, unit_ch); return URNN_E_CUDA; }
    return URNN_OK;
}

// One prepared launch: everything the kernel needs, built once per sequence and replayed every step.
struct GemmLaunch {
    CUtensorMap maps[3];
    GemmParams P;
    bool gated;
    int grid; size_t smem;
};

static inline void params_defaults(GemmParams& P) {
    memset(&P, 0, sizeof(P));
    P.nacc = 1; P.acc_mode = ACC_SINGLE; P.nmma = 3; P.gdepth = 1; P.bias_mod = 1 << 30;
    P.sink.comm.world = 1;
    P.aff.ch_per_set = 32;
}

// segments: up to 3 split maps concatenated along K.  Appends the TMA steps for accumulator `acc`, source pixel offset
// `pix_off`, starting at weight-image channel k0; returns the next k.
static inline int add_segment_steps(GemmLaunch& L, int map, int nch, int k0, int acc, long long pix_off) {
    const int unit = (nch % 32 == 0) ? 32 : 16;
    for (int c = 0; c < nch; c += unit) {
        Step& s = L.P.steps[L.P.nsteps++];
        s.map = map; s.c0 = c; s.unit_ch = unit; s.kglob = k0 + c; s.acc = acc; s.pad = 0; s.pix_off = pix_off;
    }
    return k0 + nch;
}

// Chooses ring depth / TMEM layout; fails if the weights do not fit.  Call after P is filled (N, nkb, nrows, nacc, gate_ch).
static inline int plan_gemm(GemmLaunch& L, int num_sms) {
    GemmParams& P = L.P;
    const int ncols_total = P.acc_mode == ACC_DECONV ? P.nacc * P.N : P.N;
    // accumulators: stride between accumulators of one stage, stages
    if (P.nacc == 1) {
        // up to four accumulator stages: the chain MMA issue -> commit -> epilogue wake-up -> epilogue is latency-bound per
        // stage (~4 us for a 128-column tile), so the stages in flight set the tile rate
        int stride = 32; while (stride < P.N) stride <<= 1;
        int stages = 512 / stride; if (stages > MAX_ACC_STAGES) stages = MAX_ACC_STAGES;
        if (const char* e = getenv("URNN_V2_STAGES")) { const int cap = atoi(e); if (cap >= 1 && cap < stages) stages = cap; }
        int cols = 32; while (cols < stages * stride) cols <<= 1;
        P.tmem_cols = cols; P.acc_stride = stride; P.acc_stages = stages;
    } else {
        int stride = 32; while (stride < P.N) stride <<= 1;
        const int need = P.nacc * stride;
        if (need > 512) { set_error("gemm_v2: %d accumulators of %d columns do not fit in tensor memory", P.nacc, P.N); return URNN_E_UNSUPPORTED; }
        P.acc_stride = stride; P.acc_stages = (2 * need <= 512) ? 2 : 1;
        int cols = 32; while (cols < P.acc_stages * need) cols <<= 1;
        P.tmem_cols = cols;
    }
    auto pick = [&](int gd) {
        for (int ns = 8; ns >= 2; --ns)
            if (smem_plan(P.nkb, P.nrows, ns, P.gate_ch, gd, ncols_total).total <= SMEM_MAX) return ns;
        return 0;
    };
    int best = 0;
    if (P.gate_ch == 0) { P.gdepth = 0; best = pick(0); }
    else {
        // double-buffered gate operands only when a deep ring still fits beside them
        const int ns2 = pick(2);
        if (ns2 >= 4) { P.gdepth = 2; best = ns2; } else { P.gdepth = 1; best = pick(1); }
    }
    if (best < 2) { set_error("gemm_v2: weights %dx%d (hi+lo) leave no room for the operand ring", P.nrows, P.nkb * 64); return URNN_E_UNSUPPORTED; }
    if (const char* e = getenv("URNN_V2_SLOTS")) { const int cap = atoi(e); if (cap >= 2 && cap < best) best = cap; }   // bring-up: ring depth cap
    P.nslots = best;
    P.l2_ahead = 0;
    { const char* e = getenv("URNN_V2_L2HINT"); const int v = e ? atoi(e) : 0; P.l2_hint = v == 1 ? L2_EVICT_FIRST : (v == 2 ? L2_EVICT_LAST : L2_EVICT_NORMAL); }
    const SmemPlan sp = smem_plan(P.nkb, P.nrows, P.nslots, P.gate_ch, P.gdepth, ncols_total);
    L.smem = sp.total;
    for (int s = 0; s < P.nsteps; ++s) {
        const Step& st = P.steps[s];
        MmaStep& m = P.msteps[s];
        const int kg = st.kglob, kg1 = kg + 16;        // the second K = 16 group may start the next 64-channel weight block
        m.a_desc = smem_desc_mn_sw128(sp.ring_off, (uint32_t)st.unit_ch * 256u);
        m.a_lo_delta = ((uint32_t)st.unit_ch * 128u) >> 4;
        m.b_desc[0] = smem_desc_sw128(sp.w_off + (uint32_t)(kg >> 6) * P.nrows * 128u + (uint32_t)((kg & 63) >> 4) * 32u);
        m.b_desc[1] = smem_desc_sw128(sp.w_off + (uint32_t)(kg1 >> 6) * P.nrows * 128u + (uint32_t)((kg1 & 63) >> 4) * 32u);
        m.first = ((P.acc_mode == ACC_POOL ? kg : s) == 0) ? 0u : 1u;
        m.d_off = P.acc_mode != ACC_DECONV ? (unsigned)(st.acc * P.acc_stride) : 0u;
        m.nj = (unsigned)(st.unit_ch >> 4);
    }
    const long long ntiles = P.ntot / TILE_M;
    L.grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    L.gated = P.gate_ch > 0;
    return URNN_OK;
}

static inline int launch_gemm(const GemmLaunch& L, cudaStream_t st, bool pdl = true) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(L.gated ? NTHREADS_GATED : NTHREADS_PLAIN);
    cfg.dynamicSmemBytes = L.smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    // (cheap, and correct with several devices in one process: the attribute is per device)
    if (L.gated) {
        URNN_CUDA(cudaFuncSetAttribute(gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        URNN_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<true>, L.maps[0], L.maps[1], L.maps[2], L.P));
    } else {
        URNN_CUDA(cudaFuncSetAttribute(gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        URNN_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<false>, L.maps[0], L.maps[1], L.maps[2], L.P));
    }
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}

// ---- recompute ("chain") launches of a ConvGRU cell (gemm_v2_chain.cuh)
struct ChainLaunch { CUtensorMap maps[3]; ChainParams P; int grid; size_t smem; };

static inline int add_chain_steps(ChainLaunch& L, int map, int nch, int k0, int ncols) {
    const int unit = (nch % 32 == 0) ? 32 : 16;
    for (int c = 0; c < nch; c += unit) {
        if (L.P.nsteps >= MAX_STEPS) return -1;
        L.P.step_ncols[L.P.nsteps] = ncols;
        Step& s = L.P.steps[L.P.nsteps++];
        s.map = map; s.c0 = c; s.unit_ch = unit; s.kglob = k0 + c; s.acc = 0; s.pad = 0; s.pix_off = 0;
    }
    return k0 + nch;
}

// true when the cell's [W1 ; W2] image, the gate buffers and a ring of >= min_slots fit in one CTA's shared memory
static inline int chain_slots(int nkb, int nrows, int F, int gdepth) {
    for (int ns = 8; ns >= 2; --ns) if (chain_smem(nkb, nrows, ns, F, gdepth).total <= SMEM_MAX) return ns;
    return 0;
}

static inline int plan_chain(ChainLaunch& L, int num_sms) {
    ChainParams& P = L.P;
    if (P.N % 16 || P.N > 256 || P.F % 32) { set_error("chain: %d accumulator columns / %d gated channels not supported", P.N, P.F); return URNN_E_UNSUPPORTED; }
    int cols = 32; while (cols < 2 * P.N) cols <<= 1;
    if (cols > 512) { set_error("chain: two stages of %d columns exceed tensor memory", P.N); return URNN_E_UNSUPPORTED; }
    P.tmem_cols = cols; P.acc_stride = P.N <= cols / 2 ? cols / 2 : P.N;
    int gd = 2, ns = chain_slots(P.nkb, P.nrows, P.F, 2);
    if (ns < 3) { gd = 1; ns = chain_slots(P.nkb, P.nrows, P.F, 1); }
    if (const char* e = getenv("URNN_V2_CHAIN_GD")) { const int v = atoi(e); if (v == 1 || v == 2) { gd = v; ns = chain_slots(P.nkb, P.nrows, P.F, gd); } }
    if (ns < 2) { set_error("chain: weights %dx%d (hi+lo) leave no room for the operand ring", P.nrows, P.nkb * 64); return URNN_E_UNSUPPORTED; }
    if (const char* e = getenv("URNN_V2_SLOTS")) { const int cap = atoi(e); if (cap >= 2 && cap < ns) ns = cap; }
    P.gdepth = gd; P.nslots = ns;
    L.smem = chain_smem(P.nkb, P.nrows, P.nslots, P.F, P.gdepth).total;
    const long long ntiles = P.ntot / TILE_M;
    L.grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    return URNN_OK;
}

static inline int launch_chain(const ChainLaunch& L, cudaStream_t st, bool pdl = true) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(NTHREADS_CHAIN);
    cfg.dynamicSmemBytes = L.smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    URNN_CUDA(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    URNN_CUDA(cudaLaunchKernelEx(&cfg, chain_kernel, L.maps[0], L.maps[1], L.maps[2], L.P));
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}

}  // namespace v2
}  // namespace urnn
