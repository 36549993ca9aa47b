// gemm_v2.cuh -- tcgen05 "pixel GEMM" for sm_100a, second generation:  D[128 pixels x N] = X^T[128 x K] * W^T[K x N]
//
// Precision: every operand is a bf16 hi + lo pair (v = hi + lo, 16-17 significant bits) and every K=16 step issues
// three MMAs  A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  into an fp32 accumulator in tensor memory.  (Measured on the reference's
// recurrence, tools/precision_sweep.py: single-pass bf16/tf32 operands or bf16 intermediate maps leave the config-3
// tolerance at T=180; >= 16-bit operands on BOTH sides with fp32 intermediates stay inside it.)
//
// Data path: activations live in HBM as "split maps"  [2 (hi|lo)][C][Ntot] bf16, pixel-contiguous.  A TMA tensor load
// (box 64 pixels x unit channels x {hi,lo}, SWIZZLE_128B) lands them directly in the canonical MN-major UMMA layout: no
// SIMT stage between HBM and the tensor core.  The only SIMT-produced operand is the reset-gated state r*h of the
// candidate GEMM (K = F channels), written by four gate warps.  Weights are resident in shared memory as a hi and a lo
// image (K-major SWIZZLE_128B), prepared once per sequence.
//
// One persistent CTA per SM:  warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM owner, warps 4-19 epilogue
// (tcgen05.ld: lane = pixel, column = output channel), warps 20-23 gate warps (gated launches only).
// Pipelines: operand ring (full/empty), gated-operand buffers (gfull/gempty), accumulator stages (tfull/tempty).
#pragma once
#include <cuda.h>
#include "ptx_sm100.cuh"
#include "urnn_common.cuh"

namespace urnn {
namespace v2 {

using namespace ptx;

constexpr int TILE_M = 128;
constexpr int SLOT_BYTES = 16384;         // one ring slot: 32 channels x 128 pixels x {hi,lo} 16-bit
constexpr int MAX_STEPS = 16;
constexpr int MAXG = 8;                   // 32-column groups per accumulator (N <= 256)
constexpr int NWARP_EPI = 16;                 // plain launches: 16 epilogue warps; gated launches: 8 epilogue + 12 gate warps
constexpr int NWARP_EPI_GATED = 8, NWARP_GATE = 12;
constexpr int EPI_WARP0 = 4;
constexpr int NTHREADS_PLAIN = 32 * (EPI_WARP0 + NWARP_EPI), NTHREADS_GATED = 32 * (EPI_WARP0 + NWARP_EPI_GATED + NWARP_GATE);
constexpr size_t SMEM_MAX = 229376;       // 227 KB opt-in limit minus 3 KB for static shared memory (exchange staging)

enum { EPI_STATS_F32 = 0, EPI_LRELU_SPLIT = 1, EPI_LRELU_F32 = 2 };
enum { ACC_SINGLE = 0, ACC_POOL = 1, ACC_DECONV = 2 };

// ---- normalisation statistics, mean-shifted: every epilogue warp keeps (n, S1 = sum(x-K), S2 = sum((x-K)^2), K) per group
// with its own pilot K (the first value it sees), the last CTA converts the partials to (n, mean, M2) in double and merges
// them with Chan's formula in a fixed order: no E[x^2]-E[x]^2 cancellation however large |mean|/sigma is.
struct StatSink2 {
    float4*   partial;   // [nsets][stride]
    double*   total;     // [nsets][4] = (mean, M2, n, -)
    unsigned* counter;
    int       nsets;
    int       stride;    // >= CTAs
    CommDev   comm;
};

__device__ __forceinline__ void chan_merge(double& n, double& mean, double& m2, double nb, double meanb, double m2b) {
    if (nb <= 0.0) return;
    if (n <= 0.0) { n = nb; mean = meanb; m2 = m2b; return; }
    const double nn = n + nb, d = meanb - mean;
    mean += d * (nb / nn);
    m2 += m2b + d * d * (n * nb / nn);
    n = nn;
}

// One-shot all-reduce of the (mean, M2, n) triples over NVLink peer memory (ll_allgather, urnn_common.cuh), merged in
// rank order with Chan's formula: every rank obtains bit-identical totals.
__device__ __forceinline__ void stats2_exchange(const StatSink2& s) {
    const CommDev& c = s.comm;
    if (c.world <= 1) return;
    __shared__ unsigned xin[COMM_MAX_SETS * 6], xout[COMM_MAX_WORLD * COMM_MAX_SETS * 6];
    if (threadIdx.x < s.nsets) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const unsigned long long v = (unsigned long long)__double_as_longlong(s.total[4 * threadIdx.x + k]);
            xin[6 * threadIdx.x + 2 * k] = (unsigned)v; xin[6 * threadIdx.x + 2 * k + 1] = (unsigned)(v >> 32);
        }
    }
    __syncthreads();
    ll_allgather(c, xin, s.nsets, 6, xout);
    if (threadIdx.x < s.nsets) {
        double n = 0.0, mean = 0.0, m2 = 0.0;
        for (int r = 0; r < c.world; ++r) {
            const unsigned* w = xout + (r * s.nsets + threadIdx.x) * 6;
            double t[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) t[k] = __longlong_as_double((long long)((unsigned long long)w[2 * k] | ((unsigned long long)w[2 * k + 1] << 32)));
            chan_merge(n, mean, m2, t[2], t[0], t[1]);
        }
        s.total[4 * threadIdx.x] = mean; s.total[4 * threadIdx.x + 1] = m2; s.total[4 * threadIdx.x + 2] = n;
    }
    __syncthreads();
}

// Called by ALL threads of the CTA after its partials are written.  The last CTA merges, exchanges, folds the affine.
__device__ __forceinline__ void stats2_finalize_last_cta(const StatSink2& s, int ncontrib, unsigned ncta_total, const AffineOut* aff) {
    __shared__ bool is_last;
    __shared__ double tot_s[MAXG * 4];
    // the affine parameters of "my" channel are fetched by every CTA before the fence (a wasted 8-byte load for all but the
    // last one): the finalizer is a serial tail of the launch, every dependent round trip to L2 in it costs ~1 us
    const bool want_aff = aff != nullptr && aff->scale != nullptr;
    float gamma_c = 0.f, beta_c = 0.f;
    if (want_aff && (int)threadIdx.x < aff->channels) { gamma_c = __ldg(aff->gamma + threadIdx.x); beta_c = __ldg(aff->beta + threadIdx.x); }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(s.counter, 1u);
        is_last = (t == ncta_total - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // one warp per set; the partials (one per CTA) are re-referenced to the first non-empty partial's pilot K00 in double --
    // only multiplies and adds, no divisions in the loop (a single warp runs ~8 cycles per dependent instruction) -- and
    // summed in a fixed order: lane-strided, then a fixed shuffle tree.
    const int nwarp = blockDim.x >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int set = warp; set < s.nsets; set += nwarp) {
        const float4* part = s.partial + (size_t)set * s.stride;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < ncontrib) v0 = __ldcg(part + lane);
        double k00 = 0.0;
        {
            const unsigned have = __ballot_sync(0xffffffffu, v0.x > 0.f);
            if (have) k00 = (double)__shfl_sync(0xffffffffu, v0.w, __ffs(have) - 1);
            else for (int i = 32; i < ncontrib; ++i) { const float4 v = __ldcg(part + i); if (v.x > 0.f) { k00 = (double)v.w; break; } }
        }
        double n = 0.0, a1 = 0.0, a2 = 0.0;
        for (int i = lane; i < ncontrib; i += 32) {
            const float4 v = i == lane ? v0 : __ldcg(part + i);
            if (!(v.x > 0.f)) continue;              // empty partial (all-padding CTA): its pilot may be garbage
            const double nb = (double)v.x, s1 = (double)v.y, s2 = (double)v.z, d = (double)v.w - k00;
            n += nb; a1 += s1 + nb * d; a2 += s2 + 2.0 * d * s1 + nb * d * d;
        }
        n = warp_sum(n); a1 = warp_sum(a1); a2 = warp_sum(a2);
        if (lane == 0) {
            const double m = n > 0.0 ? a1 / n : 0.0;
            double m2 = a2 - a1 * m;                 // sum (x - mean)^2
            if (m2 < 0.0) m2 = 0.0;
            s.total[4 * set] = k00 + m; s.total[4 * set + 1] = m2; s.total[4 * set + 2] = n;
            tot_s[4 * set] = k00 + m; tot_s[4 * set + 1] = m2; tot_s[4 * set + 2] = n;
        }
    }
    __syncthreads();
    if (s.comm.world > 1) {
        stats2_exchange(s);
        if ((int)threadIdx.x < 4 * s.nsets) tot_s[threadIdx.x] = s.total[threadIdx.x];
        __syncthreads();
    }
    if (want_aff) {
        for (int c = threadIdx.x; c < aff->channels; c += blockDim.x) {
            const int set = c / aff->ch_per_set;
            const double mean = tot_s[4 * set], n = tot_s[4 * set + 2];
            double var = n > 0.0 ? tot_s[4 * set + 1] / n : 0.0;
            if (var < 0.0) var = 0.0;
            const double rstd = 1.0 / sqrt(var + (double)aff->eps);
            const bool mine = c == (int)threadIdx.x;
            const double sc = (double)(mine ? gamma_c : aff->gamma[c]) * rstd;
            aff->scale[c] = (float)sc;
            aff->shift[c] = (float)((double)(mine ? beta_c : aff->beta[c]) - mean * sc);
        }
    }
    if (threadIdx.x == 0) *s.counter = 0u;
}

// ---- launch description
struct Step {                 // one TMA-loaded operand unit of a tile
    int map;                  // tensor map index (0..2)
    int c0;                   // first channel inside that map
    int unit_ch;              // 16 or 32 channels
    int kglob;                // position in the weight image's K axis
    int acc;                  // accumulator (ACC_POOL: one per 2x2 phase)
    int pad;
    long long pix_off;        // added to the tile's first pixel (phase block offsets)
};

// What the MMA issuer needs per operand unit, precomputed by plan_gemm (descriptors relative to the CTA's shared-memory
// base: the start-address field is additive, the kernel adds base >> 4 and the ring slot).
struct MmaStep {
    unsigned long long a_desc;        // hi half of the unit in ring slot 0
    unsigned long long b_desc[2];     // weight rows (hi image) of the unit's first / second K = 16 group
    unsigned a_lo_delta;              // descriptor distance hi -> lo half of the unit
    unsigned d_off;                   // accumulator column offset inside the stage
    unsigned first;                   // 0: this unit's first MMA overwrites the accumulator
    unsigned nj;                      // K = 16 groups in the unit (1 or 2)
};
constexpr int MAX_ACC_STAGES = 4;

struct GemmParams {
    Step steps[MAX_STEPS]; int nsteps;
    MmaStep msteps[MAX_STEPS];
    int nacc, acc_mode;
    // reset-gated operand segment (candidate GEMM): r*h with r = sigmoid(gate_pre*scale + shift), produced by the gate warps
    int gate_ch;                                   // 0: none, else F (multiple of 32)
    int gate_k0;                                   // position of the gated channels in the weight image's K axis
    const sp16* gate_h; long long gate_h_plane, gate_h_lo;    // split map of h: hi planes, lo planes gate_h_lo elements further
    const float* gate_pre; long long gate_pre_plane;                   // fp32 pre-GroupNorm reset-gate map [F][plane]
    const float* gate_scale; const float* gate_shift;                  // folded GroupNorm affine of those F channels
    int gdepth;                                    // gated-operand buffers (1 or 2)
    // weights: resident image, hi then lo, each nkb blocks of (nrows x 128 B) (64 channels, K-major, SWIZZLE_128B)
    const void* wimg; int nkb, nrows;
    int N;                                         // columns per accumulator (multiple of 16, <= 256)
    int nmma;                                      // 3: hi/lo split product; 1: single pass (hi parts only)
    // pixels: tiles cover [0, ntot); pixel p is real iff (p % blk_stride) < blk_valid
    long long ntot, blk_stride, blk_valid;
    // epilogue
    int epi; float slope;
    const float* bias; int nbias, bias_mod;        // bias[col % bias_mod] for col < nbias
    float* out_f32; long long out_plane;           // fp32 planes out_f32[col*out_plane + p]
    int store_c0, store_c1;                        // EPI_STATS_F32: columns [store_c0, store_c1) are stored
    sp16* out_hi; long long out_lo;       // split map destination: hi planes, lo planes out_lo elements further
    int nchw, nchw_w, nchw_w4; long long nchw_n4p;  // EPI_LRELU_F32: destination is an NCHW image (un-permute the phase-separated layout)
    long long out_acc_stride;                      // ACC_DECONV: pixel offset between the destination blocks of the accumulators
    int nstat; StatSink2 sink; AffineOut aff;      // leading 32-column groups with GroupNorm statistics
    int l2_ahead;                                  // (unused)
    unsigned long long l2_hint;                    // L2 eviction policy of the TMA operand loads
    long long* dbg;                                // bring-up: clock64 stamps of CTA 0, [event 0..7][tile 0..31]
    int nslots, tmem_cols, acc_stages, acc_stride; // ring depth; TMEM allocation; stages and column stride between accumulators
};

struct SmemPlan { uint32_t w_off, ring_off, gbuf_off, bias_off, gaff_off, bar_off, red_off, ptab_off, mtab_off, total; };

__host__ __device__ inline SmemPlan smem_plan(int nkb, int nrows, int nslots, int gate_ch, int gdepth, int ncols_total) {
    SmemPlan s;
    uint32_t o = 0;
    s.w_off = o; o += 2u * nkb * nrows * 128u;
    s.ring_off = o; o += (uint32_t)nslots * SLOT_BYTES;
    s.gbuf_off = o; o += (uint32_t)gdepth * (gate_ch / 32) * SLOT_BYTES;
    s.bias_off = o; o += ((uint32_t)ncols_total * 4u + 127u) & ~127u;
    s.gaff_off = o; o += ((uint32_t)gate_ch * 8u + 127u) & ~127u;
    s.bar_off = o; o += 512;
    s.red_off = o; o += NWARP_EPI * MAXG * 16;
    s.ptab_off = o; o += MAX_STEPS * 32;            // producer table
    s.mtab_off = o; o += MAX_STEPS * 2 * 32;        // MMA table (one entry per K=16 group)
    s.total = o + 1024;      // alignment slack
    return s;
}

// Per-step tables in shared memory: the single-thread control roles run one dependent instruction every ~7 cycles, so
// everything that can be precomputed is (measured: 1500 cycles per step with descriptors built in the loop).
struct ProdEnt { unsigned long long map; int c0; int pix_off; unsigned half_bytes; unsigned pad[3]; };          // 32 B
struct MmaEnt { unsigned long long a_hi, b_hi; unsigned a_lo_delta, acc_col, first, pad; };                     // 32 B

template <bool GATED>
__global__ void __launch_bounds__(GATED ? NTHREADS_GATED : NTHREADS_PLAIN, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
            const __grid_constant__ CUtensorMap map2, const GemmParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // provably warp-uniform (uniform registers for TMEM / MMA operands)
    const int nthreads = GATED ? NTHREADS_GATED : NTHREADS_PLAIN;
    constexpr int NWE = GATED ? NWARP_EPI_GATED : NWARP_EPI;      // epilogue warps of this instantiation
    constexpr int GATE_WARP0 = EPI_WARP0 + NWE;
    const int ncols_total = P.acc_mode == ACC_DECONV ? P.nacc * P.N : P.N;
    const SmemPlan L = smem_plan(P.nkb, P.nrows, P.nslots, P.gate_ch, P.gdepth, ncols_total);
    // barriers: full[8] | empty[8] | tfull[4] | tempty[4] | gfull[2] | gempty[2] | tmem slot
    const uint32_t full0 = base + L.bar_off, empty0 = full0 + 64, tfull0 = empty0 + 64, tempty0 = tfull0 + 32;
    const uint32_t gfull0 = tempty0 + 32, gempty0 = gfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.bar_off + 256);
    float* sbias = reinterpret_cast<float*>(sm + L.bias_off);
    float* sgaff = reinterpret_cast<float*>(sm + L.gaff_off);
    float4* red = reinterpret_cast<float4*>(sm + L.red_off);     // [NWARP_EPI][MAXG] = (n, S1, S2, K)

    if (P.dbg && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMax((unsigned long long*)&P.dbg[8 * 32 + 1], ~t_); }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // ---- one-time setup (reads only parameters and the constant weight image)
    if (tid == 0) {
        for (int s = 0; s < P.nslots; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < MAX_ACC_STAGES; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, NWE); }
        for (int a = 0; a < 2; ++a) { mbar_init(gfull0 + 8 * a, NWARP_GATE); mbar_init(gempty0 + 8 * a, 1); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map0); tma_prefetch_desc(&map1); tma_prefetch_desc(&map2); }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
    for (int i = tid; i < NWARP_EPI * MAXG; i += nthreads) red[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < ncols_total; i += nthreads) sbias[i] = (P.bias != nullptr && i < P.nbias) ? __ldg(P.bias + (i % P.bias_mod)) : 0.f;
    ProdEnt* ptab = reinterpret_cast<ProdEnt*>(sm + L.ptab_off);
    if (tid < P.nsteps) {
        const Step& st = P.steps[tid];
        const uint32_t hl = P.nmma == 3 ? 2u : 1u;
        ProdEnt pe;
        pe.map = reinterpret_cast<unsigned long long>(st.map == 0 ? (const void*)&map0 : (st.map == 1 ? (const void*)&map1 : (const void*)&map2));
        pe.c0 = st.c0; pe.pix_off = (int)st.pix_off; pe.half_bytes = (uint32_t)st.unit_ch * 128u * hl; pe.pad[0] = pe.pad[1] = pe.pad[2] = 0;
        ptab[tid] = pe;
    }
    {
        const int nvec = 2 * P.nkb * P.nrows * 8;            // 16-byte pieces of the weight image
        const char* src = reinterpret_cast<const char*>(P.wimg);
        const int rot = (int)((blockIdx.x * 37u) % (unsigned)(nvec >> 5)) << 5;   // CTAs start at different L2 lines
        for (int i = tid; i < nvec; i += nthreads) {
            const int j = i + rot >= nvec ? i + rot - nvec : i + rot;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + L.w_off + (uint32_t)j * 16u), "l"(src + (size_t)j * 16u));
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // from here on the previous kernel's outputs are consumed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if constexpr (GATED) {
        for (int i = tid; i < P.gate_ch; i += nthreads) { sgaff[2 * i] = -1.4426950408889634f * __ldg(P.gate_scale + i); sgaff[2 * i + 1] = -1.4426950408889634f * __ldg(P.gate_shift + i); }
        __syncthreads();
    }
#define V2_STAMP(ev, ti) do { if (P.dbg && blockIdx.x == 0 && lane == 0 && (ti) < 32) P.dbg[(ev) * 32 + (ti)] = clock64(); } while (0)
    const long long ntiles = P.ntot / TILE_M;
    if (P.dbg && blockIdx.x == 0 && tid == 0) P.dbg[8 * 32] = clock64();
    if (P.dbg && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); P.dbg[9 * 32 + blockIdx.x] = (long long)t_; unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); P.dbg[9 * 32 + 512 + blockIdx.x] = sm_; }
    const uint32_t wlo = (uint32_t)P.nkb * P.nrows * 128u;   // lo image offset

    if (warp == 0 || warp == 3) {
        // =========================================================================== TMA producers (two: even / odd units)
        if (lane == 0) {
            const int me = warp == 0 ? 0 : 1;
            int slot = me; uint32_t ph = 0;
            if (slot >= P.nslots) { slot -= P.nslots; ph ^= 1; }
            int s = me;                                      // step inside the tile
            long long tile = blockIdx.x;
            while (s >= P.nsteps) { s -= P.nsteps; tile += gridDim.x; }
            int ti = 0;
            while (tile < ntiles) {
                const ProdEnt pe = ptab[s];
                const int px = (int)(tile * TILE_M) + pe.pix_off;
                mbar_wait(empty0 + 8 * slot, ph ^ 1);
                const uint32_t dst = base + L.ring_off + (uint32_t)slot * SLOT_BYTES;
                const uint32_t bar = full0 + 8 * slot;
                mbar_arrive_expect_tx(bar, 2u * pe.half_bytes);
                tma_load_3d(dst, reinterpret_cast<const void*>(pe.map), bar, px, pe.c0, 0, P.l2_hint);
                tma_load_3d(dst + pe.half_bytes, reinterpret_cast<const void*>(pe.map), bar, px + 64, pe.c0, 0, P.l2_hint);
                if (s <= 1) V2_STAMP(0 + me, ti);
                slot += 2; if (slot >= P.nslots) { slot -= P.nslots; ph ^= 1; }
                s += 2; while (s >= P.nsteps) { s -= P.nsteps; tile += gridDim.x; ++ti; }
            }
        }
    } else if (warp == 1) {
        // =========================================================================== MMA issuer
        // The whole warp walks the pipeline with warp-uniform values (descriptors live in uniform registers); one elected
        // lane issues tcgen05.mma / tcgen05.commit.
        const uint32_t idesc = instr_desc_16(P.N, SPLIT_FMT);
        int slot = 0; uint32_t ph = 0; int as = 0; uint32_t aph = 0; int gd = 0; uint32_t gph = 0;
        const uint32_t wlo16 = wlo >> 4, base16 = base >> 4;
        int ti = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            mbar_wait(tempty0 + 8 * as, aph ^ 1);
            tc_fence_after();
            V2_STAMP(2, ti);
            const uint32_t d0 = tmem_base + (uint32_t)(as * P.nacc * P.acc_stride);
            for (int s = 0; s < P.nsteps; ++s) {
                const MmaStep& m = P.msteps[s];
                mbar_wait(full0 + 8 * slot, ph);
                tc_fence_after();
                if (s == 0) V2_STAMP(3, ti);
                if (s == P.nsteps - 1) V2_STAMP(4, ti);
                const uint64_t a_hi = m.a_desc + (uint64_t)(base16 + (uint32_t)slot * (SLOT_BYTES >> 4));
                const uint64_t a_lo = a_hi + m.a_lo_delta;
                const uint64_t b0 = m.b_desc[0] + base16, b1 = m.b_desc[1] + base16;
                const uint32_t d = d0 + m.d_off;
                if (elect_one()) {
                    if (P.acc_mode != ACC_DECONV) {
                        umma_f16(d, a_hi, b0, idesc, m.first);
                        umma_f16(d, a_lo, b0, idesc, 1u);
                        umma_f16(d, a_hi, b0 + wlo16, idesc, 1u);
                        if (m.nj == 2) {
                            umma_f16(d, a_hi + 128, b1, idesc, 1u);
                            umma_f16(d, a_lo + 128, b1, idesc, 1u);
                            umma_f16(d, a_hi + 128, b1 + wlo16, idesc, 1u);
                        }
                    } else {
                        for (uint32_t j = 0; j < m.nj; ++j) {
                            const uint64_t bj = j == 0 ? b0 : b1;
                            for (int a = 0; a < P.nacc; ++a) {
                                const uint32_t da = d + (uint32_t)(a * P.acc_stride);
                                const uint64_t b = bj + (uint32_t)((a * P.N * 128) >> 4);
                                umma_f16(da, a_hi + 128 * j, b, idesc, j == 0 ? m.first : 1u);
                                umma_f16(da, a_lo + 128 * j, b, idesc, 1u);
                                umma_f16(da, a_hi + 128 * j, b + wlo16, idesc, 1u);
                            }
                        }
                    }
                    umma_commit(empty0 + 8 * slot);
                }
                __syncwarp();
                if (++slot == P.nslots) { slot = 0; ph ^= 1; }
            }
            if constexpr (GATED) {
                mbar_wait(gfull0 + 8 * gd, gph);
                tc_fence_after();
                const uint32_t g_base = base + L.gbuf_off + (uint32_t)gd * (P.gate_ch / 32) * SLOT_BYTES;
                if (elect_one()) {
                    for (int u = 0; u < P.gate_ch / 32; ++u)
                        for (int j = 0; j < 2; ++j) {
                            const int k = P.gate_k0 + 32 * u + 16 * j;
                            const uint64_t b = smem_desc_sw128(base + L.w_off + (uint32_t)(k >> 6) * P.nrows * 128u + (uint32_t)((k & 63) >> 4) * 32u);
                            const uint32_t a = g_base + (uint32_t)u * SLOT_BYTES + j * 2048;
                            const uint64_t a_hi = smem_desc_mn_sw128(a, 8192), a_lo = a_hi + (4096 >> 4);
                            const uint32_t first = (P.nsteps == 0 && (u | j) == 0) ? 0u : 1u;
                            umma_f16(d0, a_hi, b, idesc, first);
                            umma_f16(d0, a_lo, b, idesc, 1u); umma_f16(d0, a_hi, b + wlo16, idesc, 1u);
                        }
                    umma_commit(gempty0 + 8 * gd);
                }
                __syncwarp();
                if (++gd == P.gdepth) { gd = 0; gph ^= 1; }
            }
            if (elect_one()) umma_commit(tfull0 + 8 * as);
            V2_STAMP(5, ti);
            __syncwarp();
            if (++as == P.acc_stages) { as = 0; aph ^= 1; }
        }
    } else if (warp >= EPI_WARP0 && warp < GATE_WARP0) {
        // =========================================================================== epilogue
        // 16 warps: lane quarter lq = warp & 3 (hardware: a warp reads TMEM lanes 32*(warp%4)..), column chunks of 16
        // are dealt round-robin to the four warps of a lane quarter.  A 32-column GroupNorm group is two chunks, usually
        // on two different warps: every (warp, group) keeps its own partial (n, S1, S2, pilot), merged by the finalizer.
        const int ew = warp - EPI_WARP0, lq = warp & 3, qtr = ew >> 2;
        const int row = lq * 32 + lane;
        const int nchunks = P.N >> 4;
        const unsigned tiles_per_blk = (unsigned)(P.blk_stride / TILE_M);
        int as = 0; uint32_t aph = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long p = tile * TILE_M + row;
            const unsigned in_blk = ((unsigned)tile % tiles_per_blk) * TILE_M + row;
            const bool valid = in_blk < (unsigned)P.blk_valid;
            const bool pair_valid = (in_blk & ~1u) < (unsigned)P.blk_valid;
            mbar_wait(tfull0 + 8 * as, aph);
            tc_fence_after();
            if (ew == 0 && lane == 0) V2_STAMP(6, (int)((tile - blockIdx.x) / gridDim.x));
            const uint32_t t0 = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(as * P.nacc * P.acc_stride);
#pragma unroll 1
            for (int c = qtr; c < nchunks; c += NWE / 4) {
                const int col0 = c * 16;
                if (P.epi == EPI_STATS_F32) {
                    float v[16];
                    tmem_ld16(t0 + col0, v);
                    const float4* bs4 = reinterpret_cast<const float4*>(sbias + col0);
                    float bsv[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const float4 b = bs4[i]; bsv[4 * i] = b.x; bsv[4 * i + 1] = b.y; bsv[4 * i + 2] = b.z; bsv[4 * i + 3] = b.w; }
                    const int g = c >> 1;
                    float pilot = 0.f;
                    if (g < P.nstat) {                       // this warp's pilot for the group: its first real value
                        const float4 r = red[ew * MAXG + g];
                        pilot = r.x > 0.f ? r.w : __shfl_sync(0xffffffffu, v[0] + bsv[0], 0);
                    }
                    float y[16], s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        y[i] = v[i] + bsv[i];
                        const float d = v[i] + (bsv[i] - pilot);
                        s1 += d; s2 = fmaf(d, d, s2);
                    }
                    if (col0 >= P.store_c0 && col0 < P.store_c1 && valid) {
                        float* o = P.out_f32 + (long long)(col0 - P.store_c0) * P.out_plane + p;
#pragma unroll
                        for (int i = 0; i < 16; ++i) { *o = y[i]; o += P.out_plane; }
                    }
                    if (g < P.nstat) {
                        const float ps = warp_sum(valid ? s1 : 0.f), pq = warp_sum(valid ? s2 : 0.f);
                        const int cnt = __popc(__ballot_sync(0xffffffffu, valid)) * 16;
                        if (lane == 0 && cnt > 0) {
                            float4 r = red[ew * MAXG + g];
                            r.x += (float)cnt; r.y += ps; r.z += pq; r.w = pilot;
                            red[ew * MAXG + g] = r;
                        }
                        __syncwarp();
                    }
                } else if (P.acc_mode == ACC_POOL) {
                    // AvgPool2 after the activation: the four 2x2 phases of this coarse pixel are the four accumulators
                    float y[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] = 0.f;
                    const float* bs = sbias + col0;
#pragma unroll 1
                    for (int a = 0; a < 4; ++a) {
                        float v[16];
                        tmem_ld16(t0 + (uint32_t)(a * P.acc_stride) + col0, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) y[i] += lrelu(v[i] + bs[i], P.slope);
                    }
                    sp16* dst = P.out_hi + ((lane & 1) ? P.out_lo : 0) + (long long)col0 * P.out_plane + (p & ~1ll);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t w = split16(valid ? 0.25f * y[i] : 0.f);
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, w, 1);
                        const uint32_t pk = (lane & 1) ? ((o >> 16) | (w & 0xFFFF0000u)) : ((w & 0xFFFFu) | (o << 16));
                        if (pair_valid) *reinterpret_cast<uint32_t*>(dst) = pk;
                        dst += P.out_plane;
                    }
                } else {
                    const int na = P.acc_mode == ACC_DECONV ? P.nacc : 1;
#pragma unroll 1
                    for (int a = 0; a < na; ++a) {
                        float v[16];
                        tmem_ld16(t0 + (uint32_t)(a * P.acc_stride) + col0, v);
                        const float* bs = sbias + a * P.N + col0;
                        const long long pa = p + (long long)a * P.out_acc_stride;
                        if (P.epi == EPI_LRELU_F32) {
                            if (valid) {
                                long long po = pa;
                                if (P.nchw) {
                                    const unsigned blk = (unsigned)(p / P.nchw_n4p), q = (unsigned)(p - (long long)blk * P.nchw_n4p);
                                    const unsigned qy = q / (unsigned)P.nchw_w4, qx = q - qy * (unsigned)P.nchw_w4, f1 = blk >> 2, f2 = blk & 3;
                                    po = (long long)(4 * qy + 2 * (f2 >> 1) + (f1 >> 1)) * P.nchw_w + (4 * qx + 2 * (f2 & 1) + (f1 & 1));
                                }
                                float* o = P.out_f32 + (long long)col0 * P.out_plane + po;
#pragma unroll
                                for (int i = 0; i < 16; ++i) { if (col0 + i < P.store_c1) *o = lrelu(v[i] + bs[i], P.slope); o += P.out_plane; }
                            }
                        } else {
                            sp16* dst = P.out_hi + ((lane & 1) ? P.out_lo : 0) + (long long)col0 * P.out_plane + (pa & ~1ll);
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const uint32_t w = split16(valid ? lrelu(v[i] + bs[i], P.slope) : 0.f);
                                const uint32_t o = __shfl_xor_sync(0xffffffffu, w, 1);
                                const uint32_t pk = (lane & 1) ? ((o >> 16) | (w & 0xFFFF0000u)) : ((w & 0xFFFFu) | (o << 16));
                                if (pair_valid) *reinterpret_cast<uint32_t*>(dst) = pk;
                                dst += P.out_plane;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * as);
            if (ew == 0 && lane == 0) V2_STAMP(7, (int)((tile - blockIdx.x) / gridDim.x));
            if (++as == P.acc_stages) { as = 0; aph ^= 1; }
        }
    } else if (GATED && warp >= GATE_WARP0) {
        // =========================================================================== gate warps: r*h -> hi/lo operand
        // task = one 16-byte piece (8 pixels of one channel): lanes cover 2 channel rows x 16 pieces (coalesced 256 B rows);
        // 12 warps, two tasks in flight per thread.  sigmoid(a*g + b) = 1 / (1 + 2^(-(a*g+b)*log2 e)): the affine is folded.
        const int gt = tid - GATE_WARP0 * 32;
        constexpr int NGT = NWARP_GATE * 32;
        const int ntask = P.gate_ch * 16;
        int gd = 0; uint32_t gph = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long p0 = tile * TILE_M;
            mbar_wait(gempty0 + 8 * gd, gph ^ 1);
            uint8_t* gb = sm + L.gbuf_off + (size_t)gd * (P.gate_ch / 32) * SLOT_BYTES;
#pragma unroll 1
            for (int t0 = gt; t0 < ntask; t0 += 2 * NGT) {
                uint4 hi[2], lo[2]; float4 g0[2], g1[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int idx = t0 + q * NGT;
                    if (idx < ntask) {
                        const int ch = idx >> 4, j = idx & 15;
                        const long long e = (long long)ch * P.gate_h_plane + p0 + 8 * j;
                        hi[q] = __ldg(reinterpret_cast<const uint4*>(P.gate_h + e));
                        lo[q] = __ldg(reinterpret_cast<const uint4*>(P.gate_h + P.gate_h_lo + e));
                        const float4* gp = reinterpret_cast<const float4*>(P.gate_pre + (long long)ch * P.gate_pre_plane + p0 + 8 * j);
                        g0[q] = __ldg(gp); g1[q] = __ldg(gp + 1);
                    }
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int idx = t0 + q * NGT;
                    if (idx < ntask) {
                        const int ch = idx >> 4, j = idx & 15;
                        const float sc = sgaff[2 * ch], sh = sgaff[2 * ch + 1];          // pre-multiplied by -log2(e)
                        const uint32_t hw[4] = {hi[q].x, hi[q].y, hi[q].z, hi[q].w}, lw[4] = {lo[q].x, lo[q].y, lo[q].z, lo[q].w};
                        const float gv[8] = {g0[q].x, g0[q].y, g0[q].z, g0[q].w, g1[q].x, g1[q].y, g1[q].z, g1[q].w};
                        uint32_t oh[4], ol[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float h0 = lo16_to_f32(hw[u]) + lo16_to_f32(lw[u]), h1 = hi16_to_f32(hw[u]) + hi16_to_f32(lw[u]);
                            const float v0 = __fdividef(h0, 1.0f + ex2_approx(fmaf(gv[2 * u], sc, sh)));
                            const float v1 = __fdividef(h1, 1.0f + ex2_approx(fmaf(gv[2 * u + 1], sc, sh)));
                            split16x2(v0, v1, oh[u], ol[u]);
                        }
                        const int cu = ch & 31;
                        uint8_t* d = gb + (size_t)(ch >> 5) * SLOT_BYTES + (j >> 3) * 8192 + cu * 128 + (((j & 7) ^ (cu & 7)) << 4);
                        *reinterpret_cast<uint4*>(d) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
                        *reinterpret_cast<uint4*>(d + 4096) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(gfull0 + 8 * gd);
            if (++gd == P.gdepth) { gd = 0; gph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (P.dbg && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); P.dbg[9 * 32 + 256 + blockIdx.x] = (long long)t_; }
    if (P.epi == EPI_STATS_F32 && tid < P.nstat) {
        // this CTA's partial of set `tid`: the 16 epilogue warps' sums re-referenced to one pilot (double, no divisions)
        const int g = tid;
        float k0 = 0.f; bool have = false;
        for (int w = 0; w < NWE; ++w) { const float4 r = red[w * MAXG + g]; if (!have && r.x > 0.f) { k0 = r.w; have = true; } }
        double n = 0.0, a1 = 0.0, a2 = 0.0;
        for (int w = 0; w < NWE; ++w) {
            const float4 r = red[w * MAXG + g];
            if (!(r.x > 0.f)) continue;              // warps that only saw padding pixels: pilot may be garbage
            const double nb = (double)r.x, s1 = (double)r.y, s2 = (double)r.z, d = (double)r.w - (double)k0;
            n += nb; a1 += s1 + nb * d; a2 += s2 + 2.0 * d * s1 + nb * d * d;
        }
        P.sink.partial[(size_t)g * P.sink.stride + blockIdx.x] = have ? make_float4((float)n, (float)a1, (float)a2, k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols); }
    if (P.dbg && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMax((unsigned long long*)&P.dbg[8 * 32 + 2], t_); }
    if (P.epi == EPI_STATS_F32 && P.nstat > 0) stats2_finalize_last_cta(P.sink, gridDim.x, gridDim.x, &P.aff);
    if (P.dbg && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMax((unsigned long long*)&P.dbg[8 * 32 + 3], t_); }
}

// ---- weight images: hi and lo bf16 parts of W^T, K-major rows of 128 bytes (64 channels), SWIZZLE_128B
struct WImgSpec {
    const float* W; long long w_ld, w_ks;     // element (row n, channel k) = W[(n % racc)*w_ld + (n / racc)*w_acc + k*w_ks]
    long long w_acc; int racc;                // rows per accumulator block (ConvTranspose phases); racc >= nrows: plain matrix
    int nrows, nrows_valid, K, k_skip;        // k_skip: leading source columns that are not part of the contraction
    unsigned long long off;                   // byte offset of the image in the arena
};
constexpr int WIMG_MAX = 32;
struct WImgBatch { WImgSpec s[WIMG_MAX]; int n; char* base; };

__global__ void __launch_bounds__(256) wimg_kernel(const WImgBatch B) {
    const WImgSpec& S = B.s[blockIdx.y];
    const int nkb = (S.K + 63) / 64;
    const int chunks_per_row = nkb * 8, total = S.nrows * chunks_per_row;
    const size_t lo_off = (size_t)nkb * S.nrows * 128;
    char* img = B.base + S.off;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int n = idx / chunks_per_row, ch = idx % chunks_per_row;
        const int kb = ch >> 3, j = ch & 7, k0 = kb * 64 + j * 8;
        uint32_t h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = k0 + 2 * u + e;
                v[e] = (k < S.K && n < S.nrows_valid)
                           ? __ldg(S.W + (long long)(n % S.racc) * S.w_ld + (long long)(n / S.racc) * S.w_acc + (long long)(k + S.k_skip) * S.w_ks) : 0.f;
            }
            const uint32_t w0 = split16(v[0]), w1 = split16(v[1]);
            h[u] = (w0 & 0xFFFFu) | (w1 << 16);
            l[u] = (w0 >> 16) | (w1 & 0xFFFF0000u);
        }
        const size_t o = (size_t)kb * S.nrows * 128 + (size_t)n * 128 + ((j ^ (n & 7)) << 4);
        *reinterpret_cast<uint4*>(img + o) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(img + lo_off + o) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}
static inline size_t wimg_bytes(int nrows, int K) { return 2 * (size_t)((K + 63) / 64) * nrows * 128; }

}  // namespace v2
}  // namespace urnn
