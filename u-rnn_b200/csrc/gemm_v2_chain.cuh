// gemm_v2_chain.cuh -- the "recompute" schedule of a ConvGRU cell on the second-generation pixel GEMM: the pre-norm gate maps
// never leave the SM.
//
//   sweep A   (gemm_kernel, statistics only, no store)      G = W1 u + b1                       -> GroupNorm-1 statistics
//   sweep B'  (chain_kernel, final = 0)   G_r in tensor memory -> gate warps r*h -> C = W2 [x | e | r*h] + b2 -> GN-2 statistics
//   sweep C'  (chain_kernel, final = 1)   G_z, G_r, C again; epilogue: z, tanh(GN2(C)), blend with h, store h' (split map)
//
// HBM traffic per cell step: 3 x (C_x + C_h) inputs (the 2nd and 3rd read hit L2 when the inputs fit) + F outputs, instead of
// the materialising schedule's extra 2F + 5F fp32 planes (encoder stage 1 at 500 x 500: 144 MB from HBM instead of 672 MB).
// Needs [W1 ; W2] hi + lo resident: 3F x K x 4 bytes of shared memory (F = 64, K = 80: 96 KB).
//
// Per tile: phase 1 MMAs (all columns from x / e; z and r columns from h) -> g1full -> 16 gate warps read G_r from TMEM
// (lane = pixel), r = sigmoid(GN1), multiply the state, re-split into the swizzled operand buffer -> gfull -> phase 2 MMAs
// (C += W2_h (r*h)) -> tfull -> 8 epilogue warps.  The MMA warp issues phase 1 of tile i+1 before phase 2 of tile i, the
// gate buffers and the accumulators are double-buffered: the gate stage of one tile overlaps the epilogue of the previous.
#pragma once
#include "gemm_v2.cuh"

namespace urnn {
namespace v2 {

constexpr int CH_NWE = 8, CH_NWG = 16;
constexpr int CH_GATE_WARP0 = EPI_WARP0 + CH_NWE;
constexpr int NTHREADS_CHAIN = 32 * (EPI_WARP0 + CH_NWE + CH_NWG);

struct ChainParams {
    Step steps[MAX_STEPS]; int step_ncols[MAX_STEPS]; int nsteps;   // TMA units: x / e (all N columns) first, then raw h (first col_c columns)
    int F, gate_k0;                              // gated channels; their position on the weight image's K axis
    const sp16* h; long long h_plane, h_lo;      // split map of the state (global loads: gate warps, blend epilogue)
    const void* wimg; int nkb, nrows;            // rows = N: [W1_z (final only) ; W1_r ; W2]
    int N, col_z, col_r, col_c;                  // accumulator columns of the three F-wide blocks (col_z < 0: absent)
    long long ntot, blk_stride, blk_valid;
    const float* bias_z; const float* bias_r; const float* bias_c;
    const float* scale_z; const float* shift_z; const float* scale_r; const float* shift_r;    // GroupNorm-1 affine (sweep A)
    const float* scale_c; const float* shift_c;  // GroupNorm-2 affine (sweep B'), final only
    int final;
    sp16* out_hi; long long out_lo, out_plane;   // final: h' split map
    int nstat; StatSink2 sink; AffineOut aff;    // !final: GroupNorm-2 statistics of C
    int nslots, gdepth, tmem_cols, acc_stride;
};

struct ChainSmem { uint32_t w_off, ring_off, gbuf_off, bias_off, aff_off, bar_off, red_off, ptab_off, total; };
__host__ __device__ inline ChainSmem chain_smem(int nkb, int nrows, int nslots, int F, int gdepth) {
    ChainSmem s; uint32_t o = 0;
    s.w_off = o; o += 2u * nkb * nrows * 128u;
    s.ring_off = o; o += (uint32_t)nslots * SLOT_BYTES;
    s.gbuf_off = o; o += (uint32_t)gdepth * (F / 32) * SLOT_BYTES;
    s.bias_off = o; o += ((uint32_t)3 * F * 4u + 127u) & ~127u;           // [z | r | c]
    s.aff_off = o; o += ((uint32_t)3 * F * 8u + 127u) & ~127u;            // float2 (scale, shift) pre-folded: [r | z | c]
    s.bar_off = o; o += 512;
    s.red_off = o; o += CH_NWE * MAXG * 16;
    s.ptab_off = o; o += MAX_STEPS * 32;
    s.total = o + 1024;
    return s;
}

__global__ void __launch_bounds__(NTHREADS_CHAIN, 1)
chain_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
             const __grid_constant__ CUtensorMap map2, const ChainParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int F = P.F;
    const ChainSmem L = chain_smem(P.nkb, P.nrows, P.nslots, F, P.gdepth);
    // barriers: full[8] | empty[8] | g1full[2] | tfull[2] | tempty[2] | gfull[2] | gempty[2] | tmem slot
    const uint32_t full0 = base + L.bar_off, empty0 = full0 + 64, g1full0 = empty0 + 64, tfull0 = g1full0 + 16, tempty0 = tfull0 + 16;
    const uint32_t gfull0 = tempty0 + 16, gempty0 = gfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.bar_off + 256);
    float* sbias = reinterpret_cast<float*>(sm + L.bias_off);          // [z | r | c]
    float2* saff = reinterpret_cast<float2*>(sm + L.aff_off);          // [r | z | c]
    float4* red = reinterpret_cast<float4*>(sm + L.red_off);
    ProdEnt* ptab = reinterpret_cast<ProdEnt*>(sm + L.ptab_off);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        for (int s = 0; s < P.nslots; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(g1full0 + 8 * a, 1); mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, CH_NWE);
            mbar_init(gfull0 + 8 * a, CH_NWG); mbar_init(gempty0 + 8 * a, 1);
        }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map0); tma_prefetch_desc(&map1); tma_prefetch_desc(&map2); }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
    for (int i = tid; i < CH_NWE * MAXG; i += NTHREADS_CHAIN) red[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 3 * F; i += NTHREADS_CHAIN) {
        const int b = i / F, c = i - b * F;
        const float* src = b == 0 ? P.bias_z : (b == 1 ? P.bias_r : P.bias_c);
        sbias[i] = src ? __ldg(src + c) : 0.f;
    }
    if (tid < P.nsteps) {
        const Step& st = P.steps[tid];
        ProdEnt pe;
        pe.map = reinterpret_cast<unsigned long long>(st.map == 0 ? (const void*)&map0 : (st.map == 1 ? (const void*)&map1 : (const void*)&map2));
        pe.c0 = st.c0; pe.pix_off = (int)st.pix_off; pe.half_bytes = (uint32_t)st.unit_ch * 256u; pe.pad[0] = pe.pad[1] = pe.pad[2] = 0;
        ptab[tid] = pe;
    }
    {
        const int nvec = 2 * P.nkb * P.nrows * 8;
        const char* src = reinterpret_cast<const char*>(P.wimg);
        const int rot = (int)((blockIdx.x * 37u) % (unsigned)(nvec >> 5)) << 5;
        for (int i = tid; i < nvec; i += NTHREADS_CHAIN) {
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
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // folded affines (outputs of the previous sweeps): value = scale * (acc + bias) + shift, pre-multiplied for ex2:
    // sigmoid(x) = 1 / (1 + 2^(-x log2 e)),  tanh(x) = 1 - 2 / (1 + 2^(2 x log2 e))
    for (int i = tid; i < 3 * F; i += NTHREADS_CHAIN) {
        const int b = i / F, c = i - b * F;                  // 0: r, 1: z, 2: c
        float sc = 0.f, sh = 0.f, bias = 0.f, k = -1.4426950408889634f;
        if (b == 0) { sc = __ldg(P.scale_r + c); sh = __ldg(P.shift_r + c); bias = sbias[F + c]; }
        else if (b == 1) { if (P.col_z >= 0) { sc = __ldg(P.scale_z + c); sh = __ldg(P.shift_z + c); bias = sbias[c]; } }
        else if (P.final) { sc = __ldg(P.scale_c + c); sh = __ldg(P.shift_c + c); bias = sbias[2 * F + c]; k = 2.8853900817779268f; }
        saff[i] = make_float2(k * sc, k * fmaf(sc, bias, sh));
    }
    __syncthreads();
    const long long ntiles = P.ntot / TILE_M;
    const uint32_t wlo16 = ((uint32_t)P.nkb * P.nrows * 128u) >> 4;

    if (warp == 0 || warp == 3) {
        // =========================================================================== TMA producers (even / odd units)
        if (lane == 0) {
            const int me = warp == 0 ? 0 : 1;
            int slot = me; uint32_t ph = 0;
            if (slot >= P.nslots) { slot -= P.nslots; ph ^= 1; }
            int s = me;
            long long tile = blockIdx.x;
            while (s >= P.nsteps) { s -= P.nsteps; tile += gridDim.x; }
            while (tile < ntiles) {
                const ProdEnt pe = ptab[s];
                const int px = (int)(tile * TILE_M) + pe.pix_off;
                mbar_wait(empty0 + 8 * slot, ph ^ 1);
                const uint32_t dst = base + L.ring_off + (uint32_t)slot * SLOT_BYTES, bar = full0 + 8 * slot;
                mbar_arrive_expect_tx(bar, 2u * pe.half_bytes);
                tma_load_3d(dst, reinterpret_cast<const void*>(pe.map), bar, px, pe.c0, 0, L2_EVICT_NORMAL);
                tma_load_3d(dst + pe.half_bytes, reinterpret_cast<const void*>(pe.map), bar, px + 64, pe.c0, 0, L2_EVICT_NORMAL);
                slot += 2; if (slot >= P.nslots) { slot -= P.nslots; ph ^= 1; }
                s += 2; while (s >= P.nsteps) { s -= P.nsteps; tile += gridDim.x; }
            }
        }
    } else if (warp == 1) {
        // =========================================================================== MMA issuer (phase 2 lags one tile)
        const uint32_t idesc_c = instr_desc_16(F, SPLIT_FMT);
        int slot = 0; uint32_t ph = 0;
        // is there a unit that initialises the C columns in phase 1?
        const bool c_init = P.nsteps > 0 && P.step_ncols[0] == P.N;
        auto phase2 = [&](long long j) {
            const int gd = (int)(j % P.gdepth), as = (int)(j & 1);
            mbar_wait(gfull0 + 8 * gd, (uint32_t)((j / P.gdepth) & 1));
            tc_fence_after();
            const uint32_t g_base = base + L.gbuf_off + (uint32_t)gd * (F / 32) * SLOT_BYTES;
            const uint32_t d = tmem_base + (uint32_t)(as * P.acc_stride) + (uint32_t)P.col_c;
            if (elect_one()) {
                for (int u = 0; u < F / 32; ++u)
                    for (int jj = 0; jj < 2; ++jj) {
                        const int k = P.gate_k0 + 32 * u + 16 * jj;
                        const uint64_t b = smem_desc_sw128(base + L.w_off + (uint32_t)(k >> 6) * P.nrows * 128u + (uint32_t)P.col_c * 128u + (uint32_t)((k & 63) >> 4) * 32u);
                        const uint32_t a = g_base + (uint32_t)u * SLOT_BYTES + jj * 2048;
                        const uint64_t a_hi = smem_desc_mn_sw128(a, 8192), a_lo = a_hi + (4096 >> 4);
                        const uint32_t first = (!c_init && (u | jj) == 0) ? 0u : 1u;
                        umma_f16(d, a_hi, b, idesc_c, first);
                        umma_f16(d, a_lo, b, idesc_c, 1u);
                        umma_f16(d, a_hi, b + wlo16, idesc_c, 1u);
                    }
                umma_commit(gempty0 + 8 * gd);
                umma_commit(tfull0 + 8 * as);
            }
            __syncwarp();
        };
        long long i = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
            const int as = (int)(i & 1);
            mbar_wait(tempty0 + 8 * as, (uint32_t)(((i >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)(as * P.acc_stride);
            for (int s = 0; s < P.nsteps; ++s) {
                const int unit_ch = P.steps[s].unit_ch, kg = P.steps[s].kglob, ncols = P.step_ncols[s];
                mbar_wait(full0 + 8 * slot, ph);
                tc_fence_after();
                const uint32_t a_base = base + L.ring_off + (uint32_t)slot * SLOT_BYTES;
                const uint64_t a0 = smem_desc_mn_sw128(a_base, (uint32_t)unit_ch * 256u);
                const uint32_t a_lo_delta = ((uint32_t)unit_ch * 128u) >> 4;
                const uint64_t b0 = smem_desc_sw128(base + L.w_off + (uint32_t)(kg >> 6) * P.nrows * 128u + (uint32_t)((kg & 63) >> 4) * 32u);
                const int kg1 = kg + 16;
                const uint64_t b1 = smem_desc_sw128(base + L.w_off + (uint32_t)(kg1 >> 6) * P.nrows * 128u + (uint32_t)((kg1 & 63) >> 4) * 32u);
                const uint32_t idesc = instr_desc_16(ncols, SPLIT_FMT);
                if (elect_one()) {
                    for (int j = 0; j < (unit_ch >> 4); ++j) {
                        const uint64_t a_hi = a0 + (uint32_t)(j * 128), a_lo = a_hi + a_lo_delta, b_hi = j == 0 ? b0 : b1;
                        umma_f16(d0, a_hi, b_hi, idesc, (s | j) == 0 ? 0u : 1u);
                        umma_f16(d0, a_lo, b_hi, idesc, 1u);
                        umma_f16(d0, a_hi, b_hi + wlo16, idesc, 1u);
                    }
                    umma_commit(empty0 + 8 * slot);
                }
                __syncwarp();
                if (++slot == P.nslots) { slot = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(g1full0 + 8 * as);
            __syncwarp();
            if (i > 0) phase2(i - 1);
        }
        if (i > 0) phase2(i - 1);
    } else if (warp >= EPI_WARP0 && warp < CH_GATE_WARP0) {
        // =========================================================================== epilogue (8 warps)
        const int ew = warp - EPI_WARP0, lq = warp & 3, half = ew >> 2;
        const int row = lq * 32 + lane;
        const int nchunks = F >> 4;
        const unsigned tiles_per_blk = (unsigned)(P.blk_stride / TILE_M);
        const unsigned short* hp = reinterpret_cast<const unsigned short*>(P.h);
        long long i = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
            const int as = (int)(i & 1);
            const long long p = tile * TILE_M + row;
            const unsigned in_blk = ((unsigned)tile % tiles_per_blk) * TILE_M + row;
            const bool valid = in_blk < (unsigned)P.blk_valid;
            const bool pair_valid = (in_blk & ~1u) < (unsigned)P.blk_valid;
            uint32_t hw[16];                                  // state of my pixel, 16 channels: hi | lo << 16 (final sweep)
            auto load_h = [&](int c) {
                const long long hoff = (long long)(16 * c) * P.h_plane + p;
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    hw[q] = (uint32_t)__ldg(hp + hoff + (long long)q * P.h_plane) | ((uint32_t)__ldg(hp + P.h_lo + hoff + (long long)q * P.h_plane) << 16);
            };
            if (P.final && half < nchunks) load_h(half);      // in flight while the accumulator is still being produced
            mbar_wait(tfull0 + 8 * as, (uint32_t)((i >> 1) & 1));
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(as * P.acc_stride);
#pragma unroll 1
            for (int c = half; c < nchunks; c += 2) {
                float vc[16];
                tmem_ld16(t0 + (uint32_t)(P.col_c + 16 * c), vc);
                if (!P.final) {
                    const float* bs = sbias + 2 * F + 16 * c;
                    const int g = c >> 1;
                    const float4 r0 = red[ew * MAXG + g];
                    const float pilot = r0.x > 0.f ? r0.w : __shfl_sync(0xffffffffu, vc[0] + bs[0], 0);
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int q = 0; q < 16; ++q) { const float d = vc[q] + (bs[q] - pilot); s1 += d; s2 = fmaf(d, d, s2); }
                    const float ps = warp_sum(valid ? s1 : 0.f), pq = warp_sum(valid ? s2 : 0.f);
                    const int cnt = __popc(__ballot_sync(0xffffffffu, valid)) * 16;
                    if (lane == 0 && cnt > 0) {
                        float4 r = red[ew * MAXG + g];
                        r.x += (float)cnt; r.y += ps; r.z += pq; r.w = pilot;
                        red[ew * MAXG + g] = r;
                    }
                    __syncwarp();
                } else {
                    float vz[16];
                    tmem_ld16(t0 + (uint32_t)(P.col_z + 16 * c), vz);
                    sp16* dst = P.out_hi + ((lane & 1) ? P.out_lo : 0) + (long long)(16 * c) * P.out_plane + (p & ~1ll);
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int ch = 16 * c + q;
                        const float2 az = saff[F + ch], ac = saff[2 * F + ch];
                        const float hv = lo16_to_f32(hw[q] & 0xFFFFu) + lo16_to_f32(hw[q] >> 16);
                        const float z = __fdividef(1.0f, 1.0f + ex2_approx(fmaf(az.x, vz[q], az.y)));
                        const float t = 1.0f - __fdividef(2.0f, 1.0f + ex2_approx(fmaf(ac.x, vc[q], ac.y)));
                        const uint32_t w = split16(valid ? fmaf(z, t - hv, hv) : 0.f);
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, w, 1);
                        const uint32_t pk = (lane & 1) ? ((o >> 16) | (w & 0xFFFF0000u)) : ((w & 0xFFFFu) | (o << 16));
                        if (pair_valid) *reinterpret_cast<uint32_t*>(dst) = pk;
                        dst += P.out_plane;
                    }
                    if (c + 2 < nchunks) load_h(c + 2);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * as);
        }
    } else if (warp >= CH_GATE_WARP0) {
        // =========================================================================== gate warps (16): TMEM G_r -> r*h -> operand
        const int gw = warp - CH_GATE_WARP0, lq = warp & 3;
        const int row = lq * 32 + lane;
        const unsigned short* hp = reinterpret_cast<const unsigned short*>(P.h);
        long long i = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
            const int as = (int)(i & 1), gd = (int)(i % P.gdepth);
            const long long p = tile * TILE_M + row;
            uint32_t hw[16];
            auto load_h = [&](int c) {
                const long long hoff = (long long)(16 * c) * P.h_plane + p;
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    hw[q] = (uint32_t)__ldg(hp + hoff + (long long)q * P.h_plane) | ((uint32_t)__ldg(hp + P.h_lo + hoff + (long long)q * P.h_plane) << 16);
            };
            load_h(gw >> 2);                                  // in flight while phase 1 of this tile runs
            mbar_wait(g1full0 + 8 * as, (uint32_t)((i >> 1) & 1));
            tc_fence_after();
            mbar_wait(gempty0 + 8 * gd, (uint32_t)(((i / P.gdepth) & 1) ^ 1));
            uint8_t* gb = sm + L.gbuf_off + (size_t)gd * (F / 32) * SLOT_BYTES;
            const uint32_t t0 = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(as * P.acc_stride) + (uint32_t)P.col_r;
            // my pixel inside the operand tile: half (row >> 6), 16-byte piece ((row & 63) >> 3), element (row & 7)
            const uint32_t px_off = (uint32_t)(row >> 6) * 8192u + (uint32_t)(row & 7) * 2u;
            const uint32_t piece = (uint32_t)((row & 63) >> 3);
#pragma unroll 1
            for (int c = gw >> 2; c < (F >> 4); c += CH_NWG / 4) {
                float v[16];
                tmem_ld16(t0 + (uint32_t)(16 * c), v);
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int ch = 16 * c + q, cu = ch & 31;
                    const float2 a = saff[ch];
                    const float hv = lo16_to_f32(hw[q] & 0xFFFFu) + lo16_to_f32(hw[q] >> 16);
                    const uint32_t w = split16(__fdividef(hv, 1.0f + ex2_approx(fmaf(a.x, v[q], a.y))));
                    uint8_t* d = gb + (size_t)(ch >> 5) * SLOT_BYTES + px_off + (uint32_t)cu * 128u + ((piece ^ (uint32_t)(cu & 7)) << 4);
                    *reinterpret_cast<unsigned short*>(d) = (unsigned short)(w & 0xFFFFu);
                    *reinterpret_cast<unsigned short*>(d + 4096) = (unsigned short)(w >> 16);
                }
                if (c + CH_NWG / 4 < (F >> 4)) load_h(c + CH_NWG / 4);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(gfull0 + 8 * gd);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (!P.final && tid < P.nstat) {
        const int g = tid;
        float k0 = 0.f; bool have = false;
        for (int w = 0; w < CH_NWE; ++w) { const float4 r = red[w * MAXG + g]; if (!have && r.x > 0.f) { k0 = r.w; have = true; } }
        double n = 0.0, a1 = 0.0, a2 = 0.0;
        for (int w = 0; w < CH_NWE; ++w) {
            const float4 r = red[w * MAXG + g];
            if (!(r.x > 0.f)) continue;
            const double nb = (double)r.x, s1 = (double)r.y, s2 = (double)r.z, d = (double)r.w - (double)k0;
            n += nb; a1 += s1 + nb * d; a2 += s2 + 2.0 * d * s1 + nb * d * d;
        }
        P.sink.partial[(size_t)g * P.sink.stride + blockIdx.x] = have ? make_float4((float)n, (float)a1, (float)a2, k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols); }
    if (!P.final && P.nstat > 0) stats2_finalize_last_cta(P.sink, gridDim.x, gridDim.x, &P.aff);
}

}  // namespace v2
}  // namespace urnn
