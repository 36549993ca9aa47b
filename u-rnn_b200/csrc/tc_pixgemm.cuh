// tc_pixgemm.cuh -- tcgen05 "pixel GEMM" for sm_100a:  D[128 pixels x NOUT] = X^T[128 x K] * W^T[K x NOUT]
// with bf16 operands (rounded on the fly from the fp32 / bf16 NCHW maps) and fp32 accumulation in tensor memory.
//
// One persistent CTA per SM, 28 warps (7 warpgroups), warp-specialised.  Activations must be rounded fp32 -> bf16 (and,
// for the candidate GEMM, multiplied by the reset gate) before the tensor core sees them, so a SIMT stage always sits
// between HBM and the MMA operand.  The operand is stored MN-major (pixel-contiguous, SWIZZLE_128B) exactly like the
// NCHW source: no transposition.  The channel axis is a virtual concatenation of up to three maps.
//
// BULK = true (default; every fp32 map, bf16 maps with planes padded to whole tiles):
//   warps 25-26 loaders   : stream the raw channel rows of a unit (32 channels x 128 pixels) into a staging ring with
//                           16-byte cp.async (8-byte for the 2x2 pooling gather, 4-byte for planes that are only 4-byte
//                           aligned); completion is collected per unit by an mbarrier, no registers are held, 4-6 units
//                           (64-96 KB) are in flight per SM.  (1-D TMA bulk copies were tried first: a 512-byte row per
//                           UBLKCP is too small, 11 GB/s per SM.)
//   warps 16-23 converters: two groups of 4 warps, alternating units: staging -> (reset gate) -> bf16 -> operand slot
//   warp  24    MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128, N=NOUT, K=16 per instruction, tcgen05.commit
//   warps 0-15  epilogue  : tcgen05.ld (TMEM lane = pixel, column = output channel); bias, addend map, GroupNorm
//                           statistics, LeakyReLU / pooling / 2x2 scatter; bf16 rows leave through a per-warp transpose tile
//                           as 16-byte pieces
// BULK = false (bf16 maps without padded planes; URNN_BULK=0): warps 8-23 are SIMT producers that load from global
//   memory themselves (64-channel units, two groups of 8 warps), warps 0-7 the epilogue.
//
// Pipelines: staging ring (raw_full / raw_empty), operand ring (full / empty), double-buffered accumulator in TMEM
// (tmem_full / tmem_empty).  Waits park in hardware (mbarrier.try_wait with a suspend-time hint), arrivals are
// warp-elected, register budgets are re-partitioned per role with setmaxnreg.  Weights stay resident in shared memory;
// the time-step driver hands every GEMM a pre-converted bf16 image (wimg_kernel), stand-alone calls convert in the prologue.
#pragma once
#include <cuda_bf16.h>
#include "urnn_common.cuh"

namespace urnn {
namespace tc {

constexpr int TILE_M = 128;           // pixels per tile (TMEM lanes)
constexpr int KBLK = 64;              // channels per ring slot
constexpr int STAGE_BYTES = TILE_M * KBLK * 2;
constexpr int NPROD = 512;            // producer threads
constexpr int NEPI = 256;             // epilogue threads (simt mode): warp w owns TMEM lanes 32*(w%4).. and column groups w/4, w/4+2, ..
constexpr int NEPI_BULK = 512;        // bulk mode: warps 8-15 join the epilogue (column groups w/4, w/4+4), warps 16-23 convert
constexpr int NLOADW = 2;             // bulk mode: loader warps (cp.async row copies), alternating units
constexpr int NTHREADS = NEPI + NPROD + 128;   // 28 warps = 7 whole warpgroups (setmaxnreg works per warpgroup); warp 27 idles
constexpr int MMA_WARP = (NEPI + NPROD) / 32;
constexpr int LOAD_WARP0 = MMA_WARP + 1;    // bulk mode: first loader warp
constexpr int UNIT_K = 32;            // bulk mode: channels per unit
constexpr int RAW_ROW = 512;          // bulk mode: staging bytes per channel row (128 fp32; bf16 rows use the first 256)
constexpr int RAW_SLOT = UNIT_K * RAW_ROW;            // 16 KB
constexpr int GRAW_SLOT = UNIT_K * 256;               // reset-gate pre-activations (bf16) of a unit
constexpr int A_SLOT_BULK = TILE_M * UNIT_K * 2;      // 8 KB operand slot
constexpr int MAX_RAW = 6, MAX_ASLOT = 4;
constexpr int EPI_STAGE = 1024;       // per epilogue warp: 16 channels x 32 pixels of bf16, transposes thread-per-pixel <-> 16-byte rows
constexpr int MAXG = 8;               // NOUT <= 256 -> at most 8 groups of 32 output channels
constexpr int MAXKB = 5;              // K <= 320 channels per contraction
constexpr size_t SMEM_CAP = 229376;   // dynamic shared memory budget (227 KB opt-in limit minus static use)

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// A bare try_wait loop polls every ~25-40 cycles (the default hardware time-out is short, and nanosleep(32) returns
// almost immediately): with 16 waiting warps that was 43-53 % of all issued instructions (ncu source view,
// profiles/) and starved the warps doing real work.  The suspend-time hint lets the hardware park the thread until the
// phase completes (or the hint expires), so a wait costs a handful of instructions.
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (done) break;
    }
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only); bytes past `src_bytes` (0..16) are zero-filled and
// not read.  Completion is collected per thread by cp_async_arrive.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
// 8-byte variant (L1-allocating; .cg only exists for 16 bytes): the 2x2 pooling gather moves (dx0, dx1) pairs
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
// 4-byte variant for fp32 rows that are only 4-byte aligned (e.g. 125 x 125 maps: planes of 15625 floats)
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
// the mbarrier receives one (pre-counted) arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// Register re-partitioning between warp roles (per warpgroup of 4 warps).  Launch: 28 warps x 72 = 64512 registers.
// epilogue warps take 80 (their 32-value accumulator slice + staging addresses spilled at 72), the MMA / loader
// warpgroup gives back to 40:  bulk mode 16x80 + 8x72 + 4x40 (x32 lanes) = 64512; simt mode 8x80 + 16x72 + 4x40.
__device__ __forceinline__ void regs_epilogue() { asm volatile("setmaxnreg.inc.sync.aligned.u32 80;" ::: "memory"); }
__device__ __forceinline__ void regs_converter() {}   // converters keep the launch allocation (72)
__device__ __forceinline__ void regs_control() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 128 B, 8-row atoms 1024 B apart): weights
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);       // start address
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// MN-major (pixel-contiguous) A tile, 128-byte swizzle: 1024-byte atoms of 8 K-rows x 64 MN elements;
// leading byte offset = distance between atoms along MN (8192), stride byte offset = along K (1024).
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo = 8192) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D fp32, A bf16 MN-major, B bf16 K-major, M=128, N=nout
__device__ __forceinline__ uint32_t instr_desc_bf16(int nout) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(nout >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t w, float& a, float& b) {
    a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xFFFF0000u);
}

// 8 consecutive elements starting at element offset `off` of an fp32 (kind 0) or bf16 (kind 1) map; `nvalid` of them
// are inside the map.  The widest aligned vector access available is used.
__device__ __forceinline__ void load8(const void* base, int kind, long off, int nvalid, float (&v)[8]) {
    if (kind == 0) {
        const float* s = reinterpret_cast<const float*>(base) + off;
        if (nvalid == 8 && (reinterpret_cast<uintptr_t>(s) & 15) == 0) {
            const float4 lo = __ldg(reinterpret_cast<const float4*>(s));
            const float4 hi = __ldg(reinterpret_cast<const float4*>(s) + 1);
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (u < nvalid) ? __ldg(s + u) : 0.f;
        }
    } else {
        const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(base) + off;
        const uintptr_t a = reinterpret_cast<uintptr_t>(s);
        if (nvalid == 8 && (a & 15) == 0) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(s));
            unpack_bf16x2(w.x, v[0], v[1]); unpack_bf16x2(w.y, v[2], v[3]);
            unpack_bf16x2(w.z, v[4], v[5]); unpack_bf16x2(w.w, v[6], v[7]);
        } else if (nvalid == 8 && (a & 7) == 0) {
            const uint2 w0 = __ldg(reinterpret_cast<const uint2*>(s));
            const uint2 w1 = __ldg(reinterpret_cast<const uint2*>(s) + 1);
            unpack_bf16x2(w0.x, v[0], v[1]); unpack_bf16x2(w0.y, v[2], v[3]);
            unpack_bf16x2(w1.x, v[4], v[5]); unpack_bf16x2(w1.y, v[6], v[7]);
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (u < nvalid) ? __bfloat162float(s[u]) : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------- parameters
// Virtual channel concatenation of up to three NCHW maps (fp32: kind 0, bf16: kind 1).  The segment `gate_seg`
// (or none: -1) is multiplied by sigmoid(gate_pre[gate_ch0 + c] * gate_scale[..] + gate_shift[..]) (reset gate;
// gate_pre is a bf16 map).
struct Segs {
    const void* src[3]; int cend[3]; int kind[3];
    long plane[3];                  // elements between channel planes of each map (>= N; internal bf16 maps are padded)
    long gate_plane;                // same for gate_pre
    int gate_seg; int gate_ch0;
    const __nv_bfloat16* gate_pre; const float* gate_scale; const float* gate_shift;
};

struct GemmParams {
    Segs seg;
    // weights: row n < nrow1 is W[n*w_ld + k]; row n >= nrow1 is W2[(n-nrow1)*w2_ld + k] for k < k2 and 0 beyond
    const float* W; long w_ld; long w_ks; int nrow1;   // element (n,k) = W[n*w_ld + k*w_ks]
    const float* W2; long w2_ld; int k2;
    const float* bias; int nbias;            // bias[n] for n < nbias, zero beyond
    int NOUT, K, N;                          // output channels (multiple of 32, <= 256), reduction size, pixels
    __nv_bfloat16* out; long out_plane;      // bf16 NCHW destination of (acc + bias [+ addend])
    const __nv_bfloat16* addend;             // optional bf16 map [NOUT][plane] added before store / statistics
    int nstat;                               // leading 32-channel groups whose (sum, sumsq) go to `sink`
    // stem epilogues (EPI_LRELU / EPI_POOL / EPI_DECONV)
    int nout_store;                          // columns actually stored (<= NOUT; NOUT is padded to 32)
    float* out_f32;                          // fp32 destination instead of `out` when non-null (EPI_LRELU)
    float slope;                             // LeakyReLU slope
    int img_w;                               // EPI_POOL: input width (producer gathers 2x2 quads); EPI_DECONV: input width
    int n_base;                              // EPI_DECONV: global column offset of this launch (n = co*4 + dy*2 + dx)
    StatSink sink; AffineOut aff;
    int nstage;                              // simt mode: K-block ring depth (64-channel slots)
    int bulk, nraw, na;                      // bulk mode: on/off, staging ring depth, operand ring depth (32-channel slots)
    int reverse;                             // walk the tiles from the last to the first (L2 reuse between sweeps)
    const void* wimg;                        // optional pre-converted weight image (bf16, swizzled, exactly the shared-memory layout):
                                             // the prologue then copies nkb*NOUT*128 bytes instead of converting fp32 weights
    int bias_mma;                            // bulk mode: the bias rides in the GEMM as two extra channels of ones (weights bf16 hi + lo
                                             // parts of the bias), so the epilogue has no bias loads / adds
    int out_vec;                             // bf16 output (and addend) planes are padded to whole tiles and 16-byte aligned:
                                             // the epilogue moves them as 16-byte row pieces through a per-warp staging tile
    int tmem_cols;                           // power of two >= 2 * NOUT
    volatile unsigned* dbg;                  // optional host-mapped progress / trace words (bring-up only)
};

enum { EPI_GN = 0, EPI_LRELU = 1, EPI_POOL = 2, EPI_DECONV = 3 };

// 8 tile slots = 2 pooling quads (slot = 4*quad + 2*dy + dx) gathered from an fp32 map of width img_w
__device__ __forceinline__ void load8_quads(const float* plane_base, long q0, long nquads, int w2, int img_w, float (&v)[8]) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const long Q = q0 + t;
        if (Q < nquads) {
            const long qy = Q / w2, qx = Q % w2;
            const float* b = plane_base + (2 * qy) * (long)img_w + 2 * qx;
            const float2 top = __ldg(reinterpret_cast<const float2*>(b));
            const float2 bot = __ldg(reinterpret_cast<const float2*>(b + img_w));
            v[4 * t + 0] = top.x; v[4 * t + 1] = top.y; v[4 * t + 2] = bot.x; v[4 * t + 3] = bot.y;
        } else {
            v[4 * t + 0] = v[4 * t + 1] = v[4 * t + 2] = v[4 * t + 3] = 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------- the kernel
// BULK is a template parameter (not P.bulk) so that each instantiation carries only one producer mode: the kernel is
// instruction-cache bound on grids with one tile per CTA (first pass through every role's code).
template <bool GATED, int EPI, bool BULK>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_gn_kernel(const GemmParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;           // swizzle atoms need 1024-byte alignment
    uint8_t* sm = smem_raw + (base - raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NOUT = P.NOUT, K = P.K, N = P.N;
    if (P.dbg && blockIdx.x == 0 && tid == 0) { unsigned t_; asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t_)); P.dbg[148 * 16 + 6 * 64] = t_; }
    const int Kb = K + (P.bias_mma ? 2 : 0);                // + the two bias channels
    const int Kp = (Kb + 15) & ~15;                         // padded to the MMA K
    const int nkb = (Kp + KBLK - 1) / KBLK;                 // K blocks per tile
    const int last_k = Kp - (nkb - 1) * KBLK;               // valid (padded) channels in the last block
    const int wblk_bytes = NOUT * 128;
    constexpr bool bulk = BULK;
    const int nslot = bulk ? P.na : P.nstage;               // operand ring depth (8 KB slots in bulk mode, 16 KB otherwise)
    const int nu = (Kp + UNIT_K - 1) / UNIT_K;              // bulk mode: 32-channel units per tile
    const uint32_t raw_slot_bytes = GATED ? (RAW_SLOT + GRAW_SLOT) : RAW_SLOT;
    const uint32_t w_off = 0;
    const uint32_t a_off = w_off + (uint32_t)nkb * wblk_bytes;
    const uint32_t raw_off = a_off + (uint32_t)(bulk ? P.na * A_SLOT_BULK : P.nstage * STAGE_BYTES);
    const uint32_t stg_off = raw_off + (bulk ? (uint32_t)P.nraw * raw_slot_bytes : 0u);
    const uint32_t bias_off = stg_off + (uint32_t)((bulk ? NEPI_BULK : NEPI) / 32) * EPI_STAGE;
    const uint32_t bar_off = bias_off + 1024;
    const uint32_t tab_off = bar_off + 512;                  // per-channel source table: MAXKB*64 entries of 32 bytes
    // barriers: full[8], empty[8], tmem_full[2], tmem_empty[2], (tmem slot), raw_full[8], raw_empty[8]
    const uint32_t full0 = base + bar_off, empty0 = full0 + 64;
    const uint32_t tfull0 = empty0 + 64, tempty0 = tfull0 + 16;
    const uint32_t rfull0 = base + bar_off + 256, rempty0 = rfull0 + 64;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + bar_off + 176);
    float* sbias = reinterpret_cast<float*>(sm + bias_off);
    struct ChanEnt { const char* ptr; const char* gptr; float sc, sh; int meta; int pad; };   // meta: 1 valid, 2 bf16, 4 gated
    ChanEnt* ctab = reinterpret_cast<ChanEnt*>(sm + tab_off);
    __shared__ float red[2][4][MAXG];   // [sum|sumsq][lane quarter][group]; each (quarter, group) has one owner warp

    // allow the next kernel in the stream to begin its own prologue as soon as SMs free up (PDL)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // ---- one-time setup: barriers, TMEM, resident weights (fp32 -> bf16, swizzled K-major rows)
    if (tid < 2 * 4 * MAXG) (&red[0][0][0])[tid] = 0.f;
    if (tid == 0) {
        // Arrivals are per WARP (lane 0 after __syncwarp), not per thread: every arrive is a SYNCS event that wakes the
        // warps parked in NANOSLEEP.SYNCS, and 128-512 arrivals per unit kept them polling (31 % of executed instructions).
        // One producer / converter group per operand slot: 8 warps (simt) or 4 warps (bulk).
        for (int s = 0; s < nslot; ++s) { mbar_init(full0 + 8 * s, bulk ? 4 : NPROD / 64); mbar_init(empty0 + 8 * s, 1); }
        if (bulk) for (int s = 0; s < P.nraw; ++s) { mbar_init(rfull0 + 8 * s, 32); mbar_init(rempty0 + 8 * s, 4); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, (bulk ? NEPI_BULK : NEPI) / 32); }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
    if constexpr (EPI == EPI_DECONV) {
        for (int i = tid; i < (NOUT >> 2); i += NTHREADS) sbias[i] = (4 * i < P.nout_store) ? __ldg(P.bias + ((P.n_base >> 2) + i)) : 0.f;
    } else {
        for (int i = tid; i < NOUT; i += NTHREADS) sbias[i] = (i < P.nbias) ? __ldg(P.bias + i) : 0.f;
    }
    if (P.wimg != nullptr) {
        // image prepared once per step by wimg_kernel: plain asynchronous 16-byte copies, rotated per CTA so that the
        // 148 CTAs do not ask the same L2 lines at the same moment
        const int nvec = nkb * (wblk_bytes >> 4);
        const int rot = (int)((blockIdx.x * 41u) % (unsigned)(nvec >> 5)) << 5;
        const char* src = reinterpret_cast<const char*>(P.wimg);
        for (int i = tid; i < nvec; i += NTHREADS) {
            const int j = i + rot >= nvec ? i + rot - nvec : i + rot;
            cp_async16(base + w_off + (uint32_t)j * 16u, src + (size_t)j * 16u, 16u);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        // no prepared image (stand-alone operator calls): convert the fp32 weights here.  Deliberately compact code --
        // the time-step driver always supplies images, and kernel size matters (instruction cache).
        const int chunks_per_row = nkb * 8, total = NOUT * chunks_per_row;
#pragma unroll 1
        for (int idx = tid; idx < total; idx += NTHREADS) {
            const int n = idx / chunks_per_row, ch = idx % chunks_per_row;
            const int kb = ch >> 3, j = ch & 7, k0 = kb * KBLK + j * 8;
            const bool second = n >= P.nrow1;
            const float* wrow = second ? (P.W2 + (long)(n - P.nrow1) * P.w2_ld) : (P.W + (long)n * P.w_ld);
            const int klim = second ? (P.k2 < K ? P.k2 : K) : K;
            const long ks = second ? 1 : P.w_ks;
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (k0 + u < klim && n < P.nout_store) ? __ldg(wrow + (long)(k0 + u) * ks) : 0.f;
            if (P.bias_mma && n < P.nbias && k0 <= K + 1 && k0 + 8 > K) {      // columns K, K+1: bias = hi + lo in bf16
                const float b = __ldg(P.bias + n);
                const float hi = __bfloat162float(__float2bfloat16(b));
#pragma unroll
                for (int u = 0; u < 8; ++u) { if (k0 + u == K) v[u] = hi; if (k0 + u == K + 1) v[u] = b - hi; }
            }
            const uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            *reinterpret_cast<uint4*>(sm + w_off + kb * wblk_bytes + n * 128 + ((j ^ (n & 7)) << 4)) = pk;
        }
    }
    // where every reduction channel lives: plane base pointer, dtype, optional reset-gate map and folded GN affine
    // (gate_scale / gate_shift are outputs of the previous kernel: read them only after the dependency wait)
    for (int k = tid; k < nkb * KBLK; k += NTHREADS) {
        ChanEnt e; e.ptr = nullptr; e.gptr = nullptr; e.sc = 0.f; e.sh = 0.f; e.meta = 0; e.pad = 0;
        if (k < K) {
            const Segs& S = P.seg;
            const int sg = (k < S.cend[0]) ? 0 : ((k < S.cend[1]) ? 1 : 2);
            const int cc = k - (sg == 0 ? 0 : S.cend[sg - 1]);
            e.ptr = reinterpret_cast<const char*>(S.src[sg]) + (long)cc * S.plane[sg] * (S.kind[sg] ? 2 : 4);
            e.meta = 1 | (S.kind[sg] ? 2 : 0);
            if (GATED && sg == S.gate_seg) {
                e.gptr = reinterpret_cast<const char*>(S.gate_pre) + (long)(S.gate_ch0 + cc) * S.gate_plane * 2;
                e.meta |= 4 | ((S.gate_ch0 + cc) << 8);
            }
        } else if (k < Kb) {
            e.meta = 8;                                      // bias channel: constant one
        }
        ctab[k] = e;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above reads only parameters; from here on the previous kernel's outputs are consumed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if constexpr (GATED) {
        for (int k = tid; k < nkb * KBLK; k += NTHREADS) {
            const int m = ctab[k].meta;
            if (m & 4) { ctab[k].sc = __ldg(P.seg.gate_scale + (m >> 8)); ctab[k].sh = __ldg(P.seg.gate_shift + (m >> 8)); }
        }
        __syncthreads();
    }
#define TC_DBG(slot, val) do { if (P.dbg && (tid & 31) == 0) P.dbg[blockIdx.x * 16 + (slot)] = (val); } while (0)
#define TC_TRACE(role, idx) do { if (P.dbg && blockIdx.x == 0 && (tid & 31) == 0 && (idx) < 64) { unsigned t_; asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t_)); P.dbg[148 * 16 + (role) * 64 + (idx)] = t_; } } while (0)
    TC_DBG(0, 0x100u | tmem_base);
    TC_TRACE(0, 0);
    const int ntiles = (N + TILE_M - 1) / TILE_M;
    const int acc_stride = P.tmem_cols >> 1;                // columns per accumulator stage

    const int nepw = bulk ? NEPI_BULK / 32 : NEPI / 32;       // epilogue warps
    if (warp >= nepw && warp < MMA_WARP) {
        // =========================================================================== producers
        // Shared layout of a ring slot (canonical UMMA MN-major, SWIZZLE_128B): 1024-byte atoms of 8 channels x 64
        // pixels; atom(mblk, kblk) at mblk*8192 + kblk*1024; inside: channel (k&7)*128 B, 16-byte chunk
        // j = (pixel%64)/8 stored at chunk position j ^ (k&7).  Units (tile, K block) are streamed two deep.
        // simt: warps 8-23, two groups of 256 threads; bulk: warps 16-23, two converter groups of 128 threads
        const int pt = bulk ? tid - NEPI_BULK : tid - NEPI;
        const Segs& S = P.seg;
        // Two independent producer groups of 256 threads; group g fills the units (tile, K block) with index u = g (mod 2).
        // A thread holds NO loads in flight when it publishes its unit: fence.proxy.async compiles to MEMBAR.ALL.CTA,
        // which waits for every outstanding load of the thread, so a per-thread software prefetch would be serialised
        // by the fence.  Memory-level parallelism comes from the two groups (and the 8-deep ring) instead.
        const int grp = bulk ? pt >> 7 : pt >> 8, tl = bulk ? pt & 127 : pt & 255;
        const int ch0 = tl >> 4, px8 = tl & 15;               // simt: chunk c handles channel-in-block ch0 + 16c, pixels px8*8..+7
        constexpr int NCH = 4;
        int trace_i = 0;
        bool aligned = (N & 3) == 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            aligned = aligned && ((reinterpret_cast<uintptr_t>(S.src[i]) & (S.kind[i] ? 7 : 15)) == 0) && ((S.plane[i] & 3) == 0);
        if (GATED) aligned = aligned && ((reinterpret_cast<uintptr_t>(S.gate_pre) & 7) == 0) && ((S.gate_plane & 3) == 0);
        // chunk c lives 2 K-atoms (2048 B) after chunk c-1: channel ch0 + 16c
        const uint32_t soff0 = (uint32_t)((px8 >> 3) * 8192 + (ch0 >> 3) * 1024 + (ch0 & 7) * 128 + (((px8 & 7) ^ (ch0 & 7)) << 4));
        const ChanEnt* myent = ctab + ch0;
        const long px_off = (long)px8 * 8;
        struct Buf { float v[NCH][8]; uint2 g[GATED ? NCH : 1][2]; float sc[GATED ? NCH : 1], sh[GATED ? NCH : 1]; int meta[NCH]; };

        auto ld8 = [&](const char* ptr, bool bf16, long tile, bool whole, int nvalid, float (&v)[8]) {
            if (!bf16) {
                const float* s = reinterpret_cast<const float*>(ptr) + tile * TILE_M + px8 * 8;
                if (whole) {
                    const float4 lo = __ldg(reinterpret_cast<const float4*>(s));
                    const float4 hi = __ldg(reinterpret_cast<const float4*>(s) + 1);
                    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = (u < nvalid) ? __ldg(s + u) : 0.f;
                }
            } else {
                const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(ptr) + tile * TILE_M + px8 * 8;
                if (whole) {
                    const uint2 w0 = __ldg(reinterpret_cast<const uint2*>(s));
                    const uint2 w1 = __ldg(reinterpret_cast<const uint2*>(s) + 1);
                    unpack_bf16x2(w0.x, v[0], v[1]); unpack_bf16x2(w0.y, v[2], v[3]);
                    unpack_bf16x2(w1.x, v[4], v[5]); unpack_bf16x2(w1.y, v[6], v[7]);
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = (u < nvalid) ? __bfloat162float(s[u]) : 0.f;
                }
            }
        };
        // fast path: the whole tile is inside the map and every plane is vector-aligned -> no per-element control flow
        auto issue_fast = [&](int tile, int kb, Buf& b) {
            const long eoff = (long)tile * TILE_M + px_off;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const ChanEnt e = myent[kb * KBLK + 16 * c];
                b.meta[c] = e.meta;
                if (e.meta & 1) {
                    if (e.meta & 2) {
                        const uint2* s = reinterpret_cast<const uint2*>(e.ptr + eoff * 2);
                        const uint2 w0 = __ldg(s), w1 = __ldg(s + 1);
                        unpack_bf16x2(w0.x, b.v[c][0], b.v[c][1]); unpack_bf16x2(w0.y, b.v[c][2], b.v[c][3]);
                        unpack_bf16x2(w1.x, b.v[c][4], b.v[c][5]); unpack_bf16x2(w1.y, b.v[c][6], b.v[c][7]);
                    } else {
                        const float4* s = reinterpret_cast<const float4*>(e.ptr + eoff * 4);
                        const float4 lo = __ldg(s), hi = __ldg(s + 1);
                        b.v[c][0] = lo.x; b.v[c][1] = lo.y; b.v[c][2] = lo.z; b.v[c][3] = lo.w;
                        b.v[c][4] = hi.x; b.v[c][5] = hi.y; b.v[c][6] = hi.z; b.v[c][7] = hi.w;
                    }
                    if constexpr (GATED) {
                        if (e.meta & 4) {
                            const uint2* s = reinterpret_cast<const uint2*>(e.gptr + eoff * 2);
                            b.g[c][0] = __ldg(s); b.g[c][1] = __ldg(s + 1);
                            b.sc[c] = e.sc; b.sh[c] = e.sh;
                        }
                    }
                }
            }
        };
        auto issue_slow = [&](int tile, int kb, Buf& b) {
            const long p = (long)tile * TILE_M + px8 * 8;
            const int nvalid = (p >= N) ? 0 : ((N - p >= 8) ? 8 : (int)(N - p));
            const bool whole = aligned && nvalid == 8;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const ChanEnt e = myent[kb * KBLK + 16 * c];
                b.meta[c] = (nvalid > 0) ? e.meta : 0;
                if (b.meta[c] & 1) {
                    if constexpr (EPI == EPI_POOL) {
                        load8_quads(reinterpret_cast<const float*>(e.ptr), p >> 2, (long)N >> 2, P.img_w >> 1, P.img_w, b.v[c]);
                    } else {
                        ld8(e.ptr, (e.meta & 2) != 0, tile, whole, nvalid, b.v[c]);
                    }
                    if constexpr (GATED) {
                        if (e.meta & 4) {
                            float gv[8];
                            ld8(e.gptr, true, tile, whole, nvalid, gv);
                            b.g[c][0] = make_uint2(pack_bf16(gv[0], gv[1]), pack_bf16(gv[2], gv[3]));
                            b.g[c][1] = make_uint2(pack_bf16(gv[4], gv[5]), pack_bf16(gv[6], gv[7]));
                            b.sc[c] = e.sc; b.sh[c] = e.sh;
                        }
                    }
                }
            }
        };
        auto commit = [&](int kb, int stage, Buf& b) {
            uint8_t* st = sm + a_off + stage * STAGE_BYTES;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                uint4 pk = make_uint4(0u, 0u, 0u, 0u);      // channels past K were tabulated as invalid: stored as zeros
                if (b.meta[c] & 1) {
                    if constexpr (GATED) {
                        if (b.meta[c] & 4) {
                            float gv[8];
                            unpack_bf16x2(b.g[c][0].x, gv[0], gv[1]); unpack_bf16x2(b.g[c][0].y, gv[2], gv[3]);
                            unpack_bf16x2(b.g[c][1].x, gv[4], gv[5]); unpack_bf16x2(b.g[c][1].y, gv[6], gv[7]);
#pragma unroll
                            for (int u = 0; u < 8; ++u) b.v[c][u] *= sigmoid_fast(fmaf(gv[u], b.sc[c], b.sh[c]));
                        }
                    }
                    pk = make_uint4(pack_bf16(b.v[c][0], b.v[c][1]), pack_bf16(b.v[c][2], b.v[c][3]),
                                    pack_bf16(b.v[c][4], b.v[c][5]), pack_bf16(b.v[c][6], b.v[c][7]));
                }
                *reinterpret_cast<uint4*>(st + soff0 + c * 2048) = pk;
            }
            fence_proxy_async();                             // generic-proxy stores -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * stage);
            if (warp == NEPI / 32 && kb == nkb - 1) { TC_TRACE(1, trace_i); ++trace_i; }
        };

        if (!bulk) {
            Buf b;
            for (long u = grp;; u += 2) {
                const long ti = u / nkb;
                const int kb = (int)(u - ti * nkb);
                long tile = blockIdx.x + ti * gridDim.x;
                if (tile >= ntiles) break;
                if (P.reverse) tile = ntiles - 1 - tile;
                const int stage = (int)(u % P.nstage);
                const uint32_t phase = (uint32_t)((u / P.nstage) & 1);
                if (EPI != EPI_POOL && aligned && (tile + 1) * TILE_M <= N) issue_fast((int)tile, kb, b);
                else issue_slow((int)tile, kb, b);
                mbar_wait<64>(empty0 + 8 * stage, phase ^ 1);
                commit(kb, stage, b);
            }
        } else {
            regs_converter();
            // ---- bulk mode: staging (raw rows landed by cp.async) -> bf16 swizzled operand slot.  Unit = 32 channels x 128
            // pixels; a thread (ch0 = 0..7) converts 8 pixels of channels ch0 + 8c, c = 0..3.  Operand slot: 1024-byte
            // atoms of 8 channels x 64 pixels, atom(mblk, kblk) at mblk*4096 + kblk*1024 -> chunk c is K-atom c.
            const uint32_t so0 = (uint32_t)((px8 >> 3) * 4096 + ch0 * 128 + (((px8 & 7) ^ ch0) << 4));
            // fp32 rows: a thread's 32 bytes are fetched as two 16-byte pieces; odd groups of four threads fetch the
            // upper piece first so that a quarter-warp touches all 32 banks (conflict-free LDS.128)
            const bool swp = ((px8 >> 2) & 1) != 0;
            for (long u = grp;; u += 2) {
                const long ti = u / nu;
                const int kb = (int)(u - ti * nu);
                long tile = blockIdx.x + ti * gridDim.x;
                if (tile >= ntiles) break;
                if (P.reverse) tile = ntiles - 1 - tile;
                const int rs = (int)(u % P.nraw), as = (int)(u % P.na);
                const uint32_t rph = (uint32_t)((u / P.nraw) & 1), aph = (uint32_t)((u / P.na) & 1);
                const long pbase = tile * TILE_M + px_off;
                const int nv = (pbase >= N) ? 0 : ((N - pbase >= 8) ? 8 : (int)(N - pbase));   // valid pixels of my 8
                const uint8_t* rawp = sm + raw_off + rs * raw_slot_bytes;
                mbar_wait<96>(rfull0 + 8 * rs, rph);
                uint4 pk[4];
                // fast paths: my four channels are of one kind and the tile is whole -> all shared loads are issued
                // back to back, without per-channel control flow (a converter warp is latency-bound, not issue-bound)
                const ChanEnt* ent = ctab + kb * UNIT_K + ch0;
                const int m0 = ent[0].meta, m1 = ent[8].meta, m2 = ent[16].meta, m3 = ent[24].meta;
                const int kind_and = m0 & m1 & m2 & m3 & 15, kind_or = (m0 | m1 | m2 | m3) & 15;
                if (kind_and == kind_or && nv == 8 && (kind_or == 1 || kind_or == 3 || (GATED && kind_or == 5))) {
                    if (kind_or == 3) {
                        uint4 w[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) w[c] = *reinterpret_cast<const uint4*>(rawp + (ch0 + 8 * c) * RAW_ROW + px8 * 16);
#pragma unroll
                        for (int c = 0; c < 4; ++c) pk[c] = w[c];                 // already bf16: a plain copy
                    } else if (!GATED || kind_or == 1) {
                        float4 f[4][2];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint8_t* r = rawp + (ch0 + 8 * c) * RAW_ROW + px8 * 32;
                            f[c][0] = *reinterpret_cast<const float4*>(r + (swp ? 16 : 0));
                            f[c][1] = *reinterpret_cast<const float4*>(r + (swp ? 0 : 16));
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4 lo = swp ? f[c][1] : f[c][0], hi = swp ? f[c][0] : f[c][1];
                            pk[c] = make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
                        }
                    } else {
                        // reset-gated rows: two channels at a time (register budget)
#pragma unroll
                        for (int c2 = 0; c2 < 4; c2 += 2) {
                            float4 f[2][2]; uint4 g[2]; float2 aff[2];
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const uint8_t* r = rawp + (ch0 + 8 * (c2 + c)) * RAW_ROW + px8 * 32;
                                f[c][0] = *reinterpret_cast<const float4*>(r + (swp ? 16 : 0));
                                f[c][1] = *reinterpret_cast<const float4*>(r + (swp ? 0 : 16));
                                g[c] = *reinterpret_cast<const uint4*>(rawp + RAW_SLOT + (ch0 + 8 * (c2 + c)) * 256 + px8 * 16);
                                aff[c] = *reinterpret_cast<const float2*>(&ent[8 * (c2 + c)].sc);
                            }
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                float gv[8];
                                unpack_bf16x2(g[c].x, gv[0], gv[1]); unpack_bf16x2(g[c].y, gv[2], gv[3]);
                                unpack_bf16x2(g[c].z, gv[4], gv[5]); unpack_bf16x2(g[c].w, gv[6], gv[7]);
                                const float4 lo = swp ? f[c][1] : f[c][0], hi = swp ? f[c][0] : f[c][1];
                                const float a = aff[c].x, b = aff[c].y;
                                pk[c2 + c] = make_uint4(
                                    pack_bf16(lo.x * sigmoid_fast(fmaf(gv[0], a, b)), lo.y * sigmoid_fast(fmaf(gv[1], a, b))),
                                    pack_bf16(lo.z * sigmoid_fast(fmaf(gv[2], a, b)), lo.w * sigmoid_fast(fmaf(gv[3], a, b))),
                                    pack_bf16(hi.x * sigmoid_fast(fmaf(gv[4], a, b)), hi.y * sigmoid_fast(fmaf(gv[5], a, b))),
                                    pack_bf16(hi.z * sigmoid_fast(fmaf(gv[6], a, b)), hi.w * sigmoid_fast(fmaf(gv[7], a, b))));
                            }
                        }
                    }
                } else
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int ch = ch0 + 8 * c;
                    const ChanEnt* e = ctab + kb * UNIT_K + ch;
                    const int meta = e->meta;
                    float v[8];
                    pk[c] = make_uint4(0u, 0u, 0u, 0u);
                    if (meta & 1) {
                        if (meta & 2) {
                            const uint4 w = *reinterpret_cast<const uint4*>(rawp + ch * RAW_ROW + px8 * 16);
                            unpack_bf16x2(w.x, v[0], v[1]); unpack_bf16x2(w.y, v[2], v[3]);
                            unpack_bf16x2(w.z, v[4], v[5]); unpack_bf16x2(w.w, v[6], v[7]);
                        } else {
                            const uint8_t* r = rawp + ch * RAW_ROW + px8 * 32;
                            const float4 f0 = *reinterpret_cast<const float4*>(r + (swp ? 16 : 0));
                            const float4 f1 = *reinterpret_cast<const float4*>(r + (swp ? 0 : 16));
                            const float4 lo = swp ? f1 : f0, hi = swp ? f0 : f1;
                            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
                        }
                        if constexpr (GATED) {
                            if (meta & 4) {
                                const uint4 g = *reinterpret_cast<const uint4*>(rawp + RAW_SLOT + ch * 256 + px8 * 16);
                                float gv[8];
                                unpack_bf16x2(g.x, gv[0], gv[1]); unpack_bf16x2(g.y, gv[2], gv[3]);
                                unpack_bf16x2(g.z, gv[4], gv[5]); unpack_bf16x2(g.w, gv[6], gv[7]);
                                const float sc = e->sc, sh = e->sh;
#pragma unroll
                                for (int q = 0; q < 8; ++q) v[q] *= sigmoid_fast(fmaf(gv[q], sc, sh));
                            }
                        }
                        if (nv < 8) {                        // last tile: rows beyond the map hold stale staging bytes
#pragma unroll
                            for (int q = 0; q < 8; ++q) v[q] = (q < nv) ? v[q] : 0.f;
                        }
                        pk[c] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    } else if (meta & 8) {
                        pk[c] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);   // bf16 ones
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(rempty0 + 8 * rs); // staging slot consumed (values are in registers)
                mbar_wait<64>(empty0 + 8 * as, aph ^ 1);
                uint8_t* st = sm + a_off + as * A_SLOT_BULK;
#pragma unroll
                for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(st + so0 + c * 1024) = pk[c];
                fence_proxy_async();                         // generic-proxy stores -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) mbar_arrive(full0 + 8 * as);
                if (warp == NEPI_BULK / 32 && kb == nu - 1) { TC_TRACE(1, trace_i); ++trace_i; }
            }
        }
    } else if (warp >= MMA_WARP) {
      regs_control();
      if (warp >= LOAD_WARP0 && warp < LOAD_WARP0 + NLOADW) {
        // =========================================================================== loaders (bulk mode)
        // Loader warp l streams the units u = l (mod NLOADW): per channel row one warp-wide 16-byte cp.async (512 B of an
        // fp32 row; a bf16 row, 256 B, uses lanes 0-15), plus the row of reset-gate pre-activations for gated channels.
        // Nothing is waited for here: the copies of up to `nraw` units stay in flight and each thread's completion is
        // collected by the unit's mbarrier (cp.async.mbarrier.arrive.noinc).
        if (bulk) {
            for (long u = warp - LOAD_WARP0;; u += NLOADW) {
                const long ti = u / nu;
                const int kb = (int)(u - ti * nu);
                long tile = blockIdx.x + ti * gridDim.x;
                if (tile >= ntiles) break;
                if (P.reverse) tile = ntiles - 1 - tile;
                const int rs = (int)(u % P.nraw);
                const uint32_t rph = (uint32_t)((u / P.nraw) & 1);
                const long p0 = tile * TILE_M;
                const int nvalid = (N - p0 >= TILE_M) ? TILE_M : (int)(N - p0);
                // valid bytes of my 16-byte chunk in an fp32 row (tail tile: the rest is zero-filled)
                const int vb = nvalid * 4 - lane * 16;
                const uint32_t fbytes = vb >= 16 ? 16u : (vb > 0 ? (uint32_t)vb : 0u);
                const uint32_t dst = base + raw_off + rs * raw_slot_bytes;
                mbar_wait<96>(rempty0 + 8 * rs, rph ^ 1);
                // rows are addressed arithmetically from the segment description (kernel parameters, uniform registers):
                // a table lookup per row would put two dependent shared-memory round trips in front of every copy
                const Segs& S = P.seg;
                const int k_lo = kb * UNIT_K, k_hi = (k_lo + UNIT_K < K) ? k_lo + UNIT_K : K;
                if constexpr (EPI == EPI_POOL) {
                    // tile slot = 4*quad + 2*dy + dx: lane q gathers quad tile*32 + q of every channel with two 8-byte
                    // copies (the dx pair of the upper and of the lower image row); single fp32 source map
                    const long Q = tile * 32 + lane, nquads = (long)N >> 2;
                    const int w2 = P.img_w >> 1;
                    const long qy = Q / w2, qx = Q - qy * w2;
                    const uint32_t nb = Q < nquads ? 8u : 0u;
                    const long rowb = S.plane[0] * 4;
                    const char* src = reinterpret_cast<const char*>(S.src[0]) + (long)k_lo * rowb + (nb ? ((2 * qy) * (long)P.img_w + 2 * qx) * 4 : 0);
                    const long down = (long)P.img_w * 4;
                    uint32_t d = dst + lane * 16;
#pragma unroll 4
                    for (int c = k_lo; c < k_hi; ++c) { cp_async8(d, src, nb); cp_async8(d + 8, src + down, nb); src += rowb; d += RAW_ROW; }
                    cp_async_arrive(rfull0 + 8 * rs);
                    continue;
                }
                int cs = 0;
#pragma unroll
                for (int sg = 0; sg < 3; ++sg) {
                    const int ce = S.cend[sg];
                    const int lo = cs > k_lo ? cs : k_lo, hi = ce < k_hi ? ce : k_hi;
                    if (lo < hi) {
                        const bool b16 = S.kind[sg] != 0;
                        const long rowb = S.plane[sg] * (b16 ? 2 : 4);
                        const char* src = reinterpret_cast<const char*>(S.src[sg]) + (long)(lo - cs) * rowb + p0 * (b16 ? 2 : 4) + lane * 16;
                        uint32_t d = dst + (uint32_t)(lo - k_lo) * RAW_ROW + lane * 16;
                        const uint32_t nb = b16 ? 16u : fbytes;          // bf16 rows live in padded planes: always whole
                        const bool row16 = b16 || (((reinterpret_cast<uintptr_t>(S.src[sg]) | (uintptr_t)rowb) & 15) == 0);
                        if (row16) {
                            if (!b16 || lane < 16) {
#pragma unroll 4
                                for (int c = lo; c < hi; ++c) { cp_async16(d, src, nb); src += rowb; d += RAW_ROW; }
                            }
                        } else {
                            // rows are only 4-byte aligned: four 4-byte copies per lane and row (pixels lane, lane+32, ...)
                            const char* s4 = src - lane * 12;            // = row base + lane * 4
                            const uint32_t d4 = d - lane * 12;
                            uint32_t v4[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) v4[q] = (lane + 32 * q < nvalid) ? 4u : 0u;
                            for (int c = lo; c < hi; ++c) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) cp_async4(d4 + (uint32_t)(c - lo) * RAW_ROW + q * 128, s4 + (long)(c - lo) * rowb + q * 128, v4[q]);
                            }
                        }
                        if (GATED && sg == S.gate_seg && lane < 16) {
                            const long growb = S.gate_plane * 2;
                            const char* g = reinterpret_cast<const char*>(S.gate_pre) + (long)(S.gate_ch0 + lo - cs) * growb + p0 * 2 + lane * 16;
                            uint32_t gd = dst + RAW_SLOT + (uint32_t)(lo - k_lo) * 256 + lane * 16;
#pragma unroll 4
                            for (int c = lo; c < hi; ++c) { cp_async16(gd, g, 16u); g += growb; gd += 256; }
                        }
                    }
                    cs = ce;
                }
                cp_async_arrive(rfull0 + 8 * rs);
            }
        }
      } else if (warp == MMA_WARP) {
        // =========================================================================== MMA issuer
        // every lane follows the pipeline (waits), lane 0 alone issues tcgen05.mma / tcgen05.commit
        const uint32_t idesc = instr_desc_bf16(NOUT);
        int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait<32>(tempty0 + 8 * as, aphase ^ 1);     // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * acc_stride);
            const int nunits = bulk ? nu : nkb;
            for (int kb = 0; kb < nunits; ++kb) {
                mbar_wait<20>(full0 + 8 * stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    if (bulk) {
                        const uint32_t a_addr = base + a_off + stage * A_SLOT_BULK;
                        const uint32_t b_addr = base + w_off + (kb >> 1) * wblk_bytes + (kb & 1) * 64;
                        const int left = (Kp - kb * UNIT_K) >> 4;
                        const int nmma = left < 2 ? left : 2;
                        for (int j = 0; j < nmma; ++j)
                            umma_f16(d_tmem, smem_desc_mn_sw128(a_addr + j * 2048, 4096), smem_desc_sw128(b_addr + j * 32), idesc,
                                     (kb | j) ? 1u : 0u);
                    } else {
                        const uint32_t a_addr = base + a_off + stage * STAGE_BYTES;
                        const uint32_t b_addr = base + w_off + kb * wblk_bytes;
                        const int nmma = ((kb == nkb - 1) ? last_k : KBLK) >> 4;
                        for (int j = 0; j < nmma; ++j)
                            umma_f16(d_tmem, smem_desc_mn_sw128(a_addr + j * 2048), smem_desc_sw128(b_addr + j * 32), idesc,
                                     (kb | j) ? 1u : 0u);
                    }
                    umma_commit(empty0 + 8 * stage);         // frees the ring slot when these MMAs retire
                }
                __syncwarp();
                if (++stage == nslot) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) umma_commit(tfull0 + 8 * as);     // accumulator complete -> epilogue
            __syncwarp();
            TC_TRACE(2, (tile - (int)blockIdx.x) / (int)gridDim.x);
            as ^= 1; if (as == 0) aphase ^= 1;
        }
      }
    } else {
        // =========================================================================== epilogue (warps 0-7; 0-15 in bulk mode)
        regs_epilogue();
        // warp w: TMEM lanes 32*(w&3).. (pixels), 32-column groups g = (w>>2), (w>>2)+2, ... (output channels)
        int as = 0; uint32_t aphase = 0;
        const int ng = NOUT >> 5;
        const int lq = warp & 3, ghalf = warp >> 2, gstep = nepw >> 2;   // column groups ghalf, ghalf + gstep, ...
        const int row = lq * 32 + lane;
        for (int vt = blockIdx.x; vt < ntiles; vt += gridDim.x) {
            const int tile = P.reverse ? ntiles - 1 - vt : vt;
            const long p = (long)tile * TILE_M + row;
            const bool valid = p < N;
            const long pc = valid ? p : 0;
            // the addend map (r-independent candidate part) does not depend on this tile's MMA: fetch it first.
            // out_vec: as four 16-byte row pieces per lane (row c = (lane>>2) + 8j of the group, pixels 8*(lane&3)..+7 of
            // this warp's 32), transposed to thread-per-pixel through the warp's staging tile; else element by element.
            uint4 adraw[EPI == EPI_GN ? 4 : 1];
            // 16-byte row pieces: lane -> row (lane>>1) of a 16-channel half group, pixels 16*(lane&1).. of this warp's 32
            const long vrow0 = (long)tile * TILE_M + lq * 32 + (lane & 1) * 16;
            const int vrow = lane >> 1;
            if constexpr (EPI == EPI_GN) {
                if (P.addend != nullptr && ghalf < ng && P.out_vec) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const uint4* a = reinterpret_cast<const uint4*>(P.addend + (long)(ghalf * 32 + hh * 16 + vrow) * P.out_plane + vrow0);
                        adraw[2 * hh] = __ldg(a); adraw[2 * hh + 1] = __ldg(a + 1);
                    }
                }
            }
            mbar_wait<32>(tfull0 + 8 * as, aphase);
            tc_fence_after();
            if (warp == 0) TC_TRACE(3, (vt - (int)blockIdx.x) / (int)gridDim.x);
            const uint32_t t_addr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(as * acc_stride);
            // not unrolled: one copy of the group code keeps the kernel small (it is instruction-cache sensitive on grids
            // with one tile per CTA); the per-group statistics are therefore reduced and accumulated right away
#pragma unroll 1
            for (int gi = 0; gi < MAXG / 2; ++gi) {
                const int g = ghalf + gstep * gi;
                if (g < ng) {
                    float v[32];
                    tmem_ld32(t_addr + g * 32, v);
                    if constexpr (EPI == EPI_GN) {
                        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                        const float* bs = sbias + g * 32;
                        if (P.out_vec) {
                            uint8_t* stg = sm + stg_off + warp * EPI_STAGE;          // [16 channels][32 pixels] bf16, warp-private
                            const uint32_t vo = (uint32_t)(vrow * 64 + (lane & 1) * 32);
                            if (P.addend != nullptr && gi > 0) {
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    const uint4* a = reinterpret_cast<const uint4*>(P.addend + (long)(g * 32 + hh * 16 + vrow) * P.out_plane + vrow0);
                                    adraw[2 * hh] = __ldg(a); adraw[2 * hh + 1] = __ldg(a + 1);
                                }
                            }
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                if (P.addend != nullptr) {
                                    *reinterpret_cast<uint4*>(stg + vo) = adraw[2 * hh];
                                    *reinterpret_cast<uint4*>(stg + vo + 16) = adraw[2 * hh + 1];
                                    __syncwarp();
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        v[16 * hh + j] += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(stg + j * 64 + lane * 2));
                                    __syncwarp();
                                }
                                if (!P.bias_mma) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[16 * hh + j] += bs[16 * hh + j];
                                }
#pragma unroll
                                for (int j = 0; j < 16; j += 2) {
                                    const int i = 16 * hh + j;
                                    const float y0 = v[i], y1 = v[i + 1];
                                    *reinterpret_cast<__nv_bfloat16*>(stg + j * 64 + lane * 2) = __float2bfloat16(y0);
                                    *reinterpret_cast<__nv_bfloat16*>(stg + (j + 1) * 64 + lane * 2) = __float2bfloat16(y1);
                                    s0 += y0; s1 += y1; q0 = fmaf(y0, y0, q0); q1 = fmaf(y1, y1, q1);
                                }
                                __syncwarp();
                                // planes are padded to whole tiles: pixels past the map are written too (never read as data)
                                uint4* o = reinterpret_cast<uint4*>(P.out + (long)(g * 32 + hh * 16 + vrow) * P.out_plane + vrow0);
                                o[0] = *reinterpret_cast<const uint4*>(stg + vo);
                                o[1] = *reinterpret_cast<const uint4*>(stg + vo + 16);
                                __syncwarp();
                            }
                        } else {
                            __nv_bfloat16* o = P.out + (long)(g * 32) * P.out_plane + pc;
                            if (P.addend != nullptr) {
                                const __nv_bfloat16* ad = P.addend + (long)(g * 32) * P.out_plane + pc;
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] += __bfloat162float(ad[(long)i * P.out_plane]);
                            }
#pragma unroll
                            for (int i = 0; i < 32; i += 2) {
                                const float y0 = v[i] + bs[i], y1 = v[i + 1] + bs[i + 1];
                                if (valid) { o[0] = __float2bfloat16(y0); o[P.out_plane] = __float2bfloat16(y1); }
                                o += 2 * P.out_plane;
                                s0 += y0; s1 += y1; q0 = fmaf(y0, y0, q0); q1 = fmaf(y1, y1, q1);
                            }
                        }
                        // GroupNorm partials of this CTA: lanes -> (lane quarter, group) owner warp, accumulated over tiles
                        const float ps = warp_sum(valid ? s0 + s1 : 0.f), pq = warp_sum(valid ? q0 + q1 : 0.f);
                        if (lane == 0) { red[0][lq][g] += ps; red[1][lq][g] += pq; }
                    } else if constexpr (EPI == EPI_LRELU) {
                        // y = LeakyReLU(acc + bias) -> NCHW (bf16 or fp32)
                        const float* bs = sbias + g * 32;
                        if (P.out_vec && P.out_f32 == nullptr) {
                            uint8_t* stg = sm + stg_off + warp * EPI_STAGE;
                            const uint32_t vo = (uint32_t)(vrow * 64 + (lane & 1) * 32);
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                if (!P.bias_mma) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[16 * hh + j] += bs[16 * hh + j];
                                }
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    *reinterpret_cast<__nv_bfloat16*>(stg + j * 64 + lane * 2) = __float2bfloat16(lrelu(v[16 * hh + j], P.slope));
                                __syncwarp();
                                const int n = g * 32 + hh * 16 + vrow;
                                if (n < P.nout_store) {
                                    uint4* o = reinterpret_cast<uint4*>(P.out + (long)n * P.out_plane + vrow0);
                                    o[0] = *reinterpret_cast<const uint4*>(stg + vo);
                                    o[1] = *reinterpret_cast<const uint4*>(stg + vo + 16);
                                }
                                __syncwarp();
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int n = g * 32 + i;
                                const float y = lrelu(v[i] + bs[i], P.slope);
                                if (valid && n < P.nout_store) {
                                    if (P.out_f32) P.out_f32[(long)n * P.out_plane + p] = y;
                                    else P.out[(long)n * P.out_plane + p] = __float2bfloat16(y);
                                }
                            }
                        }
                    } else if constexpr (EPI == EPI_POOL) {
                        // tile slot = 4*quad + sub-pixel: AvgPool2 = mean over 4 consecutive lanes, after the activation
                        const float* bs = sbias + g * 32;
                        const long Q = p >> 2;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float y = lrelu(v[i] + bs[i], P.slope);
                            y += __shfl_xor_sync(0xffffffffu, y, 1);
                            y += __shfl_xor_sync(0xffffffffu, y, 2);
                            const int n = g * 32 + i;
                            if (valid && (lane & 3) == 0 && n < P.nout_store) {
                                if (P.out_f32) P.out_f32[(long)n * P.out_plane + Q] = 0.25f * y;
                                else P.out[(long)n * P.out_plane + Q] = __float2bfloat16(0.25f * y);
                            }
                        }
                    } else {
                        // ConvTranspose2d(k2,s2): column n = co*4 + dy*2 + dx of pixel (y,x) -> out[co][2y+dy][2x+dx]
                        const long py = pc / P.img_w, px = pc % P.img_w;
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const int n = P.n_base + g * 32 + i;            // dx = 0 (i even), partner i+1 has dx = 1
                            const int co = n >> 2, dy = (n >> 1) & 1;
                            const float b = sbias[(g * 32 + i) >> 2];
                            const float y0 = lrelu(v[i] + b, P.slope), y1 = lrelu(v[i + 1] + b, P.slope);
                            if (valid && g * 32 + i < P.nout_store) {
                                const long off = (long)co * P.out_plane + (2 * py + dy) * (2L * P.img_w) + 2 * px;
                                if (P.out_f32) *reinterpret_cast<float2*>(P.out_f32 + off) = make_float2(y0, y1);
                                else *reinterpret_cast<uint32_t*>(P.out + off) = pack_bf16(y0, y1);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * as);
            if (warp == 0) TC_TRACE(4, (vt - (int)blockIdx.x) / (int)gridDim.x);
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    }
    tc_fence_before();
    if (tid == 0) TC_TRACE(5, 0);
    __syncthreads();
    if (EPI == EPI_GN && tid < P.nstat) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { a += red[0][w][tid]; b += red[1][w][tid]; }
        P.sink.partial[(size_t)tid * P.sink.stride + blockIdx.x] = make_float2(a, b);
    }
    if (warp == MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols); }
    if constexpr (EPI == EPI_GN) stats_finalize_last_cta(P.sink, gridDim.x, gridDim.x, &P.aff);
}

// dynamic shared memory needed for a given problem; 0 if it cannot fit
static inline size_t gemm_smem_fixed(int NOUT, int K, bool bulk) {
    const int Kp = (K + 15) & ~15, nkb = (Kp + KBLK - 1) / KBLK;
    return 1024 /*align slack*/ + (size_t)nkb * NOUT * 128 + (size_t)((bulk ? NEPI_BULK : NEPI) / 32) * EPI_STAGE + 1024 /*bias*/ +
           512 /*barriers*/ + (size_t)nkb * KBLK * 32 /*source table*/;
}
static inline size_t gemm_smem_bytes(int NOUT, int K, int* nstage_out) {
    const int Kp = (K + 15) & ~15, nkb = (Kp + KBLK - 1) / KBLK;
    const size_t fixed = gemm_smem_fixed(NOUT, K, false);
    const size_t cap = SMEM_CAP;
    if (fixed + 2 * STAGE_BYTES > cap || nkb > MAXKB) return 0;
    int ns = (int)((cap - fixed) / STAGE_BYTES);
    if (ns > 8) ns = 8;
    if (nstage_out) *nstage_out = ns;
    return fixed + (size_t)ns * STAGE_BYTES;
}
// bulk mode: operand ring (na slots of 8 KB) + staging ring (nraw slots of 16 KB, +8 KB each when gated); 0 if it cannot fit
static inline size_t gemm_smem_bytes_bulk(int NOUT, int K, bool gated, int* nraw_out, int* na_out) {
    const int Kp = (K + 15) & ~15, nkb = (Kp + KBLK - 1) / KBLK;
    const size_t fixed = gemm_smem_fixed(NOUT, K, true);
    if (nkb > MAXKB || fixed >= SMEM_CAP) return 0;
    const size_t avail = SMEM_CAP - fixed, slot = (size_t)RAW_SLOT + (gated ? GRAW_SLOT : 0);
    int best_raw = 0, best_na = 0;
    for (int na = 2; na <= MAX_ASLOT; ++na) {
        if (avail < (size_t)na * A_SLOT_BULK + 2 * slot) break;
        int nraw = (int)((avail - (size_t)na * A_SLOT_BULK) / slot);
        if (nraw > MAX_RAW) nraw = MAX_RAW;
        // The staging depth must be a multiple of the number of loader warps / converter groups (2): unit u and unit
        // u + nraw then belong to the same loader and the same group, so every slot has ONE producer and ONE consumer
        // and a parity wait can never be two phases behind.  With an odd depth a loader can run ahead of the other
        // group's consumption and pass a wait on the aliased parity (found as a hang; tools/sim_pipeline.py).
        nraw &= ~1;
        if (nraw >= best_raw && nraw >= 2) { best_raw = nraw; best_na = na; }   // deepest staging first, then operand slots
    }
    if (best_raw >= 2) {
        *nraw_out = best_raw; *na_out = best_na;
        return fixed + (size_t)best_na * A_SLOT_BULK + (size_t)best_raw * slot;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- weight images
// The resident B operand of one GEMM: bf16, K-major rows of 128 bytes (64 channels), SWIZZLE_128B, K blocks of
// NOUT*128 bytes -- byte for byte what the kernel prologue would build in shared memory.  One launch converts the
// weights of every GEMM of a time step (they are re-read by all 148 CTAs of ~20 launches).
struct WImgSpec {
    const float* W; long w_ld, w_ks; int nrow1;
    const float* W2; long w2_ld; int k2;
    int NOUT, K, nout_store;
    unsigned long long off;                  // byte offset of the image in the arena
};
constexpr int WIMG_MAX = 28;
struct WImgBatch { WImgSpec s[WIMG_MAX]; int n; char* base; };

__global__ void __launch_bounds__(256) wimg_kernel(const WImgBatch B) {
    const WImgSpec& S = B.s[blockIdx.y];
    const int K = S.K, Kp = (K + 15) & ~15, nkb = (Kp + KBLK - 1) / KBLK;
    const int chunks_per_row = nkb * 8, total = S.NOUT * chunks_per_row, wblk_bytes = S.NOUT * 128;
    char* img = B.base + S.off;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int n = idx / chunks_per_row, ch = idx % chunks_per_row;
        const int kb = ch >> 3, j = ch & 7, k0 = kb * KBLK + j * 8;
        const bool second = n >= S.nrow1;
        const float* wrow = second ? (S.W2 + (long)(n - S.nrow1) * S.w2_ld) : (S.W + (long)n * S.w_ld);
        const int klim = second ? (S.k2 < K ? S.k2 : K) : K;
        const long ks = second ? 1 : S.w_ks;
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (k0 + u < klim && n < S.nout_store) ? __ldg(wrow + (long)(k0 + u) * ks) : 0.f;
        const uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        *reinterpret_cast<uint4*>(img + (size_t)kb * wblk_bytes + n * 128 + ((j ^ (n & 7)) << 4)) = pk;
    }
}
static inline size_t wimg_bytes(int NOUT, int K) {
    const int Kp = (K + 15) & ~15, nkb = (Kp + KBLK - 1) / KBLK;
    return (size_t)nkb * NOUT * 128;
}

// Host side: choose the producer mode and ring depths for a filled-in GemmParams; returns the dynamic shared memory
// size (0: the weights do not fit).  Bulk mode needs 16-byte aligned channel rows: fp32 maps with N % 4 == 0, bf16 maps
// whose planes are padded to whole tiles (so a full 256-byte row is always in bounds).
static inline size_t plan_launch(GemmParams& P, int epi, bool allow_bulk, bool allow_bias_mma = false) {
    const bool gated = P.seg.gate_seg >= 0;
    bool bulk = allow_bulk;
    const long npad = ((long)P.N + TILE_M - 1) / TILE_M * TILE_M;
    if (epi == EPI_POOL) {          // 8-byte gather copies of one fp32 map
        bulk = bulk && P.seg.kind[0] == 0 && P.seg.cend[0] == P.K && (reinterpret_cast<uintptr_t>(P.seg.src[0]) & 7) == 0 &&
               (P.seg.plane[0] * 4) % 8 == 0 && (P.img_w & 1) == 0;
    }
    for (int i = 0; i < 3 && bulk && epi != EPI_POOL; ++i) {
        const int width = P.seg.cend[i] - (i ? P.seg.cend[i - 1] : 0);
        if (width <= 0) continue;
        const size_t esz = P.seg.kind[i] ? 2 : 4;
        if (P.seg.kind[i])      // bf16 maps: 16-byte rows in planes padded to whole tiles
            bulk = bulk && (reinterpret_cast<uintptr_t>(P.seg.src[i]) & 15) == 0 && ((size_t)P.seg.plane[i] * esz) % 16 == 0 && P.seg.plane[i] >= npad;
        else                    // fp32 maps: 16-byte rows need N % 4 == 0 for the tail copy; otherwise 4-byte copies (any N)
            bulk = bulk && (reinterpret_cast<uintptr_t>(P.seg.src[i]) & 3) == 0;
    }
    if (bulk && gated)
        bulk = (reinterpret_cast<uintptr_t>(P.seg.gate_pre) & 15) == 0 && (P.seg.gate_plane * 2) % 16 == 0 && P.seg.gate_plane >= npad;
    int nstage = 0, nraw = 0, na = 0;
    size_t smem = 0;
    // out_vec decides whether the staged epilogue (the one that can drop its bias adds) is used at all
    P.out_vec = (P.out != nullptr && P.out_plane >= npad && P.out_plane % 8 == 0 && (reinterpret_cast<uintptr_t>(P.out) & 15) == 0 &&
                 (P.addend == nullptr || (reinterpret_cast<uintptr_t>(P.addend) & 15) == 0) && (epi == EPI_GN || epi == EPI_LRELU)) ? 1 : 0;
    P.bias_mma = 0;
    // (off by default: the hi + lo bf16 split changes results in the last bit relative to the epilogue add, and it only
    //  pays when the two channels fit into existing padding; URNN_BIAS_MMA=1 enables it)
    if (allow_bias_mma && bulk && P.out_vec && P.out_f32 == nullptr && P.bias != nullptr) {
        // only when the two extra channels fit into the padding of the last 32-channel unit (no extra unit per tile)
        const int nu0 = (((P.K + 15) & ~15) + UNIT_K - 1) / UNIT_K, nu1 = (((P.K + 2 + 15) & ~15) + UNIT_K - 1) / UNIT_K;
        if (nu1 == nu0) smem = gemm_smem_bytes_bulk(P.NOUT, P.K + 2, gated, &nraw, &na);
        if (smem != 0 && nraw >= 4) P.bias_mma = 1; else smem = 0;
    }
    if (bulk && !P.bias_mma) { smem = gemm_smem_bytes_bulk(P.NOUT, P.K, gated, &nraw, &na); if (smem == 0) bulk = false; }
    if (!bulk) smem = gemm_smem_bytes(P.NOUT, P.K, &nstage);
    P.nstage = nstage; P.bulk = bulk ? 1 : 0; P.nraw = nraw; P.na = na;
    int cols = 32;
    while (cols < 2 * P.NOUT) cols <<= 1;
    P.tmem_cols = cols;
    return smem;
}

}  // namespace tc
}  // namespace urnn
