// ptx_sm100.cuh -- inline-PTX wrappers for sm_100a: mbarrier, TMA tensor loads, tcgen05 (MMA / TMEM), descriptors.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace urnn {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp; the compiler keeps warp-uniform operands in uniform registers around it, which a
// plain `if (lane == 0)` does not allow (it emits an ELECT / R2UR.BROADCAST waterfall in front of every UTCHMMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// parks in hardware until the phase with the given parity completes (suspend-time hint; see profiles/ round 1: a bare
// try_wait loop was 43-53 % of the issued instructions)
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA: 3-D tiled tensor load global -> shared, completion on an mbarrier (complete_tx::bytes)
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint) : "memory");
}
// L2 prefetch of the same box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ---- tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; 16-bit operands -> fp32
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// the mbarrier receives one arrival once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 128 B, 8-row atoms 1024 B apart): weights
__host__ __device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);       // start address
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// MN-major (pixel-contiguous) A tile, 128-byte swizzle: 1024-byte atoms of 8 K-rows x 64 MN elements;
// leading byte offset = distance between atoms along MN, stride byte offset = along K (1024).
__host__ __device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D fp32, A MN-major, B K-major, M=128, N=nout; both operands fp16 (format 0) or bf16 (1)
__device__ __forceinline__ uint32_t instr_desc_16(int nout, uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | ((uint32_t)(nout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- hi/lo split of an fp32 value into two 16-bit floats: v = hi + lo.
// fp16 pairs (default): 11 + 11 significant bits -- fp32-grade products with three MMAs (the dropped lo*lo term is 2^-24);
// values beyond +-65504 saturate.  bf16 pairs (URNN_SPLIT_BF16): 8 + 8 bits, fp32 range, error 2^-18 per operand --
// measured to leave the config-3 max |d state| budget at T = 180 on the full grid (0.35 vs 0.1).
#ifdef URNN_SPLIT_BF16
typedef __nv_bfloat16 sp16;
constexpr uint32_t SPLIT_FMT = 1;
__device__ __forceinline__ float lo16_to_f32(uint32_t w) { return __uint_as_float(w << 16); }              // low 16-bit element of a pair
__device__ __forceinline__ float hi16_to_f32(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }      // high 16-bit element
__device__ __forceinline__ uint32_t pack2_16(float a, float b) { const __nv_bfloat162 t = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&t); }
#else
typedef __half sp16;
constexpr uint32_t SPLIT_FMT = 0;
__device__ __forceinline__ float lo16_to_f32(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
__device__ __forceinline__ float hi16_to_f32(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
__device__ __forceinline__ uint32_t pack2_16(float a, float b) {
    const __half2 t = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
    return *reinterpret_cast<const uint32_t*>(&t);
}
#endif
// split one value: low half of the result = hi part, high half = lo part
__device__ __forceinline__ uint32_t split16(float v) {
    const uint32_t h = pack2_16(v, 0.f) & 0xFFFFu;
    const uint32_t l = pack2_16(v - lo16_to_f32(h), 0.f) & 0xFFFFu;
    return h | (l << 16);
}
// split a pixel pair: (hi pair, lo pair), each packed (first pixel in the low half)
__device__ __forceinline__ void split16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack2_16(a, b);
    lo = pack2_16(a - lo16_to_f32(hi), b - hi16_to_f32(hi));
}

}  // namespace ptx
}  // namespace urnn
