// tma_bw.cu -- what TMA tensor-load geometry streams HBM fastest on B200?  One persistent CTA per SM, one producer thread,
// DEPTH boxes in flight, consumer = the same thread (waits for the oldest box, re-issues).  Reports GB/s per geometry.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/tma_bw tools/tma_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (done) break;
    }
}
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
    }
}
__device__ int g_spin;
__device__ __forceinline__ void tma4(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// boxes are enumerated as (i0, i1, i2, i3) steps of the box size over the tensor; box index b -> coordinates
struct Geo { int nb[4]; int box[4]; int box_bytes; long long nboxes; int spin; };
static int g_host_spin = 0;

template <int DEPTH, int NPROD>
__global__ void __launch_bounds__(128, 1) tma_stream(const __grid_constant__ CUtensorMap tm, const Geo g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[DEPTH * NPROD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < DEPTH * NPROD; ++i) mbar_init(smem_u32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (warp < NPROD && lane == 0) {
        // producer `warp` handles boxes b = blockIdx.x + (NPROD*k + warp) * gridDim.x
        long long b = blockIdx.x + (long long)warp * gridDim.x;
        const long long stride = (long long)NPROD * gridDim.x;
        int issued = 0, done = 0;
        auto issue = [&](long long bi, int slot) {
            long long r = bi;
            const int i0 = (int)(r % g.nb[0]); r /= g.nb[0];
            const int i1 = (int)(r % g.nb[1]); r /= g.nb[1];
            const int i2 = (int)(r % g.nb[2]); r /= g.nb[2];
            const int i3 = (int)r;
            const uint32_t bar = smem_u32(&bars[warp * DEPTH + slot]);
            mbar_expect(bar, (uint32_t)g.box_bytes);
            tma4(base + (uint32_t)(warp * DEPTH + slot) * (uint32_t)g.box_bytes, &tm, bar, i0 * g.box[0], i1 * g.box[1], i2 * g.box[2], i3 * g.box[3]);
        };
        for (; issued < DEPTH && b < g.nboxes; ++issued, b += stride) issue(b, issued);
        const int total_first = issued;
        while (done < issued) {
            const int slot = done % DEPTH;
            if (g.spin) mbar_spin(smem_u32(&bars[warp * DEPTH + slot]), (uint32_t)((done / DEPTH) & 1));
            else mbar_wait(smem_u32(&bars[warp * DEPTH + slot]), (uint32_t)((done / DEPTH) & 1));
            ++done;
            if (b < g.nboxes) { issue(b, slot); ++issued; b += stride; }
        }
        (void)total_first;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int DEPTH, int NPROD>
static void run(const char* name, EncodeTiledFn enc, void* buf, const long long dims_[4], const long long strides_b[3], const int box_[4], int sms) {
    cuuint64_t dims[4]; cuuint64_t strides[3]; cuuint32_t box[4]; cuuint32_t es[4] = {1, 1, 1, 1};
    Geo g; g.nboxes = 1; g.box_bytes = 2; g.spin = g_host_spin;
    for (int i = 0; i < 4; ++i) { dims[i] = dims_[i]; box[i] = box_[i]; g.box[i] = box_[i]; g.nb[i] = (int)(dims_[i] / box_[i]); g.nboxes *= g.nb[i]; g.box_bytes *= box_[i]; }
    for (int i = 0; i < 3; ++i) strides[i] = strides_b[i];
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-58s encode failed %d\n", name, (int)r); return; }
    const size_t smem = (size_t)DEPTH * NPROD * g.box_bytes + 2048;
    CK(cudaFuncSetAttribute(tma_stream<DEPTH, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    tma_stream<DEPTH, NPROD><<<sms, 128, smem>>>(tm, g);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; ++i) tma_stream<DEPTH, NPROD><<<sms, 128, smem>>>(tm, g);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)g.nboxes * g.box_bytes;
    printf("%-58s depth %d prod %d box %5d B: %7.1f us  %6.0f GB/s  (%.1f GB/s per SM)\n", name, DEPTH, NPROD, g.box_bytes, ms / 3 * 1e3, bytes / (ms / 3 * 1e-3) * 1e-9,
           bytes / (ms / 3 * 1e-3) * 1e-9 / sms);
}

int main() {
    int sms; CK(cudaSetDevice(0)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const long long ntot = 251904, C = 128;              // 128 channels x hi/lo x 251904 px bf16 = 129 MB
    void* buf; CK(cudaMalloc(&buf, (size_t)ntot * C * 2 * 2)); CK(cudaMemset(buf, 1, (size_t)ntot * C * 2 * 2));
    const long long ntile = ntot / 128;
    {   // plane-major [2][C][ntot]: what gemm_v2 uses today
        const long long d[4] = {ntot, C, 2, 1}, s[3] = {ntot * 2, C * ntot * 2, 2 * C * ntot * 2};
        const int b1[4] = {64, 32, 2, 1}; run<8, 1>("plane-major, box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
        const int b2[4] = {64, 32, 1, 1}; run<8, 1>("plane-major, box 64px x 32ch", enc, buf, d, s, b2, sms);
        const int b3[4] = {64, 128, 1, 1}; run<8, 1>("plane-major, box 64px x 128ch", enc, buf, d, s, b3, sms);
        run<8, 2>("plane-major, box 64px x 32ch x hi/lo, 2 producers", enc, buf, d, s, b1, sms);
        run<4, 4>("plane-major, box 64px x 32ch x hi/lo, 4 producers", enc, buf, d, s, b1, sms);
    }
    {
        const long long d[4] = {ntot, C, 2, 1}, s[3] = {ntot * 2, C * ntot * 2, 2 * C * ntot * 2};
        const int b1[4] = {64, 32, 2, 1};
        g_host_spin = 1;
        run<8, 1>("SPIN plane-major, box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
        run<16, 1>("SPIN plane-major, box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
        run<4, 4>("SPIN plane-major, box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
        g_host_spin = 0;
        run<16, 1>("plane-major, box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
        run<2, 1>("plane-major, box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
        run<2, 4>("plane-major, box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
        run<6, 4>("plane-major, box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
    }
    if (0) {   // tile-major [tile][2][C][128 px]: rows of one tile are contiguous (256 B apart)
        const long long d[4] = {128, C, 2, ntile}, s[3] = {256, C * 256, 2 * C * 256};
        const int b1[4] = {64, 32, 2, 1}; run<8, 1>("tile-major, box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
        const int b2[4] = {64, 128, 1, 1}; run<8, 1>("tile-major, box 64px x 128ch", enc, buf, d, s, b2, sms);
        run<8, 2>("tile-major, box 64px x 32ch x hi/lo, 2 producers", enc, buf, d, s, b1, sms);
        run<4, 4>("tile-major, box 64px x 32ch x hi/lo, 4 producers", enc, buf, d, s, b1, sms);
    }
    if (0) {   // half-tile-major [halftile][2][C][64 px]: a box is one fully contiguous block
        const long long d[4] = {64, C, 2, 2 * ntile}, s[3] = {128, C * 128, 2 * C * 128};
        const int b1[4] = {64, 32, 2, 1}; run<8, 1>("half-tile-major (dense), box 64px x 32ch x hi/lo", enc, buf, d, s, b1, sms);
        const int b2[4] = {64, 128, 1, 1}; run<8, 1>("half-tile-major (dense), box 64px x 128ch", enc, buf, d, s, b2, sms);
        const int b3[4] = {64, 128, 2, 1}; run<4, 1>("half-tile-major (dense), box 64px x 128ch x hi/lo", enc, buf, d, s, b3, sms);
        run<4, 4>("half-tile-major (dense), box 64px x 32ch x hi/lo, 4 prod", enc, buf, d, s, b1, sms);
    }
    return 0;
}
