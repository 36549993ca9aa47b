// pattern_bw.cu -- how fast can ANY kernel read K fp32 planes in 128-pixel tiles (512 B per channel per tile) and
// write NOUT bf16 planes, i.e. the access pattern of one sweep?  Sets the practical roofline for the planar NCHW layout.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// one CTA per tile (grid-stride); thread -> (channel row, 8-pixel chunk); 64 KB+ in flight per SM through occupancy
__global__ void __launch_bounds__(256) tile_stream(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int K, int NOUT,
                                                   long N, int ntiles, float* sink) {
    float acc = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long p0 = (long)tile * 128;
        for (int i = threadIdx.x; i < K * 16; i += 256) {
            int c = i >> 4, q = i & 15;
            const float4* s = reinterpret_cast<const float4*>(in + (long)c * N + p0 + q * 8);
            float4 a = __ldg(s), b = __ldg(s + 1);
            acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
        }
        for (int i = threadIdx.x; i < NOUT * 16; i += 256) {
            int c = i >> 4, q = i & 15;
            uint4 v = make_uint4(__float_as_uint(acc), 1u, 2u, 3u);
            *reinterpret_cast<uint4*>(out + (long)c * N + p0 + q * 8) = v;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    const long N = 250000; const int K = 224, NOUT = 128;
    float* in; __nv_bfloat16* out; float* sink;
    CK(cudaMalloc(&in, (size_t)K * N * 4)); CK(cudaMalloc(&out, (size_t)NOUT * N * 2)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 0, (size_t)K * N * 4));
    int ntiles = (int)((N + 127) / 128);
    for (int cfg = 0; cfg < 3; ++cfg) {
        int grid = cfg == 0 ? 148 * 8 : (cfg == 1 ? 148 * 4 : ntiles);
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        tile_stream<<<grid, 256>>>(in, out, K, NOUT, N, ntiles, sink);
        CK(cudaEventRecord(a));
        for (int i = 0; i < 10; ++i) tile_stream<<<grid, 256>>>(in, out, K, NOUT, N, ntiles, sink);
        CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        double us = ms * 100.0, bytes = (double)K * N * 4 + (double)NOUT * N * 2;
        printf("grid %5d: %.1f us, %.0f GB/s (224 fp32 planes in, 128 bf16 planes out, 128-pixel tiles)\n", grid, us, bytes / us * 1e-3);
    }
    return 0;
}
