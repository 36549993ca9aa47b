// metrics.cu -- device-side post-processing of an inference run: the reference de-normalises the
// (T, H, W) prediction on the host (test.py:468, Dynamic2DFlood.py:379-385 r_MinMaxScaler) and computes R2 / MSE / RMSE /
// MAE / PeakR2 / CSI with numpy (test.py:607-675 compute_metrics).  Here the same quantities are streaming reductions over
// chunks of time steps that are still in HBM: per-element arithmetic in fp32 exactly as numpy does it on float32 arrays
// (pred_mm = p * flood_max; pred_m = pred_mm / 1000; d = pred_m - gt_m), sums in double, temporal maxima kept as fp32 maps,
// wet / dry counts as integers (bit-exact).  One pass over the data: 8 bytes of HBM traffic per (cell, step).
//
// Workspace layout (doubles first):
//   acc[8]        0: sum d^2   1: sum |d|   2: sum g   3: sum g^2   4: elements seen   5..7: tp, fp, fn (finalize)
//   step[T][4]    per time step: sum d^2, sum g, sum g^2, (unused)
//   pred_max[HW], gt_max[HW]   temporal maxima in mm (fp32)
#include <cuda_runtime.h>
#include <stdint.h>
#include "urnn_common.cuh"
#include "urnn_internal.h"

namespace urnn {
namespace metrics {

constexpr int BLOCK = 256;
constexpr int MAX_CHUNK = 64;          // time steps per accumulate launch (per-block shared partials)

struct Ws { double* acc; double* step; float* pred_max; float* gt_max; size_t total; };

static Ws carve(int H, int W, int T, void* ws) {
    Ws w; char* b = (char*)ws; size_t o = 0;
    w.acc = (double*)(b + o); o += 8 * sizeof(double);
    w.step = (double*)(b + o); o += (size_t)T * 4 * sizeof(double);
    o = (o + 255) / 256 * 256;
    w.pred_max = (float*)(b + o); o += (size_t)H * W * sizeof(float);
    o = (o + 255) / 256 * 256;
    w.gt_max = (float*)(b + o); o += (size_t)H * W * sizeof(float);
    w.total = (o + 255) / 256 * 256;
    return w;
}

__global__ void __launch_bounds__(BLOCK) reset_kernel(double* acc, double* step, int T, float* pred_max, float* gt_max, long long hw) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x, n = (long long)gridDim.x * BLOCK;
    for (long long k = i; k < hw; k += n) { pred_max[k] = -INFINITY; gt_max[k] = -INFINITY; }
    for (long long k = i; k < 4LL * T; k += n) step[k] = 0.0;
    if (i < 8) acc[i] = 0.0;
}

// thread = pixel, loop over the chunk's time steps (coalesced across the warp at every step)
__global__ void __launch_bounds__(BLOCK) accumulate_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long hw,
                                                           int t0, int nsteps, float scale, double* acc, double* step,
                                                           float* pred_max, float* gt_max) {
    __shared__ double sstep[MAX_CHUNK][3];
    __shared__ double sacc[4];
    for (int i = threadIdx.x; i < nsteps * 3; i += BLOCK) sstep[i / 3][i % 3] = 0.0;
    if (threadIdx.x < 4) sacc[threadIdx.x] = 0.0;
    __syncthreads();
    const long long p = (long long)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = p < hw;
    const int lane = threadIdx.x & 31;
    float pm = -INFINITY, gm = -INFINITY;
    if (valid) { pm = pred_max[p]; gm = gt_max[p]; }
    double a_d2 = 0.0, a_ad = 0.0, a_g = 0.0, a_g2 = 0.0;
    for (int t = 0; t < nsteps; ++t) {
        float d2 = 0.f, g = 0.f, g2 = 0.f;
        if (valid) {
            const float pmm = __ldg(pred + (long long)t * hw + p) * scale;      // r_MinMaxScaler(min = 0): data * (max - 0) + 0
            const float gmm = __ldg(gt + (long long)t * hw + p);
            pm = fmaxf(pm, pmm); gm = fmaxf(gm, gmm);
            const float pmet = __fdiv_rn(pmm, 1000.0f), gmet = __fdiv_rn(gmm, 1000.0f);
            const float d = pmet - gmet;
            d2 = d * d; g = gmet; g2 = gmet * gmet;
            a_d2 += (double)d2; a_ad += (double)fabsf(d); a_g += (double)g; a_g2 += (double)g2;
        }
        // per-step sums: 32 values in fp32 (relative error 3e-7), then double
        const float w_d2 = warp_sum(d2), w_g = warp_sum(g), w_g2 = warp_sum(g2);
        if (lane == 0) { atomicAdd(&sstep[t][0], (double)w_d2); atomicAdd(&sstep[t][1], (double)w_g); atomicAdd(&sstep[t][2], (double)w_g2); }
    }
    if (valid) { pred_max[p] = pm; gt_max[p] = gm; }
    a_d2 = warp_sum(a_d2); a_ad = warp_sum(a_ad); a_g = warp_sum(a_g); a_g2 = warp_sum(a_g2);
    if (lane == 0) { atomicAdd(&sacc[0], a_d2); atomicAdd(&sacc[1], a_ad); atomicAdd(&sacc[2], a_g); atomicAdd(&sacc[3], a_g2); }
    __syncthreads();
    for (int i = threadIdx.x; i < nsteps * 3; i += BLOCK) atomicAdd(&step[(size_t)(t0 + i / 3) * 4 + i % 3], sstep[i / 3][i % 3]);
    if (threadIdx.x < 4) atomicAdd(&acc[threadIdx.x], sacc[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long nvalid = hw;
        atomicAdd(&acc[4], (double)nvalid * nsteps);
    }
}

// wet / dry agreement of the temporal-maximum maps (test.py:654-663): integer counts
__global__ void __launch_bounds__(BLOCK) csi_kernel(const float* __restrict__ pred_max, const float* __restrict__ gt_max, long long hw,
                                                    float thres, unsigned long long* counts) {
    unsigned tp = 0, fp = 0, fn = 0;
    for (long long p = (long long)blockIdx.x * BLOCK + threadIdx.x; p < hw; p += (long long)gridDim.x * BLOCK) {
        const bool a = pred_max[p] > thres, b = gt_max[p] > thres;
        tp += a && b; fp += a && !b; fn += !a && b;
    }
    tp = __reduce_add_sync(0xffffffffu, tp); fp = __reduce_add_sync(0xffffffffu, fp); fn = __reduce_add_sync(0xffffffffu, fn);
    if ((threadIdx.x & 31) == 0) {
        if (tp) atomicAdd(&counts[0], (unsigned long long)tp);
        if (fp) atomicAdd(&counts[1], (unsigned long long)fp);
        if (fn) atomicAdd(&counts[2], (unsigned long long)fn);
    }
}

// out[12]: R2, MSE, RMSE, MAE, PeakR2, CSI, tp, fp, fn, t_peak, elements, (unused)
__global__ void finalize_kernel(const double* acc, const double* step, int T, long long hw, const unsigned long long* counts, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double n = acc[4];
    const double ss_res = acc[0], sg = acc[2], sg2 = acc[3];
    const double mean = n > 0 ? sg / n : 0.0;
    double ss_tot = sg2 - sg * mean; if (ss_tot < 0.0) ss_tot = 0.0;
    out[0] = 1.0 - ss_res / (ss_tot + 1e-10);
    const double mse = n > 0 ? ss_res / n : 0.0;
    out[1] = mse; out[2] = sqrt(mse); out[3] = n > 0 ? acc[1] / n : 0.0;
    int tp_ = 0; double best = -INFINITY;
    for (int t = 0; t < T; ++t) { const double m = step[(size_t)t * 4 + 1] / (double)hw; if (m > best) { best = m; tp_ = t; } }   // np.argmax: first maximum
    const double r = step[(size_t)tp_ * 4], g1 = step[(size_t)tp_ * 4 + 1], g2 = step[(size_t)tp_ * 4 + 2];
    double tot_p = g2 - g1 * (g1 / (double)hw); if (tot_p < 0.0) tot_p = 0.0;
    out[4] = 1.0 - r / (tot_p + 1e-10);
    const double tp = (double)counts[0], fp = (double)counts[1], fn = (double)counts[2];
    out[5] = tp / (tp + fp + fn + 1e-10);
    out[6] = tp; out[7] = fp; out[8] = fn; out[9] = (double)tp_; out[10] = n; out[11] = 0.0;
}

}  // namespace metrics
}  // namespace urnn

using namespace urnn;
using namespace urnn::metrics;

extern "C" {

size_t urnn_metrics_workspace_bytes(int32_t H, int32_t W, int32_t T) {
    if (H <= 0 || W <= 0 || T <= 0) return 0;
    return carve(H, W, T, nullptr).total;
}

static int check(int32_t H, int32_t W, int32_t T, const void* ws, size_t ws_bytes) {
    URNN_CHECK_ARG(H > 0 && W > 0 && T > 0, "metrics: H, W, T must be positive");
    URNN_CHECK_ARG(ws != nullptr, "metrics: workspace is NULL");
    if (ws_bytes < carve(H, W, T, nullptr).total) { set_error("metrics: workspace %zu < %zu bytes", ws_bytes, carve(H, W, T, nullptr).total); return URNN_E_WORKSPACE; }
    return URNN_OK;
}

int urnn_metrics_reset(int32_t H, int32_t W, int32_t T, void* ws, size_t ws_bytes, void* stream) {
    URNN_TRY(check(H, W, T, ws, ws_bytes));
    const Ws w = carve(H, W, T, ws);
    const long long hw = (long long)H * W;
    const int grid = (int)((hw + BLOCK - 1) / BLOCK < 4096 ? (hw + BLOCK - 1) / BLOCK : 4096);
    reset_kernel<<<grid, BLOCK, 0, (cudaStream_t)stream>>>(w.acc, w.step, T, w.pred_max, w.gt_max, hw);
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}

int urnn_metrics_accumulate(int32_t H, int32_t W, int32_t T, int32_t t0, int32_t nsteps, const float* pred_norm_dev,
                            const float* gt_mm_dev, float flood_max, void* ws, size_t ws_bytes, void* stream) {
    URNN_TRY(check(H, W, T, ws, ws_bytes));
    URNN_CHECK_ARG(pred_norm_dev && gt_mm_dev, "metrics: prediction / ground-truth pointer is NULL");
    URNN_CHECK_ARG(t0 >= 0 && nsteps >= 0 && t0 + nsteps <= T, "metrics: steps [%d, %d) outside [0, %d)", t0, t0 + nsteps, T);
    const Ws w = carve(H, W, T, ws);
    const long long hw = (long long)H * W;
    for (int s = 0; s < nsteps; s += MAX_CHUNK) {
        const int ns = nsteps - s < MAX_CHUNK ? nsteps - s : MAX_CHUNK;
        accumulate_kernel<<<(unsigned)((hw + BLOCK - 1) / BLOCK), BLOCK, 0, (cudaStream_t)stream>>>(
            pred_norm_dev + (long long)s * hw, gt_mm_dev + (long long)s * hw, hw, t0 + s, ns, flood_max, w.acc, w.step, w.pred_max, w.gt_max);
        URNN_LAUNCH_CHECK();
    }
    return URNN_OK;
}

int urnn_metrics_finalize(int32_t H, int32_t W, int32_t T, float flood_thres, void* ws, size_t ws_bytes, double* out12_dev, void* stream) {
    URNN_TRY(check(H, W, T, ws, ws_bytes));
    URNN_CHECK_ARG(out12_dev != nullptr, "metrics: output pointer is NULL");
    const Ws w = carve(H, W, T, ws);
    const long long hw = (long long)H * W;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(w.acc + 5);     // tp, fp, fn as 64-bit integers
    URNN_CUDA(cudaMemsetAsync(counts, 0, 3 * sizeof(unsigned long long), st));
    const int grid = (int)((hw + BLOCK - 1) / BLOCK < 2048 ? (hw + BLOCK - 1) / BLOCK : 2048);
    csi_kernel<<<grid, BLOCK, 0, st>>>(w.pred_max, w.gt_max, hw, flood_thres, counts);
    URNN_LAUNCH_CHECK();
    finalize_kernel<<<1, 32, 0, st>>>(w.acc, w.step, T, hw, counts, out12_dev);
    URNN_LAUNCH_CHECK();
    return URNN_OK;
}

}  // extern "C"
