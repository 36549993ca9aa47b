/*
 * urnn_b200.h -- C ABI of liburnn_b200.so: the U-RNN ConvGRU encoder-decoder time step on B200 (sm_100a).
 *
 * The reference (holmescao/U-RNN) is pure Python/PyTorch and has no FFI of its own; the seam it offers
 * is the nn.Module plug-in point (net_params.py:90-100,127-137 constructs the CGRU_cell objects that
 * model.py:52-63 injects into ED).  Each entry point below replaces the ATen op sequence of one reference
 * forward/backward and cites it.  Paths are relative to /root/reference/code/src/lib/model/networks/.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch types.  All tensor pointers are DEVICE pointers to
 *     fp32, NCHW-contiguous, batch 1 (the reference's fixed B=1,S=1 contract) unless a name
 *     ends in _host.  Weight pointers use the reference's own state_dict tensor layouts unchanged
 *     (nn.Conv2d (Cout,Cin,k,k); nn.ConvTranspose2d (Cin,Cout,2,2); LayerNorm (16,H,W)).
 *   - Every call is stream-ordered on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     allocates nothing and is CUDA-graph capturable.  The caller owns every buffer including the
 *     workspace: query urnn_*_workspace_bytes first.  Workspace contents need not be preserved or zeroed.
 *   - Return value: 0 on success, a negative URNN_E_* code otherwise; urnn_last_error() returns a
 *     thread-local message for the last failing call on this thread.
 *   - Inputs are never modified.  Outputs are fully overwritten.
 *   - Devices / threads: every call acts on the CUDA device that is current on the calling thread; nothing is cached per
 *     process (function attributes are set per launch, copy streams of the *_host entry points exist once per device,
 *     launch plans of urnn_ed_step_fwd are cached per (device, descriptor, parameter block, workspace, communicator)).
 *     The spatial-sharding communicator (urnn_comm_*) is the one process-wide object: one process drives one GPU.
 */
#ifndef URNN_B200_H
#define URNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define URNN_ABI_VERSION 1

/* error codes */
#define URNN_OK            0
#define URNN_E_INVALID    -1   /* bad argument / unsupported shape */
#define URNN_E_WORKSPACE  -2   /* workspace too small */
#define URNN_E_CUDA       -3   /* a CUDA runtime call failed */
#define URNN_E_UNSUPPORTED -4  /* valid request that this build does not implement */

/* cell variants (ConvRNN.py:82-89) */
#define URNN_CELL_ENCODER 0    /* hidden = h (F ch)          : gates from cat(x,h)     */
#define URNN_CELL_DECODER 1    /* hidden = cat(e,d) (2F ch)  : gates from cat(x,e,d)   */

/* arithmetic of the gate contractions */
#define URNN_MATH_FP32 0       /* fp32 FFMA, fp32 accumulate: the parity mode (atol 1e-5 vs reference)                */
#define URNN_MATH_F16X3 3     /* tcgen05 kind::f16 with every GEMM operand split into fp16 hi + lo (3 MMAs per K step: 22      */
                               /* significant bits on both sides), fp32 accumulate / statistics / pre-norm maps; TMA-fed.     */
                               /* Whole-step entry points only (ksize 1).  The mode that meets the T=180 tolerance.          */
#define URNN_MATH_BF16 2       /* tcgen05 kind::f16: GEMM operands rounded to bf16, fp32 accumulate in TMEM, fp32     */
                               /* statistics / states; applies to the cell contractions and the stage stems (k=1)     */

int         urnn_abi_version(void);
const char* urnn_last_error(void);
/* number of kernels launched by this library in this process since load (bench.py's gpu_launches) */
uint64_t    urnn_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * (Skip-)ConvGRU cell, one time step.  Replaces CGRU_cell.forward, ConvRNN.py:111-194:
 *   g = GN(conv1(cat(x,[e,]h)));  z = sigmoid(g[:F]);  r = sigmoid(g[F:]);
 *   c = tanh(GN(conv2(cat(x,[e,] r*h))));  h_out = (1-z)*h + z*c
 * GroupNorm: 32 channels per group, biased variance over the whole H x W grid, eps (ConvRNN.py:97,103).
 * ------------------------------------------------------------------------------------------------ */
typedef struct urnn_cell_desc {
    int32_t H, W;        /* spatial size at this scale                                             */
    int32_t Cx;          /* channels of x (weights always carry these columns, even if x == NULL)  */
    int32_t F;           /* num_features; multiple of 32 (GroupNorm(F//32), ConvRNN.py:103)        */
    int32_t ksize;       /* odd filter size; 1 is the production path (network.yaml:27,46)         */
    int32_t variant;     /* URNN_CELL_ENCODER / URNN_CELL_DECODER                                  */
    int32_t math;        /* URNN_MATH_*                                                            */
    float   eps;         /* GroupNorm eps (1e-5)                                                   */
} urnn_cell_desc;

typedef struct urnn_cell_params {   /* the cell's 8 tensors, ConvRNN.py:94-104 */
    const float* w1;     /* conv1.0.weight (2F, Cx+Ch, k, k), Ch = F (encoder) or 2F (decoder)     */
    const float* b1;     /* conv1.0.bias   (2F)                                                    */
    const float* gn1_w;  /* conv1.1.weight (2F)                                                    */
    const float* gn1_b;  /* conv1.1.bias   (2F)                                                    */
    const float* w2;     /* conv2.0.weight (F, Cx+Ch, k, k)                                        */
    const float* b2;     /* conv2.0.bias   (F)                                                     */
    const float* gn2_w;  /* conv2.1.weight (F)                                                     */
    const float* gn2_b;  /* conv2.1.bias   (F)                                                     */
} urnn_cell_params;

size_t urnn_cgru_fwd_workspace_bytes(const urnn_cell_desc* d);
/* x: (Cx,H,W) or NULL (= zeros, ConvRNN.py:143-146);  e: (F,H,W) encoder skip state (decoder only, else NULL);
 * h: (F,H,W) previous state (decoder: d_{t-1});  h_out: (F,H,W). */
int urnn_cgru_fwd(const urnn_cell_desc* d, const urnn_cell_params* p,
                  const float* x, const float* e, const float* h, float* h_out,
                  void* ws, size_t ws_bytes, void* stream);

/* Backward of the cell (autograd of ConvRNN.py:111-194; the reference's reentrant checkpointing means
 * recompute-in-backward, ConvRNN.py:154-158).  Inputs are the forward inputs plus dL/dh_out.  Gradient
 * outputs may be NULL to skip them; dx/de/dh are overwritten, the 8 parameter gradients are ACCUMULATED
 * (+=) into grads (same layout as urnn_cell_params, non-const) as autograd does for reused parameters. */
typedef struct urnn_cell_grads {
    float *w1, *b1, *gn1_w, *gn1_b, *w2, *b2, *gn2_w, *gn2_b;
} urnn_cell_grads;
size_t urnn_cgru_bwd_workspace_bytes(const urnn_cell_desc* d);
int urnn_cgru_bwd(const urnn_cell_desc* d, const urnn_cell_params* p,
                  const float* x, const float* e, const float* h, const float* dh_out,
                  float* dx, float* de, float* dh, const urnn_cell_grads* grads,
                  void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage stems (utils.py:73-125 make_layers; encoder.py:142-157; decoder.py:150-165).
 * ------------------------------------------------------------------------------------------------ */
/* y = [AvgPool2](LeakyReLU(conv1x1(x)+b, slope)).  pool = 1 (none) or 2 (AvgPool2d(2,2) AFTER the
 * activation, utils.py:92-94 layer order; needs H,W even).  x (Cin,H,W) -> y (Cout,H/pool,W/pool). */
int urnn_conv1x1_lrelu_fwd(int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t pool, float slope, int32_t math,
                           const float* x, const float* w, const float* b, float* y, void* stream);
/* dx overwritten (may be NULL); dw,db accumulated (may be NULL). */
size_t urnn_conv1x1_lrelu_bwd_workspace_bytes(int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t pool);
int urnn_conv1x1_lrelu_bwd(int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t pool, float slope,
                           const float* x, const float* w, const float* b, const float* dy,
                           float* dx, float* dw, float* db, void* ws, size_t ws_bytes, void* stream);

/* y = LeakyReLU(ConvTranspose2d(k=2,s=2,p=0)(x)+b): x (Cin,H,W), w (Cin,Cout,2,2) -> y (Cout,2H,2W). */
int urnn_deconv2x2_lrelu_fwd(int32_t Cin, int32_t Cout, int32_t H, int32_t W, float slope, int32_t math,
                             const float* x, const float* w, const float* b, float* y, void* stream);
size_t urnn_deconv2x2_lrelu_bwd_workspace_bytes(int32_t Cin, int32_t Cout, int32_t H, int32_t W);
int urnn_deconv2x2_lrelu_bwd(int32_t Cin, int32_t Cout, int32_t H, int32_t W, float slope,
                             const float* x, const float* w, const float* b, const float* dy,
                             float* dx, float* dw, float* db, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Head.  Replaces YOLOXHead.forward + correction_depth, head/flood_head.py:131-202, with width 16
 * (model.py:62-63): stem, 2 cls blocks, 2 reg blocks = conv1x1(16->16, no bias) -> LayerNorm([16,H,W]) -> SiLU
 * (head/network_blocks.py:74-101); predictions conv1x1(16->1)+bias -> sigmoid / LeakyReLU(0.2);
 * depth *= (prob >= cls_thred).
 * ------------------------------------------------------------------------------------------------ */
typedef struct urnn_head_params {
    const float* conv_w[5];   /* stems, cls_convs.0, cls_convs.1, reg_convs.0, reg_convs.1: conv.weight (16,16,1,1) */
    const float* ln_w[5];     /* matching ln.weight (16,H,W) */
    const float* ln_b[5];     /* matching ln.bias   (16,H,W) */
    const float* cls_pred_w;  /* cls_preds.conv.weight (1,16,1,1) */
    const float* cls_pred_b;  /* cls_preds.conv.bias   (1)        */
    const float* reg_pred_w;  /* reg_preds.conv.weight (1,16,1,1) */
    const float* reg_pred_b;  /* reg_preds.conv.bias   (1)        */
} urnn_head_params;
typedef struct urnn_head_grads {
    float* conv_w[5]; float* ln_w[5]; float* ln_b[5];
    float *cls_pred_w, *cls_pred_b, *reg_pred_w, *reg_pred_b;
} urnn_head_grads;

size_t urnn_head_fwd_workspace_bytes(int32_t H, int32_t W);
/* feat (16,H,W) -> out (2,H,W): out[0] = masked depth, out[1] = wet probability (flood_head.py:164-175). */
int urnn_head_fwd(int32_t H, int32_t W, float cls_thred, float ln_eps, float slope,
                  const urnn_head_params* p, const float* feat, float* out,
                  void* ws, size_t ws_bytes, void* stream);
/* dout (2,H,W) -> dfeat (16,H,W) overwritten; parameter gradients accumulated.  The mask is a constant
 * (flood_head.py:201 .float() of a comparison), so with dout[1] == 0 every cls-branch gradient is exactly 0. */
size_t urnn_head_bwd_workspace_bytes(int32_t H, int32_t W);
int urnn_head_bwd(int32_t H, int32_t W, float cls_thred, float ln_eps, float slope,
                  const urnn_head_params* p, const float* feat, const float* dout,
                  float* dfeat, const urnn_head_grads* grads,
                  void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole encoder-decoder time step.  Replaces ED.forward, model.py:65-121 (Encoder.forward encoder.py:187-215,
 * Decoder.forward decoder.py:173-217, head) for the published topology: downsample factors (1,2,2),
 * upsample (2,2,1) (network.yaml), H and W multiples of 4.
 * ------------------------------------------------------------------------------------------------ */
typedef struct urnn_ed_desc {
    int32_t H, W, Cin;        /* Cin = 2*historical_nums + 3 (utils/net_config.py:119-142) */
    int32_t enc_conv[3];      /* encoder.conv_out_channels  (16,64,96)  */
    int32_t enc_gru[3];       /* encoder.gru_channels       (64,96,96)  */
    int32_t dec_gru[3];       /* decoder.gru_channels, deepest first (96,96,64) */
    int32_t dec_conv[3];      /* decoder.conv_out_channels  (96,96,16); dec_conv[2] must be 16 */
    int32_t ksize;            /* must be 1 (ED does not work with k>1 in the reference either) */
    int32_t math;             /* URNN_MATH_* */
    float   cls_thred, gn_eps, ln_eps, lrelu_slope;
} urnn_ed_desc;

typedef struct urnn_ed_params {
    const float* enc_stem_w[3]; const float* enc_stem_b[3];  /* encoder.stage{1,2,3}.conv{1,2,3}_leaky_1 */
    urnn_cell_params enc_cell[3];                            /* encoder.rnn{1,2,3} */
    urnn_cell_params dec_cell[3];                            /* decoder.rnn{3,2,1}  (deepest first) */
    const float* dec_stem_w[3]; const float* dec_stem_b[3];  /* decoder.stage3.deconv1, stage2.deconv2, stage1.conv3 */
    urnn_head_params head;
} urnn_ed_params;

size_t urnn_ed_step_workspace_bytes(const urnn_ed_desc* d);
/* input (Cin,H,W); states_in/out[6] = e1,e2,e3,d(1/4),d(1/2),d(1x) in the reference's order
 * (utils/general.py:50-95); out (2,H,W) as urnn_head_fwd.  states_out must not alias states_in. */
int urnn_ed_step_fwd(const urnn_ed_desc* d, const urnn_ed_params* p, const float* input,
                     const float* const* states_in, float* const* states_out, float* out,
                     void* ws, size_t ws_bytes, void* stream);

/* The inference loop of test.py:356-375 with HOST buffers: for t in [0,T): copy inputs_host[t] (Cin,H,W)
 * host->device, run one step, copy the masked depth (H,W) device->host into out_host[t].  Copies are
 * double-buffered on internal streams and overlap the next step; the call returns after the last copy
 * has landed (it synchronises -- this is the one blocking entry point).  inputs_host / out_host should be
 * page-locked.  states[6] are device buffers updated in place across the T steps (ping-pong inside ws). */
/* test.py:356-367 with DEVICE buffers: T steps, inputs_dev (T, Cin, H, W) and out_dev (T, H, W) resident in HBM, the six
 * states (reference order, fp32 NCHW) updated in place.  Workspace: urnn_ed_sequence_dev_workspace_bytes.  Stream-ordered,
 * no synchronisation.  prob_dev may be NULL; otherwise it receives the (T, H, W) wet probabilities.
 * URNN_MATH_F16X3 on one GPU: the encoder of step t+1 and the decoder + head of step t run on two internal streams that
 * fork from `stream` at entry and join it before the states are written back, so the call keeps its stream-ordered
 * meaning for the caller (URNN_V2_PIPE=0: everything on `stream`).  The same holds for the two _host entry points. */
size_t urnn_ed_sequence_dev_workspace_bytes(const urnn_ed_desc* d);
int urnn_ed_sequence_dev(const urnn_ed_desc* d, const urnn_ed_params* p, int32_t T,
                         const float* inputs_dev, float* out_dev, float* prob_dev, float* const* states,
                         void* ws, size_t ws_bytes, void* stream);

/* ---- Device-side post-processing of an inference run.
 * Replaces, for predictions that are still in HBM, the reference's host path test.py:468 (r_MinMaxScaler,
 * Dynamic2DFlood.py:379-385, min = 0) + test.py:607-675 compute_metrics: R2, MSE, RMSE, MAE (metres), PeakR2, CSI as streaming
 * reductions over chunks of time steps.  pred_norm_dev: (nsteps, H, W) normalised model output (what urnn_ed_sequence_dev
 * writes); gt_mm_dev: (nsteps, H, W) ground truth in mm; both fp32 device buffers.  Per-element arithmetic is fp32 exactly as
 * numpy's on float32 arrays, sums are fp64, the wet / dry counts are integers (bit-exact).  Call reset once, accumulate for
 * consecutive chunks (t0 = index of the chunk's first step in the event of T steps), finalize once.
 * out12_dev (device, 12 doubles): R2, MSE, RMSE, MAE, PeakR2, CSI, tp, fp, fn, t_peak, elements, 0.  Stream-ordered. */
size_t urnn_metrics_workspace_bytes(int32_t H, int32_t W, int32_t T);
int urnn_metrics_reset(int32_t H, int32_t W, int32_t T, void* ws, size_t ws_bytes, void* stream);
int urnn_metrics_accumulate(int32_t H, int32_t W, int32_t T, int32_t t0, int32_t nsteps, const float* pred_norm_dev,
                            const float* gt_mm_dev, float flood_max, void* ws, size_t ws_bytes, void* stream);
int urnn_metrics_finalize(int32_t H, int32_t W, int32_t T, float flood_thres, void* ws, size_t ws_bytes, double* out12_dev,
                          void* stream);

/* Host-only helper (no GPU work): position of pixel (y, x) of the level-`level` map (0: H x W, 1: H/2 x W/2, 2: H/4 x W/4)
 * inside a channel plane of the library's internal phase-separated layout (URNN_MATH_F16X3), and the
 * padded plane size in *plane_elems.  Returns -1 for an invalid request.  Used by the tests to pin the layout. */
int64_t urnn_layout_index(int32_t H, int32_t W, int32_t level, int32_t y, int32_t x, int64_t* plane_elems);

/* Measurement entry point (URNN_MATH_F16X3): T steps like urnn_ed_sequence_dev (same workspace), every launch bracketed by
 * CUDA events on `stream`; synchronises once per step.  op_ms[i] = mean milliseconds of launch i, names = max_ops records of
 * 24 bytes ("stem1", "enc1.A", "enc1.B", "enc1.blend", ..., "head"), *nops = number of launches per step. */
int urnn_ed_profile_dev(const urnn_ed_desc* d, const urnn_ed_params* p, int32_t T, const float* inputs_dev,
                        float* const* states, void* ws, size_t ws_bytes, void* stream,
                        float* op_ms, char* names, int32_t max_ops, int32_t* nops);

size_t urnn_ed_sequence_host_workspace_bytes(const urnn_ed_desc* d);
int urnn_ed_sequence_host(const urnn_ed_desc* d, const urnn_ed_params* p, int32_t T,
                          const float* inputs_host, float* out_host, float* const* states,
                          void* ws, size_t ws_bytes, void* stream);

/* The same loop for one rainfall EVENT given as the dataset provides it (dataset/Dynamic2DFlood.py:200-236): three
 * static maps and a scalar rainfall series, all on the HOST.  Replaces test.py:356-375 INCLUDING the per-step input
 * assembly preprocess_inputs (dataset/Dynamic2DFlood.py:265-366): the dense (C_in,H,W) tensor is never built.  With
 * scalar rainfall 2*hist of the C_in channels are spatially constant, so their contribution to the stage-1 stem
 * (1x1 conv) is folded into a per-step bias, b_t = b + W[:, :2h] . [rain(t-h+1..t)/rain_max | cumsum(..)/cumsum_max]
 * (zero-padded before the event start), and the stem reads only the 3 normalised static maps
 * (DEM-min)/(max-min), (impervious-0.05)/0.9, manhole.  H2D traffic: 3 maps + 2T scalars per event.
 * d->Cin must equal 2*hist + 3.  states[6] are updated in place; out_host (T,H,W) receives the masked depth. */
typedef struct urnn_event_desc {
    int32_t T, hist;                  /* time steps of the event, historical_nums */
    float rain_max, cumsum_rain_max;  /* MinMaxScaler bounds of the rainfall channels */
    float dem_min, dem_max;           /* per-event DEM bounds (inputs["min_DEM"], inputs["max_DEM"]) */
} urnn_event_desc;
size_t urnn_ed_event_host_workspace_bytes(const urnn_ed_desc* d, const urnn_event_desc* ev);
int urnn_ed_event_host(const urnn_ed_desc* d, const urnn_ed_params* p, const urnn_event_desc* ev,
                       const float* dem_host, const float* impervious_host, const float* manhole_host,
                       const float* rainfall_host, const float* cumsum_rainfall_host,
                       float* out_host, float* const* states, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Spatial sharding across the GPUs of one NVLink/NVSwitch box (one process per GPU).  The reference never shards
 * the grid (its only parallelism is DDP over events, main.py:385-387); this is the B200 scaling axis of the path.
 * Each rank owns a band of rows of every map (input, 6 states, head LayerNorm affine); weights are replicated.
 * All ops are pixel-local for 1x1 filters except the GroupNorm / LayerNorm statistics (ConvRNN.py:97,103;
 * head/network_blocks.py:94), which the kernels all-reduce inside their statistics epilogue through peer-mapped
 * exchange buffers (CUDA IPC over NVLink; no NCCL call, no host round trip, CUDA-graph capturable).  Bands must
 * have equal size (the element count of every normalisation is local count x world).
 *   1. every rank: urnn_comm_local_init(world, rank, handle)  -> allocates its exchange buffer (~70 KB), returns
 *      a 64-byte IPC handle;  2. all-gather the handles by any means (torch.distributed);  3. every rank:
 *      urnn_comm_connect(all_handles).  From then on every statistics reduction of this process is global.
 *      urnn_comm_destroy() returns to single-GPU behaviour.  These three calls synchronise / allocate. */
#define URNN_COMM_HANDLE_BYTES 64
int urnn_comm_local_init(int32_t world, int32_t rank, void* handle_out);
int urnn_comm_connect(const void* all_handles);
int urnn_comm_destroy(void);
int urnn_comm_world(void);   /* 1 when no communicator is active */

#ifdef __cplusplus
}
#endif
#endif /* URNN_B200_H */
