// urnn_internal.h -- internal (C++) interfaces between the translation units of liburnn_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../include/urnn_b200.h"

namespace urnn {

struct CommDev;
void current_comm(CommDev* out, int lane = 0);      // the active cross-GPU communicator (world = 1 when none); lane: capi.cu
int comm_world();

// per-cell workspace view shared by the fp32 and tcgen05 paths
struct CellWsView {
    float *G, *C, *scale1, *shift1, *scale2, *shift2;   // pre-GN gate / candidate maps, folded GN affines
    float2 *partial1, *partial2;
    double2 *total1, *total2;
    unsigned* counter;
    int gx;                                            // partial stride (>= CTAs of any producer kernel)
};
size_t cell_ws_view(const urnn_cell_desc* d, void* ws, size_t ws_bytes, CellWsView* out);
int cgru_blend_launch(const CellWsView& w, const float* h, float* h_out, int F, long N, cudaStream_t st);

// tcgen05 bf16 path (cgru_tc.cu).  x / stem inputs may be fp32 (kind 0) or bf16 (kind 1) maps; stem outputs go
// to exactly one of y_bf16 / y_f32.
// x_plane / y_plane: elements between channel planes (0 = H*W); internal bf16 maps use tc_pad_plane(H*W) so that every
// channel row of a 128-pixel tile is a 16-byte aligned, in-bounds TMA bulk copy.
long tc_pad_plane(long n);
void tc_reset_direction();
int tc_counters_begin(const urnn_cell_desc* d, void* cell_ws, size_t ws_bytes, cudaStream_t st);
void tc_counters_end();
// weight images of a time step (cgru_tc.cu): record the step's GEMMs, convert once, replay
bool tc_recording();
void tc_wimg_begin_record(void* arena, size_t cap);
int tc_wimg_convert(cudaStream_t st);
void tc_wimg_off();
size_t cgru_fwd_bf16_workspace(const urnn_cell_desc* d);
int cgru_fwd_bf16(const urnn_cell_desc* d, const urnn_cell_params* p, const void* x, int xkind, const float* e,
                  const float* h, float* h_out, void* ws, size_t ws_bytes, cudaStream_t st, long x_plane = 0);
int cgru_fwd_bf16_standalone(const urnn_cell_desc* d, const urnn_cell_params* p, const void* x, int xkind, const float* e,
                             const float* h, float* h_out, void* ws, size_t ws_bytes, cudaStream_t st);
int conv1x1_lrelu_fwd_tc(int Cin, int Cout, int H, int W, int pool, float slope, const void* x, int xkind,
                         const float* w, long w_ld, const float* b, __nv_bfloat16* y_bf16, float* y_f32, cudaStream_t st,
                         long x_plane = 0, long y_plane = 0);
int deconv2x2_lrelu_fwd_tc(int Cin, int Cout, int H, int W, float slope, const void* x, int xkind, const float* w,
                           const float* b, __nv_bfloat16* y_bf16, float* y_f32, cudaStream_t st, long y_plane = 0);

// URNN_MATH_F16X3 (urnn_v2.cu): the step on the second-generation pixel GEMM; states stay in the internal split layout
// between the steps of a sequence
struct V2Seq;
long long v2_layout_index(int H, int W, int level, int y, int x, long long* ntot);
size_t v2_step_workspace_bytes(const urnn_ed_desc* d);
int v2_step_fwd_nchw(const urnn_ed_desc* d, const urnn_ed_params* p, const float* x, int cin, const float* w, long long w_ld, const float* b,
                     const float* const* sin, float* const* sout, float* out, void* ws, size_t ws_bytes, cudaStream_t st);
// pipelined: encoder(t+1) and decoder + head (t) on two internal streams (single GPU; URNN_V2_PIPE=0 disables)
V2Seq* v2_seq_begin(const urnn_ed_desc* d, const urnn_ed_params* p, const float* const* states, void* ws, size_t ws_bytes, cudaStream_t st, int* rc,
                    bool pipelined = false);
// one step; work queued on `st` before the call is respected, completion is reported through the two events (not joined
// into `st`): *in_done = the step's input has been consumed, *out_done = `out` / depth_dst / prob_dst are complete
int v2_seq_step(V2Seq* s, int t, const float* x, int cin, const float* w, long long w_ld, const float* b, float* out,
                float* depth_dst, float* prob_dst, cudaStream_t st, cudaEvent_t* in_done, cudaEvent_t* out_done);
int v2_seq_profile(V2Seq* s, int T, const float* inputs, size_t in_elems, int cin, const float* w, long long w_ld, const float* b, float* out,
                   cudaStream_t st, float* op_ms, char* names, int max_ops, int* nops);
int v2_seq_end(V2Seq* s, int T, float* const* states, cudaStream_t st);

// fp32 FFMA path (urnn_fp32.cu)
int cgru_fwd_fp32_passes(const urnn_cell_desc* d, const urnn_cell_params* p, const float* x, const float* e,
                         const float* h, CellWsView* wout, void* ws, size_t ws_bytes, cudaStream_t st);
size_t cgru_fwd_fp32_workspace(const urnn_cell_desc* d);
int cgru_fwd_fp32(const urnn_cell_desc* d, const urnn_cell_params* p, const float* x, const float* e,
                  const float* h, float* h_out, void* ws, size_t ws_bytes, cudaStream_t st);
int conv1x1_lrelu_fwd_fp32(int Cin, int Cout, int H, int W, int pool, float slope, const float* x,
                           const float* w, long w_ld, const float* b, float* y, cudaStream_t st);
int deconv2x2_lrelu_fwd_fp32(int Cin, int Cout, int H, int W, float slope, const float* x, const float* w,
                             const float* b, float* y, cudaStream_t st);
size_t head_fwd_fp32_workspace(int H, int W);
// comm_lane: exchange lane of the LayerNorm statistics (current_comm)
int head_fwd_fp32(int H, int W, float cls_thred, float ln_eps, float slope, const urnn_head_params* p,
                  const float* feat, float* out, void* ws, size_t ws_bytes, cudaStream_t st, int comm_lane = 0);

// backward, fp32 (urnn_bwd.cu)
size_t conv1x1_lrelu_bwd_workspace(int Cin, int Cout, int H, int W, int pool);
int conv1x1_lrelu_bwd_fp32(int Cin, int Cout, int H, int W, int pool, float slope, const float* x, const float* w,
                           const float* b, const float* dy, float* dx, float* dw, float* db, void* ws, size_t ws_bytes,
                           cudaStream_t st);
size_t deconv2x2_lrelu_bwd_workspace(int Cin, int Cout, int H, int W);
int deconv2x2_lrelu_bwd_fp32(int Cin, int Cout, int H, int W, float slope, const float* x, const float* w, const float* b,
                             const float* dy, float* dx, float* dw, float* db, void* ws, size_t ws_bytes, cudaStream_t st);
size_t head_bwd_workspace(int H, int W);
int head_bwd_fp32(int H, int W, float cls_thred, float ln_eps, float slope, const urnn_head_params* p, const float* feat,
                  const float* dout, float* dfeat, const urnn_head_grads* g, void* ws, size_t ws_bytes, cudaStream_t st);
size_t cgru_bwd_workspace(const urnn_cell_desc* d);
int cgru_bwd_fp32(const urnn_cell_desc* d, const urnn_cell_params* p, const float* x, const float* e, const float* h,
                  const float* dh_out, float* dx, float* de, float* dh, const urnn_cell_grads* gr, void* ws, size_t ws_bytes,
                  cudaStream_t st);

}  // namespace urnn
