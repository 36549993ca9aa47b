// capi.cu -- the extern "C" surface of liburnn_b200.so (see include/urnn_b200.h) and the whole-step driver.
#include <atomic>
#include <mutex>
#include <string.h>

#include "urnn_common.cuh"
#include "urnn_internal.h"

namespace urnn {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int check_cell(const urnn_cell_desc* d, const urnn_cell_params* p, const float* e, const float* h,
                      const float* h_out) {
    URNN_CHECK_ARG(d && p, "cgru: null descriptor/params");
    URNN_CHECK_ARG(d->H > 0 && d->W > 0, "cgru: bad grid %dx%d", d->H, d->W);
    URNN_CHECK_ARG(d->F > 0 && d->F % 32 == 0, "cgru: num_features=%d must be a multiple of 32 (GroupNorm(F//32))", d->F);
    URNN_CHECK_ARG(d->Cx >= 0, "cgru: Cx=%d", d->Cx);
    URNN_CHECK_ARG(d->ksize >= 1 && d->ksize % 2 == 1 && d->ksize <= 7, "cgru: filter_size=%d must be odd and <= 7", d->ksize);
    URNN_CHECK_ARG(d->variant == URNN_CELL_ENCODER || d->variant == URNN_CELL_DECODER, "cgru: bad variant %d", d->variant);
    URNN_CHECK_ARG((long)d->H * d->W < (1L << 30), "cgru: more than 2^30 cells per map are not supported (32-bit pixel indices)");
    URNN_CHECK_ARG(h && h_out, "cgru: null state pointer");
    URNN_CHECK_ARG(((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(h_out)) & 15) == 0, "cgru: h and h_out must be 16-byte aligned");
    URNN_CHECK_ARG((d->variant == URNN_CELL_DECODER) == (e != nullptr), "cgru: e must be given exactly for the decoder variant");
    URNN_CHECK_ARG(p->w1 && p->b1 && p->gn1_w && p->gn1_b && p->w2 && p->b2 && p->gn2_w && p->gn2_b, "cgru: null parameter");
    return URNN_OK;
}

// ---- cross-GPU communicator (process-global: one process drives one GPU)
static CommDev g_comm = {1, 0, {nullptr}, {nullptr}, nullptr};
static void* g_comm_local = nullptr;
static void* g_comm_peer[COMM_MAX_WORLD] = {nullptr};
static const size_t kSlotBytes = (size_t)COMM_RING * COMM_MAX_WORLD * COMM_MAX_SETS * COMM_WORDS * sizeof(unsigned long long);
static const size_t kFlagBytes = (size_t)COMM_RING * COMM_MAX_WORLD * sizeof(unsigned);
static const size_t kCommBytes = kSlotBytes + kFlagBytes + 256;

// Two independent exchange lanes (own slots, flags and counter each) share one IPC allocation: lane 0 carries every
// exchange of the per-operator / fp32 / backward paths and the encoder half of an f16x3 step, lane 1 the decoder half and the
// head of an f16x3 step -- the two halves run on different streams, and a lane needs all ranks to issue its exchanges in
// the same order.
constexpr int kCommLanes = 2;
void current_comm(CommDev* out, int lane) {
    *out = g_comm;
    if (lane <= 0 || lane >= kCommLanes || g_comm.world <= 1) return;
    const size_t off = (size_t)lane * kCommBytes;
    for (int r = 0; r < COMM_MAX_WORLD; ++r) {
        if (out->slots[r]) out->slots[r] = (unsigned long long*)((char*)out->slots[r] + off);
        if (out->flags[r]) out->flags[r] = (unsigned*)((char*)out->flags[r] + off);
    }
    out->seq = (unsigned*)((char*)out->seq + off);
}
int comm_world() { return g_comm.world; }

}  // namespace urnn

using namespace urnn;

extern "C" {

int urnn_abi_version(void) { return URNN_ABI_VERSION; }
const char* urnn_last_error(void) { return g_err; }
uint64_t urnn_launch_count(void) { return g_launches.load(); }

size_t urnn_cgru_fwd_workspace_bytes(const urnn_cell_desc* d) {
    if (!d || d->F <= 0 || d->F % 32) return 0;
    size_t a = cgru_fwd_fp32_workspace(d);
#ifndef URNN_NO_TC
    size_t b = cgru_fwd_bf16_workspace(d);
    if (b > a) a = b;
#endif
    return a;
}

int urnn_cgru_fwd(const urnn_cell_desc* d, const urnn_cell_params* p, const float* x, const float* e,
                  const float* h, float* h_out, void* ws, size_t ws_bytes, void* stream) {
    URNN_TRY(check_cell(d, p, e, h, h_out));
    cudaStream_t st = (cudaStream_t)stream;
#ifndef URNN_NO_TC
    tc_reset_direction();
#endif
    switch (d->math) {
        case URNN_MATH_F16X3:      // per-operator calls of the split mode run the fp32 FFMA kernels (same or better precision)
        case URNN_MATH_FP32: return cgru_fwd_fp32(d, p, x, e, h, h_out, ws, ws_bytes, st);
#ifndef URNN_NO_TC
        case URNN_MATH_BF16: return cgru_fwd_bf16_standalone(d, p, x, 0, e, h, h_out, ws, ws_bytes, st);
#endif
        default: set_error("cgru_fwd: math mode %d not built", d->math); return URNN_E_UNSUPPORTED;
    }
}

size_t urnn_cgru_bwd_workspace_bytes(const urnn_cell_desc* d) {
    if (!d || d->F <= 0 || d->F % 32) return 0;
    return cgru_bwd_workspace(d);
}
int urnn_cgru_bwd(const urnn_cell_desc* d, const urnn_cell_params* p, const float* x, const float* e, const float* h,
                  const float* dh_out, float* dx, float* de, float* dh, const urnn_cell_grads* grads, void* ws,
                  size_t ws_bytes, void* stream) {
    URNN_TRY(check_cell(d, p, e, h, dh_out ? (const float*)dh_out : nullptr));
    URNN_CHECK_ARG(dh_out, "cgru_bwd: null dh_out");
    return cgru_bwd_fp32(d, p, x, e, h, dh_out, dx, de, dh, grads, ws, ws_bytes, (cudaStream_t)stream);
}

int urnn_conv1x1_lrelu_fwd(int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t pool, float slope, int32_t math,
                           const float* x, const float* w, const float* b, float* y, void* stream) {
    URNN_CHECK_ARG(Cin > 0 && Cout > 0 && H > 0 && W > 0, "conv1x1: bad shape");
    URNN_CHECK_ARG(pool == 1 || pool == 2, "conv1x1: pool must be 1 or 2");
    URNN_CHECK_ARG(pool == 1 || (W % 2 == 0 && H % 2 == 0), "conv1x1: AvgPool2 needs even H=%d, W=%d", H, W);
    URNN_CHECK_ARG(pool == 1 || (reinterpret_cast<uintptr_t>(x) & 7) == 0, "conv1x1: pooled input must be 8-byte aligned");
    URNN_CHECK_ARG(x && w && b && y, "conv1x1: null pointer");
#ifndef URNN_NO_TC
    if (math == URNN_MATH_BF16 && Cout <= 256) tc_reset_direction();
    if (math == URNN_MATH_BF16 && Cout <= 256)
        return conv1x1_lrelu_fwd_tc(Cin, Cout, H, W, pool, slope, x, 0, w, Cin, b, nullptr, y, (cudaStream_t)stream);
#endif
    return conv1x1_lrelu_fwd_fp32(Cin, Cout, H, W, pool, slope, x, w, Cin, b, y, (cudaStream_t)stream);
}
size_t urnn_conv1x1_lrelu_bwd_workspace_bytes(int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t pool) {
    return conv1x1_lrelu_bwd_workspace(Cin, Cout, H, W, pool);
}
int urnn_conv1x1_lrelu_bwd(int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t pool, float slope, const float* x,
                           const float* w, const float* b, const float* dy, float* dx, float* dw, float* db, void* ws,
                           size_t ws_bytes, void* stream) {
    URNN_CHECK_ARG(Cin > 0 && Cout > 0 && H > 0 && W > 0 && (pool == 1 || pool == 2), "conv1x1_lrelu_bwd: bad shape");
    URNN_CHECK_ARG(pool == 1 || (H % 2 == 0 && W % 2 == 0), "conv1x1_lrelu_bwd: AvgPool2 needs even H, W");
    URNN_CHECK_ARG(x && w && b && dy && ws, "conv1x1_lrelu_bwd: null pointer");
    URNN_CHECK_ARG((dw == nullptr) == (db == nullptr), "conv1x1_lrelu_bwd: dw and db must be given together");
    return conv1x1_lrelu_bwd_fp32(Cin, Cout, H, W, pool, slope, x, w, b, dy, dx, dw, db, ws, ws_bytes, (cudaStream_t)stream);
}

int urnn_deconv2x2_lrelu_fwd(int32_t Cin, int32_t Cout, int32_t H, int32_t W, float slope, int32_t math, const float* x,
                             const float* w, const float* b, float* y, void* stream) {
    URNN_CHECK_ARG(Cin > 0 && Cout > 0 && H > 0 && W > 0, "deconv2x2: bad shape");
    URNN_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 7) == 0, "deconv2x2: output must be 8-byte aligned");
    URNN_CHECK_ARG(x && w && b && y, "deconv2x2: null pointer");
#ifndef URNN_NO_TC
    if (math == URNN_MATH_BF16) tc_reset_direction();
    if (math == URNN_MATH_BF16)
        return deconv2x2_lrelu_fwd_tc(Cin, Cout, H, W, slope, x, 0, w, b, nullptr, y, (cudaStream_t)stream);
#endif
    return deconv2x2_lrelu_fwd_fp32(Cin, Cout, H, W, slope, x, w, b, y, (cudaStream_t)stream);
}
size_t urnn_deconv2x2_lrelu_bwd_workspace_bytes(int32_t Cin, int32_t Cout, int32_t H, int32_t W) {
    return deconv2x2_lrelu_bwd_workspace(Cin, Cout, H, W);
}
int urnn_deconv2x2_lrelu_bwd(int32_t Cin, int32_t Cout, int32_t H, int32_t W, float slope, const float* x, const float* w,
                             const float* b, const float* dy, float* dx, float* dw, float* db, void* ws, size_t ws_bytes,
                             void* stream) {
    URNN_CHECK_ARG(Cin > 0 && Cout > 0 && H > 0 && W > 0, "deconv2x2_lrelu_bwd: bad shape");
    URNN_CHECK_ARG(x && w && b && dy && ws, "deconv2x2_lrelu_bwd: null pointer");
    URNN_CHECK_ARG((dw == nullptr) == (db == nullptr), "deconv2x2_lrelu_bwd: dw and db must be given together");
    return deconv2x2_lrelu_bwd_fp32(Cin, Cout, H, W, slope, x, w, b, dy, dx, dw, db, ws, ws_bytes, (cudaStream_t)stream);
}

size_t urnn_head_fwd_workspace_bytes(int32_t H, int32_t W) { return head_fwd_fp32_workspace(H, W); }
int urnn_head_fwd(int32_t H, int32_t W, float cls_thred, float ln_eps, float slope, const urnn_head_params* p,
                  const float* feat, float* out, void* ws, size_t ws_bytes, void* stream) {
    URNN_CHECK_ARG(H > 0 && W > 0 && p && feat && out, "head: bad argument");
    return head_fwd_fp32(H, W, cls_thred, ln_eps, slope, p, feat, out, ws, ws_bytes, (cudaStream_t)stream);
}
size_t urnn_head_bwd_workspace_bytes(int32_t H, int32_t W) { return head_bwd_workspace(H, W); }
int urnn_head_bwd(int32_t H, int32_t W, float cls_thred, float ln_eps, float slope, const urnn_head_params* p,
                  const float* feat, const float* dout, float* dfeat, const urnn_head_grads* grads, void* ws,
                  size_t ws_bytes, void* stream) {
    URNN_CHECK_ARG(H > 0 && W > 0 && p && feat && dout && dfeat && grads && ws, "head_bwd: bad argument");
    return head_bwd_fp32(H, W, cls_thred, ln_eps, slope, p, feat, dout, dfeat, grads, ws, ws_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ whole step
struct EdPlan {
    urnn_cell_desc enc[3], dec[3];
    int h[3], w[3];                       // spatial size per scale (1x, 1/2, 1/4)
    float *s[3], *up3, *up2, *feat;       // stem outputs, deconv outputs, decoder features
    void* cell_ws; size_t cell_ws_bytes;
    void* head_ws; size_t head_ws_bytes;
    void* wimg; size_t wimg_bytes;
    size_t total;
};

static int ed_plan(const urnn_ed_desc* d, void* ws, size_t ws_bytes, EdPlan* pl) {
    URNN_CHECK_ARG(d, "ed: null descriptor");
    URNN_CHECK_ARG(d->H > 0 && d->W > 0 && d->H % 4 == 0 && d->W % 4 == 0,
                   "ed: H=%d W=%d must be multiples of 4 (two 2x pools and two 2x deconvs must round-trip)", d->H, d->W);
    URNN_CHECK_ARG(d->ksize == 1, "ed: filter_size=%d; the encoder-decoder only works with 1x1 gates", d->ksize);
    URNN_CHECK_ARG(d->dec_conv[2] == 16, "ed: decoder.conv_out_channels[-1]=%d must be 16 (head width, model.py:62-63)", d->dec_conv[2]);
    URNN_CHECK_ARG(d->Cin > 0, "ed: Cin=%d", d->Cin);
    URNN_CHECK_ARG(d->dec_gru[0] == d->enc_gru[2] && d->dec_gru[1] == d->enc_gru[1] && d->dec_gru[2] == d->enc_gru[0],
                   "ed: decoder gru_channels must mirror the encoder's (skip concat, decoder.py:135)");
    for (int k = 0; k < 3; ++k) { pl->h[k] = d->H >> k; pl->w[k] = d->W >> k; }
    for (int k = 0; k < 3; ++k) {
        pl->enc[k] = urnn_cell_desc{pl->h[k], pl->w[k], d->enc_conv[k], d->enc_gru[k], 1, URNN_CELL_ENCODER, d->math, d->gn_eps};
        int sc = 2 - k;                   // decoder index 0 is the deepest scale
        int cx = (k == 0) ? d->dec_conv[0] : d->dec_conv[k - 1];
        pl->dec[k] = urnn_cell_desc{pl->h[sc], pl->w[sc], cx, d->dec_gru[k], 1, URNN_CELL_DECODER, d->math, d->gn_eps};
    }
    Arena a(ws, ws_bytes);
    // fp32 maps in fp32 mode; bf16 maps with planes padded to whole 128-pixel tiles in bf16 mode (never larger than this)
    auto map_floats = [](int c, int h, int w) { size_t n = (size_t)h * w, np = (n + 127) / 128 * 128; return (size_t)c * (n > (np + 1) / 2 ? n : (np + 1) / 2); };
    for (int k = 0; k < 3; ++k) pl->s[k] = a.take<float>(map_floats(d->enc_conv[k], pl->h[k], pl->w[k]));
    pl->up3 = a.take<float>(map_floats(d->dec_conv[0], pl->h[1], pl->w[1]));
    pl->up2 = a.take<float>(map_floats(d->dec_conv[1], pl->h[0], pl->w[0]));
    pl->feat = a.take<float>((size_t)16 * d->H * d->W);
    size_t cw = 0;
    for (int k = 0; k < 3; ++k) {
        size_t a1 = urnn_cgru_fwd_workspace_bytes(&pl->enc[k]), a2 = urnn_cgru_fwd_workspace_bytes(&pl->dec[k]);
        URNN_CHECK_ARG(a1 && a2, "ed: gru_channels must be multiples of 32");
        cw = a1 > cw ? a1 : cw; cw = a2 > cw ? a2 : cw;
    }
    pl->cell_ws_bytes = cw;
    pl->cell_ws = a.take<char>(cw);
    pl->wimg_bytes = (d->math == URNN_MATH_BF16) ? ((size_t)4 << 20) : 0;       // bf16 weight images of the step's GEMMs
    pl->wimg = a.take<char>(pl->wimg_bytes);
    pl->head_ws_bytes = urnn_head_fwd_workspace_bytes(d->H, d->W);
    pl->head_ws = a.take<char>(pl->head_ws_bytes);
    pl->total = align_up(a.off, 256);
    return URNN_OK;
}

size_t urnn_ed_step_workspace_bytes(const urnn_ed_desc* d) {
#ifndef URNN_NO_TC
    if (d && d->math == URNN_MATH_F16X3) {
        if (d->H <= 0 || d->W <= 0 || d->H % 4 || d->W % 4) { set_error("ed: H=%d W=%d must be positive multiples of 4", d->H, d->W); return 0; }
        return v2_step_workspace_bytes(d) + (size_t)2 * d->H * d->W * sizeof(float) + 256;
    }
#endif
    EdPlan pl;
    if (ed_plan(d, nullptr, 0, &pl) != URNN_OK) return 0;
    return pl.total;
}

// The stage-1 stem as the step sees it.  Dense input: {input, Cin, W, Cin, b}.  Event mode (scalar rainfall): the 2*hist
// rainfall channels are spatially constant, so their contribution W[:, :2h].r_t is folded into a per-step bias and the
// stem only reads the 3 static maps: {static3, 3, W + 2h, Cin, b_t}.
struct Stem1 { const float* x; int cin; const float* w; long w_ld; const float* b; };

static int ed_step_impl(const urnn_ed_desc* d, const urnn_ed_params* p, const Stem1& s1,
                        const float* const* sin, float* const* sout, float* out, void* ws, size_t ws_bytes,
                        void* stream) {
    URNN_CHECK_ARG(d && p && s1.x && sin && sout && out, "ed_step: null pointer");
#ifndef URNN_NO_TC
    if (d->math == URNN_MATH_F16X3) {
        for (int i = 0; i < 6; ++i) URNN_CHECK_ARG(sin[i] && sout[i], "ed_step: state %d null", i);
        return v2_step_fwd_nchw(d, p, s1.x, s1.cin, s1.w, s1.w_ld, s1.b, sin, sout, out, ws, ws_bytes, (cudaStream_t)stream);
    }
#endif
    EdPlan pl;
    URNN_TRY(ed_plan(d, ws, ws_bytes, &pl));
    const float* input = s1.x;
    URNN_CHECK_ARG(p && input && sin && sout && out, "ed_step: null pointer");
    if (pl.total > ws_bytes) { set_error("ed_step: workspace %zu < %zu bytes", ws_bytes, pl.total); return URNN_E_WORKSPACE; }
    for (int i = 0; i < 6; ++i) URNN_CHECK_ARG(sin[i] && sout[i] && sin[i] != sout[i], "ed_step: state %d null or aliased", i);
    const float sl = d->lrelu_slope;
    cudaStream_t st = (cudaStream_t)stream;
#ifndef URNN_NO_TC
    if (d->math == URNN_MATH_BF16) {
        // tcgen05 route: stem outputs that only feed GEMMs are kept as bf16 maps (the consumer would round them anyway)
        __nv_bfloat16* sb[3] = {(__nv_bfloat16*)pl.s[0], (__nv_bfloat16*)pl.s[1], (__nv_bfloat16*)pl.s[2]};
        __nv_bfloat16 *up3 = (__nv_bfloat16*)pl.up3, *up2 = (__nv_bfloat16*)pl.up2;
        long np[3];
        for (int k = 0; k < 3; ++k) np[k] = tc_pad_plane((long)pl.h[k] * pl.w[k]);
        // the launch sequence of the step (stems, 6 cells, deconvs); run once to record the GEMM weights, once for real
        auto body = [&]() -> int {
            tc_reset_direction();
            const float* cur = input; int cin = s1.cin;
            for (int k = 0; k < 3; ++k) {
                int hin = (k == 0) ? pl.h[0] : pl.h[k - 1], win = (k == 0) ? pl.w[0] : pl.w[k - 1];
                URNN_TRY(conv1x1_lrelu_fwd_tc(cin, d->enc_conv[k], hin, win, k == 0 ? 1 : 2, sl, cur, 0,
                                              k == 0 ? s1.w : p->enc_stem_w[k], k == 0 ? s1.w_ld : (long)cin,
                                              k == 0 ? s1.b : p->enc_stem_b[k], sb[k], nullptr, st, 0, np[k]));
                URNN_TRY(check_cell(&pl.enc[k], &p->enc_cell[k], nullptr, sin[k], sout[k]));
                URNN_TRY(cgru_fwd_bf16(&pl.enc[k], &p->enc_cell[k], sb[k], 1, nullptr, sin[k], sout[k], pl.cell_ws, pl.cell_ws_bytes, st, np[k]));
                cur = sout[k]; cin = d->enc_gru[k];
            }
            URNN_TRY(cgru_fwd_bf16(&pl.dec[0], &p->dec_cell[0], nullptr, 0, sout[2], sin[3], sout[3], pl.cell_ws, pl.cell_ws_bytes, st));
            URNN_TRY(deconv2x2_lrelu_fwd_tc(d->dec_gru[0], d->dec_conv[0], pl.h[2], pl.w[2], sl, sout[3], 0, p->dec_stem_w[0],
                                            p->dec_stem_b[0], up3, nullptr, st, np[1]));
            URNN_TRY(cgru_fwd_bf16(&pl.dec[1], &p->dec_cell[1], up3, 1, sout[1], sin[4], sout[4], pl.cell_ws, pl.cell_ws_bytes, st, np[1]));
            URNN_TRY(deconv2x2_lrelu_fwd_tc(d->dec_gru[1], d->dec_conv[1], pl.h[1], pl.w[1], sl, sout[4], 0, p->dec_stem_w[1],
                                            p->dec_stem_b[1], up2, nullptr, st, np[0]));
            URNN_TRY(cgru_fwd_bf16(&pl.dec[2], &p->dec_cell[2], up2, 1, sout[0], sin[5], sout[5], pl.cell_ws, pl.cell_ws_bytes, st, np[0]));
            URNN_TRY(conv1x1_lrelu_fwd_tc(d->dec_gru[2], 16, pl.h[0], pl.w[0], 1, sl, sout[5], 0, p->dec_stem_w[2], d->dec_gru[2],
                                          p->dec_stem_b[2], nullptr, pl.feat, st));
            return URNN_OK;
        };
        URNN_TRY(tc_counters_begin(&pl.enc[0], pl.cell_ws, pl.cell_ws_bytes, st));
        struct StepGuard { ~StepGuard() { tc_counters_end(); tc_wimg_off(); } } step_guard;
        tc_wimg_begin_record(pl.wimg, pl.wimg_bytes);
        URNN_TRY(body());
        URNN_TRY(tc_wimg_convert(st));
        URNN_TRY(body());
        return urnn_head_fwd(d->H, d->W, d->cls_thred, d->ln_eps, sl, &p->head, pl.feat, out, pl.head_ws, pl.head_ws_bytes, stream);
    }
#endif
    // ---- encoder (encoder.py:187-215): stem conv (+pool) then ConvGRU, stage k feeds stage k+1
    const float* cur = input; int cin = s1.cin;
    for (int k = 0; k < 3; ++k) {
        int hin = (k == 0) ? pl.h[0] : pl.h[k - 1], win = (k == 0) ? pl.w[0] : pl.w[k - 1];
        URNN_TRY(conv1x1_lrelu_fwd_fp32(cin, d->enc_conv[k], hin, win, k == 0 ? 1 : 2, sl, cur,
                                        k == 0 ? s1.w : p->enc_stem_w[k], k == 0 ? s1.w_ld : (long)cin,
                                        k == 0 ? s1.b : p->enc_stem_b[k], pl.s[k], st));
        URNN_TRY(urnn_cgru_fwd(&pl.enc[k], &p->enc_cell[k], pl.s[k], nullptr, sin[k], sout[k], pl.cell_ws,
                               pl.cell_ws_bytes, stream));
        cur = sout[k]; cin = d->enc_gru[k];
    }
    // ---- decoder (decoder.py:173-217): deepest first; x of the deepest stage is None -> zeros
    URNN_TRY(urnn_cgru_fwd(&pl.dec[0], &p->dec_cell[0], nullptr, sout[2], sin[3], sout[3], pl.cell_ws, pl.cell_ws_bytes, stream));
    URNN_TRY(urnn_deconv2x2_lrelu_fwd(d->dec_gru[0], d->dec_conv[0], pl.h[2], pl.w[2], sl, d->math, sout[3], p->dec_stem_w[0],
                                      p->dec_stem_b[0], pl.up3, stream));
    URNN_TRY(urnn_cgru_fwd(&pl.dec[1], &p->dec_cell[1], pl.up3, sout[1], sin[4], sout[4], pl.cell_ws, pl.cell_ws_bytes, stream));
    URNN_TRY(urnn_deconv2x2_lrelu_fwd(d->dec_gru[1], d->dec_conv[1], pl.h[1], pl.w[1], sl, d->math, sout[4], p->dec_stem_w[1],
                                      p->dec_stem_b[1], pl.up2, stream));
    URNN_TRY(urnn_cgru_fwd(&pl.dec[2], &p->dec_cell[2], pl.up2, sout[0], sin[5], sout[5], pl.cell_ws, pl.cell_ws_bytes, stream));
    URNN_TRY(urnn_conv1x1_lrelu_fwd(d->dec_gru[2], 16, pl.h[0], pl.w[0], 1, sl, d->math, sout[5], p->dec_stem_w[2], p->dec_stem_b[2],
                                    pl.feat, stream));
    // ---- head (flood_head.py:131-177)
    URNN_TRY(urnn_head_fwd(d->H, d->W, d->cls_thred, d->ln_eps, sl, &p->head, pl.feat, out, pl.head_ws, pl.head_ws_bytes, stream));
    return URNN_OK;
}

// ------------------------------------------------------------------------------------------------ host-buffer sequence
// test.py:356-375 with host buffers: H2D of step t+1's input and D2H of step t-1's depth map overlap step t.
// Copy streams / events of the host-buffer entry points: one set per device, created on first use on that device
// (these entry points are the only ones that own CUDA objects).
struct HostIo { cudaStream_t s_in = nullptr, s_out = nullptr; cudaEvent_t ev_in[2], ev_step[2], ev_out[2], ev_free[2], ev_start; };
static std::mutex g_io_mu;
static HostIo g_io[64];
static bool g_io_ready[64] = {false};
static int host_io(HostIo** out) {
    int dev = 0;
    URNN_CUDA(cudaGetDevice(&dev));
    URNN_CHECK_ARG(dev >= 0 && dev < 64, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_io_mu);
    HostIo& io = g_io[dev];
    if (!g_io_ready[dev]) {
        URNN_CUDA(cudaStreamCreateWithFlags(&io.s_in, cudaStreamNonBlocking));
        URNN_CUDA(cudaStreamCreateWithFlags(&io.s_out, cudaStreamNonBlocking));
        URNN_CUDA(cudaEventCreateWithFlags(&io.ev_start, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            URNN_CUDA(cudaEventCreateWithFlags(&io.ev_in[i], cudaEventDisableTiming));
            URNN_CUDA(cudaEventCreateWithFlags(&io.ev_step[i], cudaEventDisableTiming));
            URNN_CUDA(cudaEventCreateWithFlags(&io.ev_out[i], cudaEventDisableTiming));
            URNN_CUDA(cudaEventCreateWithFlags(&io.ev_free[i], cudaEventDisableTiming));
        }
        g_io_ready[dev] = true;
    }
    *out = &io;
    return URNN_OK;
}

// frees a sequence context on every exit path
struct SeqGuard { V2Seq* seq = nullptr; ~SeqGuard() { if (seq) v2_seq_end(seq, 0, nullptr, nullptr); } };

struct SeqPlan {
    float* in[2]; float* out[2]; float* st[2][6]; void* step_ws; size_t step_ws_bytes; size_t total;
    size_t state_elems[6];
};
static int seq_plan(const urnn_ed_desc* d, void* ws, size_t ws_bytes, SeqPlan* sp) {
    EdPlan pl;
    URNN_TRY(ed_plan(d, nullptr, 0, &pl));
    const size_t N = (size_t)d->H * d->W;
    const int ch[6] = {d->enc_gru[0], d->enc_gru[1], d->enc_gru[2], d->dec_gru[0], d->dec_gru[1], d->dec_gru[2]};
    const int sc[6] = {0, 1, 2, 2, 1, 0};
    Arena a(ws, ws_bytes);
    for (int i = 0; i < 2; ++i) { sp->in[i] = a.take<float>((size_t)d->Cin * N); sp->out[i] = a.take<float>(2 * N); }
    for (int k = 0; k < 6; ++k) {
        sp->state_elems[k] = (size_t)ch[k] * (N >> (2 * sc[k]));
        sp->st[0][k] = nullptr;                                   // ping buffer = the caller's state buffer
        sp->st[1][k] = a.take<float>(sp->state_elems[k]);
    }
    sp->step_ws_bytes = urnn_ed_step_workspace_bytes(d);
    if (sp->step_ws_bytes == 0) return URNN_E_INVALID;
    sp->step_ws = a.take<char>(sp->step_ws_bytes);
    sp->total = align_up(a.off, 256);
    return URNN_OK;
}

int urnn_ed_step_fwd(const urnn_ed_desc* d, const urnn_ed_params* p, const float* input,
                     const float* const* sin, float* const* sout, float* out, void* ws, size_t ws_bytes,
                     void* stream) {
    URNN_CHECK_ARG(d && p, "ed_step: null descriptor/params");
    Stem1 s1{input, d->Cin, p->enc_stem_w[0], (long)d->Cin, p->enc_stem_b[0]};
    return ed_step_impl(d, p, s1, sin, sout, out, ws, ws_bytes, stream);
}

size_t urnn_ed_sequence_host_workspace_bytes(const urnn_ed_desc* d) {
    SeqPlan sp;
    if (seq_plan(d, nullptr, 0, &sp) != URNN_OK) return 0;
    return sp.total;
}

int urnn_ed_sequence_host(const urnn_ed_desc* d, const urnn_ed_params* p, int32_t T, const float* inputs_host,
                          float* out_host, float* const* states, void* ws, size_t ws_bytes, void* stream) {
    URNN_CHECK_ARG(d && p && inputs_host && out_host && states && T > 0, "ed_sequence_host: bad argument");
    SeqPlan sp;
    URNN_TRY(seq_plan(d, ws, ws_bytes, &sp));
    if (sp.total > ws_bytes) { set_error("ed_sequence_host: workspace %zu < %zu bytes", ws_bytes, sp.total); return URNN_E_WORKSPACE; }
    for (int k = 0; k < 6; ++k) { URNN_CHECK_ARG(states[k], "ed_sequence_host: null state %d", k); sp.st[0][k] = states[k]; }
    HostIo* io = nullptr;
    URNN_TRY(host_io(&io));
    cudaStream_t s_in = io->s_in, s_out = io->s_out;
    cudaEvent_t *ev_in = io->ev_in, *ev_step = io->ev_step, *ev_out = io->ev_out, *ev_free = io->ev_free;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = (size_t)d->H * d->W, in_elems = (size_t)d->Cin * N;
    // the copy streams must not run ahead of work already queued on the caller's stream
    URNN_CUDA(cudaEventRecord(ev_free[0], st));
    URNN_CUDA(cudaStreamWaitEvent(s_in, ev_free[0], 0));
    URNN_CUDA(cudaStreamWaitEvent(s_out, ev_free[0], 0));
    URNN_CUDA(cudaMemcpyAsync(sp.in[0], inputs_host, in_elems * sizeof(float), cudaMemcpyHostToDevice, s_in));
    URNN_CUDA(cudaEventRecord(ev_in[0], s_in));
    // URNN_MATH_F16X3: the states stay in the internal split layout for the whole sequence
    SeqGuard sg; V2Seq*& seq = sg.seq;
#ifndef URNN_NO_TC
    if (d->math == URNN_MATH_F16X3) {
        int rc = URNN_OK;
        seq = v2_seq_begin(d, p, states, sp.step_ws, sp.step_ws_bytes, st, &rc, true);
        if (!seq) return rc;
    }
#endif
    cudaEvent_t e_in[2] = {nullptr, nullptr}, e_out[2] = {nullptr, nullptr};   // f16x3: completion events of the pipelined steps
    for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        if (t + 1 < T) {                                    // prefetch the next input into the other buffer
            if (t >= 1) URNN_CUDA(cudaStreamWaitEvent(s_in, seq ? e_in[b ^ 1] : ev_step[b ^ 1], 0));   // step t-1 has consumed it
            URNN_CUDA(cudaMemcpyAsync(sp.in[b ^ 1], inputs_host + (size_t)(t + 1) * in_elems, in_elems * sizeof(float),
                                      cudaMemcpyHostToDevice, s_in));
            URNN_CUDA(cudaEventRecord(ev_in[b ^ 1], s_in));
        }
        URNN_CUDA(cudaStreamWaitEvent(st, ev_in[b], 0));
        if (t >= 2) URNN_CUDA(cudaStreamWaitEvent(st, ev_out[b], 0));             // out[b] has been drained to the host
        const float* sin[6]; float* sout[6];
        for (int k = 0; k < 6; ++k) { sin[k] = sp.st[b][k]; sout[k] = sp.st[b ^ 1][k]; }
        if (seq) {
            URNN_TRY(v2_seq_step(seq, t, sp.in[b], d->Cin, p->enc_stem_w[0], (long long)d->Cin, p->enc_stem_b[0], sp.out[b], nullptr, nullptr, st,
                                 &e_in[b], &e_out[b]));
            URNN_CUDA(cudaStreamWaitEvent(s_out, e_out[b], 0));
        } else {
            URNN_TRY(urnn_ed_step_fwd(d, p, sp.in[b], sin, sout, sp.out[b], sp.step_ws, sp.step_ws_bytes, stream));
            URNN_CUDA(cudaEventRecord(ev_step[b], st));
            URNN_CUDA(cudaStreamWaitEvent(s_out, ev_step[b], 0));
        }
        URNN_CUDA(cudaMemcpyAsync(out_host + (size_t)t * N, sp.out[b], N * sizeof(float), cudaMemcpyDeviceToHost, s_out));
        URNN_CUDA(cudaEventRecord(ev_out[b], s_out));
    }
    if (seq) { V2Seq* q = seq; seq = nullptr; URNN_TRY(v2_seq_end(q, T, states, st)); }
    else if (T & 1) {                                       // final states live in the workspace buffers: copy back
        for (int k = 0; k < 6; ++k)
            URNN_CUDA(cudaMemcpyAsync(sp.st[0][k], sp.st[1][k], sp.state_elems[k] * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    URNN_CUDA(cudaStreamSynchronize(s_out));
    URNN_CUDA(cudaStreamSynchronize(st));
    return URNN_OK;
}

int64_t urnn_layout_index(int32_t H, int32_t W, int32_t level, int32_t y, int32_t x, int64_t* plane_elems) {
#ifndef URNN_NO_TC
    long long n = 0;
    const long long r = v2_layout_index(H, W, level, y, x, &n);
    if (plane_elems) *plane_elems = (int64_t)n;
    return (int64_t)r;
#else
    (void)H; (void)W; (void)level; (void)y; (void)x; (void)plane_elems; return -1;
#endif
}

// ------------------------------------------------------------------------------------------------ device-buffer sequence
size_t urnn_ed_sequence_dev_workspace_bytes(const urnn_ed_desc* d) {
    SeqPlan sp;
    if (seq_plan(d, nullptr, 0, &sp) != URNN_OK) return 0;
    return sp.total;
}

int urnn_ed_sequence_dev(const urnn_ed_desc* d, const urnn_ed_params* p, int32_t T, const float* inputs_dev, float* out_dev,
                         float* prob_dev, float* const* states, void* ws, size_t ws_bytes, void* stream) {
    URNN_CHECK_ARG(d && p && inputs_dev && out_dev && states && T > 0, "ed_sequence_dev: bad argument");
    SeqPlan sp;
    URNN_TRY(seq_plan(d, ws, ws_bytes, &sp));
    if (sp.total > ws_bytes) { set_error("ed_sequence_dev: workspace %zu < %zu bytes", ws_bytes, sp.total); return URNN_E_WORKSPACE; }
    for (int k = 0; k < 6; ++k) { URNN_CHECK_ARG(states[k], "ed_sequence_dev: null state %d", k); sp.st[0][k] = states[k]; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = (size_t)d->H * d->W, in_elems = (size_t)d->Cin * N;
    SeqGuard sg; V2Seq*& seq = sg.seq;
#ifndef URNN_NO_TC
    if (d->math == URNN_MATH_F16X3) {
        int rc = URNN_OK;
        seq = v2_seq_begin(d, p, states, sp.step_ws, sp.step_ws_bytes, st, &rc, true);
        if (!seq) return rc;
    }
#endif
    for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        const float* x = inputs_dev + (size_t)t * in_elems;
        if (seq) {                                          // the two planes of the step's result are copied out on the decoder's stream
            URNN_TRY(v2_seq_step(seq, t, x, d->Cin, p->enc_stem_w[0], (long long)d->Cin, p->enc_stem_b[0], sp.out[0], out_dev + (size_t)t * N,
                                 prob_dev ? prob_dev + (size_t)t * N : nullptr, st, nullptr, nullptr));
            continue;
        }
        const float* sin[6]; float* sout[6];
        for (int k = 0; k < 6; ++k) { sin[k] = sp.st[b][k]; sout[k] = sp.st[b ^ 1][k]; }
        URNN_TRY(urnn_ed_step_fwd(d, p, x, sin, sout, sp.out[0], sp.step_ws, sp.step_ws_bytes, stream));
        URNN_CUDA(cudaMemcpyAsync(out_dev + (size_t)t * N, sp.out[0], N * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (prob_dev) URNN_CUDA(cudaMemcpyAsync(prob_dev + (size_t)t * N, sp.out[0] + N, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (seq) { V2Seq* q = seq; seq = nullptr; URNN_TRY(v2_seq_end(q, T, states, st)); }
    else if (T & 1) {
        for (int k = 0; k < 6; ++k)
            URNN_CUDA(cudaMemcpyAsync(sp.st[0][k], sp.st[1][k], sp.state_elems[k] * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return URNN_OK;
}

int urnn_ed_profile_dev(const urnn_ed_desc* d, const urnn_ed_params* p, int32_t T, const float* inputs_dev, float* const* states,
                        void* ws, size_t ws_bytes, void* stream, float* op_ms, char* names, int32_t max_ops, int32_t* nops) {
    URNN_CHECK_ARG(d && p && inputs_dev && states && T > 0 && op_ms && names && nops, "ed_profile_dev: bad argument");
#ifndef URNN_NO_TC
    URNN_CHECK_ARG(d->math == URNN_MATH_F16X3, "ed_profile_dev: only URNN_MATH_F16X3 is instrumented");
    SeqPlan sp;
    URNN_TRY(seq_plan(d, ws, ws_bytes, &sp));
    if (sp.total > ws_bytes) { set_error("ed_profile_dev: workspace %zu < %zu bytes", ws_bytes, sp.total); return URNN_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    SeqGuard sg; int rc = URNN_OK;
    sg.seq = v2_seq_begin(d, p, states, sp.step_ws, sp.step_ws_bytes, st, &rc);
    if (!sg.seq) return rc;
    int n = 0;
    URNN_TRY(v2_seq_profile(sg.seq, T, inputs_dev, (size_t)d->Cin * d->H * d->W, d->Cin, p->enc_stem_w[0], (long long)d->Cin, p->enc_stem_b[0],
                            sp.out[0], st, op_ms, names, max_ops, &n));
    *nops = n;
    V2Seq* q = sg.seq; sg.seq = nullptr;
    return v2_seq_end(q, T, states, st);
#else
    set_error("ed_profile_dev: built without the tcgen05 path"); return URNN_E_UNSUPPORTED;
#endif
}

// ------------------------------------------------------------------------------------------------ event mode
__global__ void event_static_maps_kernel(const float* __restrict__ dem, const float* __restrict__ imp,
                                         const float* __restrict__ man, float dem_min, float dem_max,
                                         float* __restrict__ out, long N) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    out[i] = (dem[i] - dem_min) / (dem_max - dem_min);          // MinMaxScaler, Dynamic2DFlood.py:289-292,369-377
    out[N + i] = (imp[i] - 0.05f) / (0.95f - 0.05f);
    out[2 * N + i] = (man[i] - 0.f) / (1.f - 0.f);
}

// bias_all[t][o] = b[o] + sum_j W[o][j] rain_n(t-h+1+j) + sum_j W[o][h+j] cum_n(t-h+1+j)   (zero before the event start)
__global__ void event_fold_rain_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ rain,
                                       const float* __restrict__ cum, int T, int hist, int cin, int cout, float rain_max,
                                       float cum_max, float* __restrict__ bias_all) {
    int t = blockIdx.x, o = threadIdx.x;
    if (o >= cout) return;
    float acc = b[o];
    for (int j = 0; j < hist; ++j) {
        int s = t - hist + 1 + j;
        if (s < 0) continue;
        acc = fmaf(w[(long)o * cin + j], (rain[s] - 0.f) / (rain_max - 0.f), acc);
        acc = fmaf(w[(long)o * cin + hist + j], (cum[s] - 0.f) / (cum_max - 0.f), acc);
    }
    bias_all[(long)t * cout + o] = acc;
}

struct EventPlan { float* raw; float* maps; float* series; float* bias_all; float* out[2]; float* st[6]; size_t state_elems[6];
                   void* step_ws; size_t step_ws_bytes; size_t total; };
static int event_plan(const urnn_ed_desc* d, const urnn_event_desc* ev, void* ws, size_t ws_bytes, EventPlan* ep) {
    URNN_CHECK_ARG(d && ev && ev->T > 0 && ev->hist >= 0, "ed_event: bad descriptor");
    URNN_CHECK_ARG(d->Cin == 2 * ev->hist + 3, "ed_event: Cin=%d must equal 2*hist+3 (hist=%d)", d->Cin, ev->hist);
    EdPlan pl;
    URNN_TRY(ed_plan(d, nullptr, 0, &pl));
    const size_t N = (size_t)d->H * d->W;
    const int ch[6] = {d->enc_gru[0], d->enc_gru[1], d->enc_gru[2], d->dec_gru[0], d->dec_gru[1], d->dec_gru[2]};
    const int sc[6] = {0, 1, 2, 2, 1, 0};
    Arena a(ws, ws_bytes);
    ep->raw = a.take<float>(3 * N);
    ep->maps = a.take<float>(3 * N);
    ep->series = a.take<float>(2 * (size_t)ev->T);
    ep->bias_all = a.take<float>((size_t)ev->T * d->enc_conv[0]);
    for (int i = 0; i < 2; ++i) ep->out[i] = a.take<float>(2 * N);
    for (int k = 0; k < 6; ++k) { ep->state_elems[k] = (size_t)ch[k] * (N >> (2 * sc[k])); ep->st[k] = a.take<float>(ep->state_elems[k]); }
    ep->step_ws_bytes = urnn_ed_step_workspace_bytes(d);
    if (ep->step_ws_bytes == 0) return URNN_E_INVALID;
    ep->step_ws = a.take<char>(ep->step_ws_bytes);
    ep->total = align_up(a.off, 256);
    return URNN_OK;
}

size_t urnn_ed_event_host_workspace_bytes(const urnn_ed_desc* d, const urnn_event_desc* ev) {
    EventPlan ep;
    if (event_plan(d, ev, nullptr, 0, &ep) != URNN_OK) return 0;
    return ep.total;
}

int urnn_ed_event_host(const urnn_ed_desc* d, const urnn_ed_params* p, const urnn_event_desc* ev, const float* dem_host,
                       const float* impervious_host, const float* manhole_host, const float* rainfall_host,
                       const float* cumsum_rainfall_host, float* out_host, float* const* states, void* ws,
                       size_t ws_bytes, void* stream) {
    URNN_CHECK_ARG(p && dem_host && impervious_host && manhole_host && rainfall_host && cumsum_rainfall_host && out_host && states,
                   "ed_event: null pointer");
    EventPlan ep;
    URNN_TRY(event_plan(d, ev, ws, ws_bytes, &ep));
    if (ep.total > ws_bytes) { set_error("ed_event: workspace %zu < %zu bytes", ws_bytes, ep.total); return URNN_E_WORKSPACE; }
    HostIo* io = nullptr;
    URNN_TRY(host_io(&io));
    cudaStream_t s_out = io->s_out;
    cudaEvent_t *ev_step = io->ev_step, *ev_out = io->ev_out, ev_start = io->ev_start;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = (size_t)d->H * d->W;
    const int T = ev->T, cout0 = d->enc_conv[0];
    // ---- once per event: upload (3 maps + 2T scalars), normalise, fold the rainfall into T stage-1 biases
    URNN_CUDA(cudaMemcpyAsync(ep.raw, dem_host, N * sizeof(float), cudaMemcpyHostToDevice, st));
    URNN_CUDA(cudaMemcpyAsync(ep.raw + N, impervious_host, N * sizeof(float), cudaMemcpyHostToDevice, st));
    URNN_CUDA(cudaMemcpyAsync(ep.raw + 2 * N, manhole_host, N * sizeof(float), cudaMemcpyHostToDevice, st));
    URNN_CUDA(cudaMemcpyAsync(ep.series, rainfall_host, T * sizeof(float), cudaMemcpyHostToDevice, st));
    URNN_CUDA(cudaMemcpyAsync(ep.series + T, cumsum_rainfall_host, T * sizeof(float), cudaMemcpyHostToDevice, st));
    event_static_maps_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(ep.raw, ep.raw + N, ep.raw + 2 * N, ev->dem_min, ev->dem_max, ep.maps, (long)N);
    URNN_LAUNCH_CHECK();
    URNN_CHECK_ARG(cout0 <= 1024, "ed_event: stage-1 width %d too large", cout0);
    event_fold_rain_kernel<<<T, ((cout0 + 31) / 32) * 32, 0, st>>>(p->enc_stem_w[0], p->enc_stem_b[0], ep.series, ep.series + T, T, ev->hist,
                                                                   d->Cin, cout0, ev->rain_max, ev->cumsum_rain_max, ep.bias_all);
    URNN_LAUNCH_CHECK();
    URNN_CUDA(cudaEventRecord(ev_start, st));
    URNN_CUDA(cudaStreamWaitEvent(s_out, ev_start, 0));
    SeqGuard sg; V2Seq*& seq = sg.seq;
#ifndef URNN_NO_TC
    if (d->math == URNN_MATH_F16X3) {
        int rc = URNN_OK;
        seq = v2_seq_begin(d, p, states, ep.step_ws, ep.step_ws_bytes, st, &rc, true);
        if (!seq) return rc;
    }
#endif
    cudaEvent_t e_out[2] = {nullptr, nullptr};
    // ---- T steps, states ping-pong between the caller's buffers and the workspace
    for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        if (t >= 2) URNN_CUDA(cudaStreamWaitEvent(st, ev_out[b], 0));
        const float* sin[6]; float* sout[6];
        for (int k = 0; k < 6; ++k) { sin[k] = b ? ep.st[k] : states[k]; sout[k] = b ? states[k] : ep.st[k]; }
        Stem1 s1{ep.maps, 3, p->enc_stem_w[0] + 2 * ev->hist, (long)d->Cin, ep.bias_all + (size_t)t * cout0};
        if (seq) {
            URNN_TRY(v2_seq_step(seq, t, s1.x, s1.cin, s1.w, s1.w_ld, s1.b, ep.out[b], nullptr, nullptr, st, nullptr, &e_out[b]));
            URNN_CUDA(cudaStreamWaitEvent(s_out, e_out[b], 0));
        } else {
            URNN_TRY(ed_step_impl(d, p, s1, sin, sout, ep.out[b], ep.step_ws, ep.step_ws_bytes, stream));
            URNN_CUDA(cudaEventRecord(ev_step[b], st));
            URNN_CUDA(cudaStreamWaitEvent(s_out, ev_step[b], 0));
        }
        URNN_CUDA(cudaMemcpyAsync(out_host + (size_t)t * N, ep.out[b], N * sizeof(float), cudaMemcpyDeviceToHost, s_out));
        URNN_CUDA(cudaEventRecord(ev_out[b], s_out));
    }
    if (seq) { V2Seq* q = seq; seq = nullptr; URNN_TRY(v2_seq_end(q, T, states, st)); }
    else if (T & 1) {
        for (int k = 0; k < 6; ++k)
            URNN_CUDA(cudaMemcpyAsync(states[k], ep.st[k], ep.state_elems[k] * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    URNN_CUDA(cudaStreamSynchronize(s_out));
    URNN_CUDA(cudaStreamSynchronize(st));
    return URNN_OK;
}

// ------------------------------------------------------------------------------------------------ communicator
int urnn_comm_world(void) { return g_comm.world; }

int urnn_comm_local_init(int32_t world, int32_t rank, void* handle_out) {
    URNN_CHECK_ARG(world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world && handle_out, "comm_local_init: bad argument");
    URNN_CHECK_ARG(sizeof(cudaIpcMemHandle_t) <= URNN_COMM_HANDLE_BYTES, "comm_local_init: IPC handle larger than 64 bytes");
    if (g_comm_local) urnn_comm_destroy();
    URNN_CUDA(cudaMalloc(&g_comm_local, kCommLanes * kCommBytes));
    URNN_CUDA(cudaMemset(g_comm_local, 0, kCommLanes * kCommBytes));
    URNN_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    URNN_CUDA(cudaIpcGetMemHandle(&h, g_comm_local));
    memset(handle_out, 0, URNN_COMM_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    g_comm.world = 1; g_comm.rank = rank;          // not active until urnn_comm_connect
    g_comm_peer[rank] = g_comm_local;
    g_comm.seq = (unsigned*)((char*)g_comm_local + kSlotBytes + kFlagBytes);
    // remember the intended world in the (still inactive) communicator through the peer table size
    for (int r = 0; r < COMM_MAX_WORLD; ++r) if (r != rank) g_comm_peer[r] = nullptr;
    g_comm.slots[rank] = (unsigned long long*)g_comm_local;
    g_comm.flags[rank] = (unsigned*)((char*)g_comm_local + kSlotBytes);
    // world is stored negated until connected
    g_comm.world = -world;
    return URNN_OK;
}

int urnn_comm_connect(const void* all_handles) {
    URNN_CHECK_ARG(all_handles && g_comm_local && g_comm.world < 0, "comm_connect: call urnn_comm_local_init first");
    const int world = -g_comm.world, rank = g_comm.rank;
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles + (size_t)r * URNN_COMM_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        URNN_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        g_comm_peer[r] = ptr;
        g_comm.slots[r] = (unsigned long long*)ptr;
        g_comm.flags[r] = (unsigned*)((char*)ptr + kSlotBytes);
    }
    g_comm.world = world;
    return URNN_OK;
}

int urnn_comm_destroy(void) {
    cudaDeviceSynchronize();
    const int world = g_comm.world < 0 ? -g_comm.world : g_comm.world;
    for (int r = 0; r < world && r < COMM_MAX_WORLD; ++r)
        if (r != g_comm.rank && g_comm_peer[r]) { cudaIpcCloseMemHandle(g_comm_peer[r]); g_comm_peer[r] = nullptr; }
    if (g_comm_local) { cudaFree(g_comm_local); g_comm_local = nullptr; }
    g_comm = CommDev{1, 0, {nullptr}, {nullptr}, nullptr};
    return URNN_OK;
}

}  // extern "C"
