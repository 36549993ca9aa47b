// plan_check.cu -- host-only check of the launch planner (tc::plan_launch): prints, for every GEMM shape of the
// encoder-decoder step, the producer mode, ring depths and shared-memory size.  Runs without a GPU.
#include <cstdio>
#include <cstring>
#include "../u-rnn_b200/csrc/tc_pixgemm.cuh"
namespace urnn { void set_error(const char*, ...) {} void count_launch(int) {} }
using namespace urnn;

static void one(const char* name, int NOUT, int K, int N, bool gated, int epi, int k0, int kind0, bool padded_out) {
    tc::GemmParams P; memset(&P, 0, sizeof(P));
    const long npad = ((long)N + 127) / 128 * 128;
    char* base = (char*)0x10000000;                      // fake, 256-byte aligned device addresses (never dereferenced)
    P.seg.src[0] = base; P.seg.src[1] = base + (1 << 26); P.seg.src[2] = base + (2 << 26);
    P.seg.cend[0] = k0; P.seg.cend[1] = K; P.seg.cend[2] = K;
    P.seg.kind[0] = kind0; P.seg.plane[0] = kind0 ? npad : N; P.seg.plane[1] = P.seg.plane[2] = N;
    P.seg.gate_seg = gated ? 1 : -1; P.seg.gate_pre = (const __nv_bfloat16*)(base + (3 << 26)); P.seg.gate_plane = npad;
    P.NOUT = NOUT; P.K = K; P.N = N; P.img_w = 500;
    P.out = (__nv_bfloat16*)(base + (4 << 26)); P.out_plane = padded_out ? npad : N; P.bias = (const float*)base;
    const size_t smem = tc::plan_launch(P, epi, true);
    printf("%s NOUT=%d K=%d N=%d gated=%d epi=%d bulk=%d nraw=%d na=%d nstage=%d out_vec=%d smem=%zu cap=%zu\n", name, NOUT, K, N,
           (int)gated, epi, P.bulk, P.nraw, P.na, P.nstage, P.out_vec, smem, tc::SMEM_CAP);
}

int main() {
    const int Ns[3] = {500 * 500, 250 * 250, 125 * 125};
    one("stem1", 32, 63, Ns[0], false, tc::EPI_LRELU, 63, 0, true);
    one("enc1A", 192, 80, Ns[0], false, tc::EPI_GN, 16, 1, true);
    one("enc1B", 64, 64, Ns[0], true, tc::EPI_GN, 0, 0, true);
    one("stem2", 64, 64, Ns[0], false, tc::EPI_POOL, 64, 0, true);
    one("enc2A", 192, 160, Ns[1], false, tc::EPI_GN, 64, 1, true);
    one("enc2B", 96, 160, Ns[1], true, tc::EPI_GN, 64, 1, true);
    one("stem3", 96, 96, Ns[1], false, tc::EPI_POOL, 96, 0, true);
    one("enc3A", 192, 192, Ns[2], false, tc::EPI_GN, 96, 1, true);
    one("enc3B", 96, 192, Ns[2], true, tc::EPI_GN, 96, 1, true);
    one("dec3A", 192, 192, Ns[2], false, tc::EPI_GN, 96, 0, true);
    one("dec3B", 96, 192, Ns[2], true, tc::EPI_GN, 96, 0, true);
    one("up3", 192, 96, Ns[2], false, tc::EPI_DECONV, 96, 0, true);
    one("dec2A", 192, 288, Ns[1], false, tc::EPI_GN, 96, 1, true);
    one("dec2B", 96, 288, Ns[1], true, tc::EPI_GN, 96, 1, true);
    one("up2", 192, 96, Ns[1], false, tc::EPI_DECONV, 96, 0, true);
    one("dec1A", 192, 224, Ns[0], false, tc::EPI_GN, 96, 1, true);
    one("dec1B", 64, 64, Ns[0], true, tc::EPI_GN, 0, 0, true);
    one("final", 32, 64, Ns[0], false, tc::EPI_LRELU, 64, 0, false);
    return 0;
}
