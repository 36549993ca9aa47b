// v2_plan_check.cu -- host-only check of the f16x3 launch planner (v2::plan_gemm / v2::plan_chain, csrc/v2_host.cuh): builds the
// GEMM launches of one encoder-decoder step exactly as csrc/urnn_v2.cu does (shapes of net_params.py:36-137 at a given
// grid), prints ring depth / TMEM layout / shared memory per launch and verifies the precomputed MMA descriptors.
// Runs without a GPU:  nvcc -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 -o v2_plan_check tools/v2_plan_check.cu
//                      ./v2_plan_check [H W]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include "../u-rnn_b200/csrc/v2_host.cuh"
namespace urnn {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
void count_launch(int) {}
}
using namespace urnn;
using namespace urnn::v2;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL %s: ", name); printf(__VA_ARGS__); printf("\n"); ++g_fail; } } while (0)

static void verify(const char* name, GemmLaunch& L, int K, long long ntot, int num_sms) {
    GemmParams& P = L.P;
    P.ntot = ntot;
    const int rc = plan_gemm(L, num_sms);
    if (rc != URNN_OK) { printf("FAIL %s: plan_gemm rc=%d\n", name, rc); ++g_fail; return; }
    const int ncols_total = P.acc_mode == ACC_DECONV ? P.nacc * P.N : P.N;
    const SmemPlan sp = smem_plan(P.nkb, P.nrows, P.nslots, P.gate_ch, P.gdepth, ncols_total);
    CHECK(L.smem == sp.total && L.smem <= (size_t)SMEM_MAX, "smem %zu", L.smem);
    CHECK(P.nslots >= 2 && P.nslots <= 8, "nslots %d", P.nslots);
    CHECK(P.tmem_cols >= 32 && P.tmem_cols <= 512 && (P.tmem_cols & (P.tmem_cols - 1)) == 0, "tmem_cols %d", P.tmem_cols);
    CHECK(P.acc_stages >= 1 && P.acc_stages <= MAX_ACC_STAGES && P.acc_stages * P.nacc * P.acc_stride <= P.tmem_cols, "stages %d x nacc %d x stride %d > %d",
          P.acc_stages, P.nacc, P.acc_stride, P.tmem_cols);
    CHECK(P.acc_stride >= P.N, "acc_stride %d < N %d", P.acc_stride, P.N);
    CHECK(L.grid == (int)(ntot / TILE_M < num_sms ? ntot / TILE_M : num_sms), "grid %d", L.grid);
    if (P.gate_ch) CHECK(P.gdepth == 1 || P.gdepth == 2, "gdepth %d", P.gdepth);
    // MMA descriptors: start addresses inside the ring / the weight image, one overwriting MMA per accumulator, K covered once
    int kcov[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nfirst[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t wbytes = (uint32_t)P.nkb * P.nrows * 128u;
    for (int s = 0; s < P.nsteps; ++s) {
        const Step& st = P.steps[s]; const MmaStep& m = P.msteps[s];
        CHECK(m.nj == (unsigned)(st.unit_ch >> 4) && (m.nj == 1 || m.nj == 2), "step %d nj %u", s, m.nj);
        CHECK((uint32_t)((m.a_desc & 0x3FFF) << 4) == sp.ring_off, "step %d A start %u != ring %u", s, (uint32_t)((m.a_desc & 0x3FFF) << 4), sp.ring_off);
        CHECK(((m.a_desc >> 16) & 0x3FFF) == (uint64_t)((st.unit_ch * 256) >> 4), "step %d A leading byte offset", s);
        CHECK(m.a_lo_delta == (unsigned)((st.unit_ch * 128) >> 4), "step %d lo delta", s);
        for (unsigned j = 0; j < m.nj; ++j) {
            const uint32_t b = (uint32_t)((m.b_desc[j] & 0x3FFF) << 4), k = (uint32_t)(st.kglob + 16 * j);
            CHECK(b >= sp.w_off && b < sp.w_off + wbytes, "step %d B start outside the weight image", s);
            CHECK(b == sp.w_off + (k >> 6) * P.nrows * 128u + ((k & 63) >> 4) * 32u, "step %d B start for k=%u", s, k);
        }
        const int acc = P.acc_mode == ACC_DECONV ? 0 : st.acc;
        CHECK(m.d_off == (unsigned)(acc * P.acc_stride), "step %d accumulator offset", s);
        kcov[acc] += st.unit_ch; nfirst[acc] += m.first == 0;
    }
    const int nacc_k = P.acc_mode == ACC_DECONV ? 1 : P.nacc;
    for (int a = 0; a < nacc_k; ++a) {
        CHECK(kcov[a] == K - P.gate_ch, "accumulator %d covers %d of %d TMA-fed channels", a, kcov[a], K - P.gate_ch);
        CHECK(nfirst[a] == (K - P.gate_ch > 0 ? 1 : 0), "accumulator %d has %d overwriting MMAs", a, nfirst[a]);
    }
    printf("%s N=%d K=%d nacc=%d gate=%d nslots=%d gdepth=%d stages=%d tmem=%d stride=%d grid=%d smem=%zu cap=%d\n", name, P.N, K, P.nacc, P.gate_ch, P.nslots,
           P.gdepth, P.acc_stages, P.tmem_cols, P.acc_stride, L.grid, L.smem, (int)SMEM_MAX);
}

static void fresh(GemmLaunch& L) { memset(&L, 0, sizeof(L)); params_defaults(L.P); L.P.nmma = 3; }

int main(int argc, char** argv) {
    const int H = argc > 2 ? atoi(argv[1]) : 500, W = argc > 2 ? atoi(argv[2]) : 500, num_sms = 148;
    const long long n4p = (((long long)(H / 4) * (W / 4)) + 127) / 128 * 128;
    const long long ntot[3] = {16 * n4p, 4 * n4p, n4p};
    struct Cell { const char* name; int Cx, Ce, F, level; };
    const Cell cells[6] = {{"enc1", 16, 0, 64, 0}, {"enc2", 64, 0, 96, 1}, {"enc3", 96, 0, 96, 2},
                           {"dec3", 0, 96, 96, 2}, {"dec2", 96, 96, 96, 1}, {"dec1", 96, 64, 64, 0}};
    char name[64];
    for (const Cell& c : cells) {
        const int K = c.Cx + c.Ce + c.F, nkb = (K + 63) / 64;
        int nchunk = 1;
        while ((size_t)2 * (2 * c.F / nchunk) * nkb * 128 > (size_t)150 * 1024) nchunk *= 2;
        for (int ci = 0; ci < nchunk; ++ci) {             // sweep A
            GemmLaunch L; fresh(L);
            int k = 0, m = 0;
            if (c.Cx) { k = add_segment_steps(L, m, c.Cx, k, 0, 0); ++m; }
            if (c.Ce) { k = add_segment_steps(L, m, c.Ce, k, 0, 0); ++m; }
            k = add_segment_steps(L, m, c.F, k, 0, 0);
            L.P.N = 2 * c.F / nchunk; L.P.nrows = L.P.N; L.P.nkb = nkb; L.P.epi = EPI_STATS_F32; L.P.nstat = L.P.N / 32;
            snprintf(name, sizeof(name), nchunk > 1 ? "%s.A%d" : "%s.A", c.name, ci + 1);
            verify(name, L, K, ntot[c.level], num_sms);
        }
        {                                                  // sweep B (gated)
            GemmLaunch L; fresh(L);
            int k = 0, m = 0;
            if (c.Cx) { k = add_segment_steps(L, m, c.Cx, k, 0, 0); ++m; }
            if (c.Ce) { k = add_segment_steps(L, m, c.Ce, k, 0, 0); ++m; }
            L.P.gate_ch = c.F; L.P.gate_k0 = k;
            L.P.N = c.F; L.P.nrows = c.F; L.P.nkb = nkb; L.P.epi = EPI_STATS_F32; L.P.nstat = c.F / 32;
            snprintf(name, sizeof(name), "%s.B", c.name);
            verify(name, L, K, ntot[c.level], num_sms);
        }
    }
    const int pools[2][3] = {{64, 64, 1}, {96, 96, 2}};     // Cin, Cout, destination level
    for (auto& p : pools) {
        GemmLaunch L; fresh(L);
        L.P.nacc = 4; L.P.acc_mode = ACC_POOL;
        for (int a = 0; a < 4; ++a) add_segment_steps(L, 0, p[0], 0, a, a * ntot[p[2]]);
        L.P.N = p[1]; L.P.nrows = p[1]; L.P.nkb = (p[0] + 63) / 64; L.P.epi = EPI_LRELU_SPLIT;
        snprintf(name, sizeof(name), "stem%d", p[2] + 1);
        verify(name, L, p[0], ntot[p[2]], num_sms);
    }
    for (int lvl = 2; lvl >= 1; --lvl) {                    // transposed convolutions 96 -> 96, source level lvl
        const int Cin = 96, Cout = 96, nkb = 2;
        int per = 4;
        while (per > 1 && (size_t)2 * per * Cout * nkb * 128 > (size_t)110 * 1024) per /= 2;
        GemmLaunch L; fresh(L);
        L.P.nacc = per; L.P.acc_mode = ACC_DECONV;
        add_segment_steps(L, 0, Cin, 0, 0, 0);
        L.P.N = Cout; L.P.nrows = per * Cout; L.P.nkb = nkb; L.P.epi = EPI_LRELU_SPLIT;
        snprintf(name, sizeof(name), "deconv%d(x%d)", lvl + 1, 4 / per);
        verify(name, L, Cin, ntot[lvl], num_sms);
    }
    {
        GemmLaunch L; fresh(L);
        add_segment_steps(L, 0, 64, 0, 0, 0);
        L.P.N = 16; L.P.nrows = 16; L.P.nkb = 1; L.P.epi = EPI_LRELU_F32;
        snprintf(name, sizeof(name), "stem_out");
        verify(name, L, 64, ntot[0], num_sms);
    }
    // recompute schedule: which cells would fit (opt-in, URNN_V2_Y=1)
    for (const Cell& c : cells) {
        const int K = c.Cx + c.Ce + c.F, nkb = (K + 63) / 64;
        const int s2 = 3 * c.F <= 256 ? chain_slots(nkb, 3 * c.F, c.F, 2) : 0, s1 = 3 * c.F <= 256 ? chain_slots(nkb, 3 * c.F, c.F, 1) : 0;
        printf("chain %s rows=%d K=%d slots_gd2=%d slots_gd1=%d fits=%d\n", c.name, 3 * c.F, K, s2, s1, (int)(s2 >= 3 || s1 >= 3));
    }
    printf("%s (%d failures)\n", g_fail ? "PLAN CHECK FAILED" : "PLAN CHECK PASSED", g_fail);
    return g_fail ? 1 : 0;
}
