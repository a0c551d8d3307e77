// emu.cu — CPU emulation of the CTA-level CMux phases (cmux_core.cuh), thread by thread, so that
// the exact code the kernels execute can be checked against the oracle without a GPU.
// TEST INFRASTRUCTURE: built host-only (nvcc, no device code is launched), loaded by tests/.
#include <cstring>
#include <vector>
#include "cmux_core.cuh"
#include "twiddle_host.h"

template <int K, int G>
struct EmuCta {
    CmuxSmem<K, G> sm;
    std::vector<CmuxRegs<K, G>> rg;
    EmuCta() : rg(CMUX_THREADS) {
        make_twiddle_table(sm.tw);
        memset(sm.acc, 0, sizeof(sm.acc));
        for (int g = 0; g < G; g++) sm.rot[g] = 0;
    }
#define ALL(stmt) for (int tid = 0; tid < CMUX_THREADS; tid++) { stmt; }
    template <int BASE_LOG, int LEVELS, int MODE>
    void cmux_step(const cd *ggsw, const uint64_t *const *ext) {
        ALL((phase_load_decompose<K, G, BASE_LOG, LEVELS, MODE>(tid, sm, rg[tid], ext)));
        for (int lev = LEVELS; lev >= 1; lev--) {
            if (lev != LEVELS) ALL((phase_next_digits<K, G, BASE_LOG, LEVELS>(tid, rg[tid], lev)));
            ALL((phase_fwd1<K, G>(tid, sm, rg[tid])));
            ALL((phase_fwd2<K, G>(tid, sm, rg[tid])));
            ALL((phase_fwd3<K, G>(tid, sm, rg[tid])));
            ALL((phase_mac<K, G>(tid, sm, rg[tid], ggsw + (size_t)(LEVELS - lev) * (K + 1) * POLY_M * (K + 1))));
        }
        ALL((phase_inv0<K, G>(tid, sm, rg[tid])));
        ALL((phase_inv1<K, G>(tid, sm, rg[tid])));
        ALL((phase_inv2<K, G>(tid, sm, rg[tid])));
        ALL((phase_inv3<K, G>(tid, sm, rg[tid])));
    }
};

// forward transform of one torus / integer polynomial through the group code path (group 0)
static void emu_forward(const uint64_t *poly, cd *out) {
    static cd tw[256];
    static bool init = false;
    if (!init) { make_twiddle_table(tw); init = true; }
    cd xb[XB_ELEMS];
    cd v[16][16];
    for (int lane = 0; lane < 16; lane++) { load_torus_poly(v[lane], lane, poly); fft256_fwd_pass1(v[lane], lane, tw, xb); }
    for (int lane = 0; lane < 16; lane++) fft256_fwd_pass2(v[lane], lane, xb);
    for (int lane = 0; lane < 16; lane++)
        for (int k2 = 0; k2 < 16; k2++) out[lane + 16 * k2] = v[lane][rev4(k2)];
}
// standard GGSW [level][row][col][N] -> kernel Fourier layout [level][row][col][p]
template <int K>
static void emu_convert_ggsw(const uint64_t *ggsw_std, int levels, cd *out) {
    cd tmp[POLY_M];
    for (int l = 0; l < levels; l++)
        for (int r = 0; r <= K; r++)
            for (int c = 0; c <= K; c++) {
                emu_forward(ggsw_std + (((size_t)l * (K + 1) + r) * (K + 1) + c) * POLY_N, tmp);
                for (int p = 0; p < POLY_M; p++) out[(((size_t)(levels - 1 - l) * (K + 1) + r) * (K + 1) + c) * POLY_M + p] = tmp[p];
            }
}

extern "C" void emu_fft_forward_torus(const uint64_t *poly, double *out) { emu_forward(poly, (cd *)out); }

extern "C" void emu_fft_roundtrip(const uint64_t *poly, uint64_t *out) {
    // forward as torus then the inverse path of the CMux (adds into out)
    static EmuCta<1, 1> *cta = new EmuCta<1, 1>();
    cd f[POLY_M];
    emu_forward(poly, f);
    memset(cta->sm.acc, 0, sizeof(cta->sm.acc));
    for (int tid = 0; tid < CMUX_THREADS; tid++) { cta->rg[tid].facc[0][0] = f[tid]; cta->rg[tid].facc[0][1] = cmk(0, 0); }
    for (int tid = 0; tid < CMUX_THREADS; tid++) phase_inv0<1, 1>(tid, cta->sm, cta->rg[tid]);
    for (int tid = 0; tid < CMUX_THREADS; tid++) phase_inv1<1, 1>(tid, cta->sm, cta->rg[tid]);
    for (int tid = 0; tid < CMUX_THREADS; tid++) phase_inv2<1, 1>(tid, cta->sm, cta->rg[tid]);
    for (int tid = 0; tid < CMUX_THREADS; tid++) phase_inv3<1, 1>(tid, cta->sm, cta->rg[tid]);
    memcpy(out, cta->sm.acc[0][0], POLY_N * 8);
}

// acc[g] += ggsw (x) (acc[g] * X^rot[g] - acc[g])   for g < G, through the emulated CTA
template <int K, int G, int BASE_LOG, int LEVELS>
static void emu_cmux_rotate_t(const uint64_t *ggsw_std, const int *rot, uint64_t *acc) {
    EmuCta<K, G> *cta = new EmuCta<K, G>();
    std::vector<cd> gf((size_t)LEVELS * (K + 1) * POLY_M * (K + 1));
    emu_convert_ggsw<K>(ggsw_std, LEVELS, gf.data());
    memcpy(cta->sm.acc, acc, sizeof(cta->sm.acc));
    for (int g = 0; g < G; g++) cta->sm.rot[g] = rot[g];
    cta->template cmux_step<BASE_LOG, LEVELS, DIFF_ROTATE>(gf.data(), nullptr);
    memcpy(acc, cta->sm.acc, sizeof(cta->sm.acc));
    delete cta;
}
extern "C" int emu_cmux_rotate(int K, int G, int base_log, int levels, const uint64_t *ggsw_std, const int *rot, uint64_t *acc) {
#define CASE(k, g, b, l) if (K == k && G == g && base_log == b && levels == l) { emu_cmux_rotate_t<k, g, b, l>(ggsw_std, rot, acc); return 0; }
    CASE(1, 1, 8, 5) CASE(1, 4, 8, 5) CASE(4, 3, 8, 5) CASE(1, 2, 15, 1) CASE(4, 3, 15, 1) CASE(4, 1, 8, 5)
    return -1;
}
