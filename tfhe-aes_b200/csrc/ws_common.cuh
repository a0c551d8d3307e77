// ws_common.cuh — helpers shared by the warp-specialised PBS kernels (pbs_ws_kernel.cu, pbs_ws1_kernel.cu):
// mbarrier / TMA bulk-copy wrappers and the modulus switch of the bootstrap input.
#pragma once
#include "cmux_core.cuh"
#include "kernels.h"

#define WS_THREADS 512
#define WS_FFT_THREADS 256
#ifndef WS_MAC_REGS3
#define WS_MAC_REGS3 112
#endif
#define WS_MAC_WARPS ((WS_THREADS - WS_FFT_THREADS) / 32)

__device__ __forceinline__ unsigned ws_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(ws_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(ws_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(ws_smem_u32(bar)), "r"(parity) : "memory");
}
// wait for two barriers with one polling loop (the two try_wait latencies overlap)
__device__ __forceinline__ void ws_mbar_wait2(uint64_t *bar_a, unsigned parity_a, uint64_t *bar_b, unsigned parity_b) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "WAIT2_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%2], %3;\n\t"
        "and.pred p, p, q;\n\t"
        "@p bra.uni WAIT2_DONE;\n\t"
        "bra.uni WAIT2_LOOP;\n\t"
        "WAIT2_DONE:\n\t}" ::"r"(ws_smem_u32(bar_a)), "r"(parity_a), "r"(ws_smem_u32(bar_b)), "r"(parity_b) : "memory");
}
__device__ __forceinline__ void ws_bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ws_smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(ws_smem_u32(bar))
                 : "memory");
}

// per-lane wait: lanes of one warp may wait on different barriers (the loop diverges and reconverges)
__device__ __forceinline__ void ws_mbar_wait_lane(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITL_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAITL_DONE;\n\t"
        "bra WAITL_LOOP;\n\t"
        "WAITL_DONE:\n\t}" ::"r"(ws_smem_u32(bar)), "r"(parity) : "memory");
}

// ---- tensor memory as a parking lot for per-thread state (32x32b shape: thread t of a warp owns TMEM lane 32 * (warp % 4) + t) ----
__device__ __forceinline__ void ws_tmem_st8(unsigned taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void ws_tmem_ld8(unsigned taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void ws_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// byte b of four words -> one word (digits of one decomposition level of four coefficients)
__device__ __forceinline__ uint32_t ws_gather_byte(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int b) {
    const uint32_t lo = __byte_perm(w0, w1, 0x0040 + b * 0x11);   // bytes: w0[b], w1[b]
    const uint32_t hi = __byte_perm(w2, w3, 0x0040 + b * 0x11);
    return __byte_perm(lo, hi, 0x5410);
}

// modulus switch to 2N: a~ = (a * in_scale [+ pre_add on the body] + 2^53) >> 54
__device__ __forceinline__ int ws_mod_switch_2n(const PbsArgs &a, int ct, int i) {
    uint64_t x = a.lwe_in[(size_t)ct * (a.lwe_dim + 1) + i] * a.in_scale;
    if (i == a.lwe_dim) x += a.pre_add_body;
    return (int)((x + (1ull << 53)) >> 54) & (2 * POLY_N - 1);
}

