// pbs_ws_kernel.cu — warp-specialised PBS (blind rotation + sample extract), sm_100a.
//
// Same arithmetic as pbs_kernel (fp_kernels.cu / cmux_core.cuh), different schedule.  One CTA of 512
// threads holds G ciphertexts for all n CMux steps:
//   warps 0-7  "FFT warps" : per step: rotate-subtract + signed decomposition, per level a forward FFT of
//                            every digit polynomial into its hand-over slot, then the K+1 inverse FFTs and the
//                            accumulator update.  Each 16-lane group owns polynomial (ct, r) and only ever
//                            touches its own accumulator polynomial and its own slot, so FFT groups never wait
//                            for each other (warp-level syncs only).
//   warps 8-15 "MAC warps" : thread p owns Fourier point p of all G ciphertexts and multiplies row r of the
//                            current level as soon as the G slots (*, r) are full, while the FFT warps already
//                            work on the next level.  The Fourier bootstrap key is streamed L2 -> shared memory
//                            by a thread-private cp.async ring (BSK_RING rows ahead, all CTAs in step).
// The two roles have different register needs (FFT: 16 complex + 32 decomposition states; MAC: G*(K+1)
// complex accumulators), so the register file is re-balanced with setmaxnreg (152 / 104 per thread):
// four warps per scheduler instead of two, and FP64 (FFT, MAC), ALU (decomposition) and LSU phases of the
// two roles overlap instead of alternating.
// Hand-shake per row r (mbarriers in shared memory): FULL[r] (G*16 FFT lanes arrive, MAC threads wait),
// EMPTY[r] (256 MAC threads arrive, the FFT lanes of row r wait), INV (MAC threads arrive after leaving the
// Fourier accumulators in the slots, FFT lanes wait).
#include "cmux_core.cuh"
#include "kernels.h"

#define WS_THREADS 512
#define WS_FFT_THREADS 256
#define WS_RING 4

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(addr), "r"(parity) : "memory");
}

template <int K, int G>
struct WsSmem {
    uint64_t acc[G][K + 1][POLY_N];
    cd xb[CMUX_GROUPS][XB_ELEMS];
    cd tw[256];
    uint64_t full[K + 1];
    uint64_t empty[K + 1];
    uint64_t inv;
    uint64_t pad_;
};

template <int K, int G, int BASE_LOG, int LEVELS>
__global__ void __launch_bounds__(WS_THREADS, 1) pbs_ws_kernel(PbsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WsSmem<K, G> &sm = *reinterpret_cast<WsSmem<K, G> *>(smem_raw);
    cd *ring = reinterpret_cast<cd *>(smem_raw + sizeof(WsSmem<K, G>));              // [WS_RING][K+1][256]
    uint16_t *ahat = reinterpret_cast<uint16_t *>(ring + WS_RING * POLY_M * (K + 1));
    const int tid = threadIdx.x;
    const int n = a.lwe_dim, np = a.lwe_dim + 1;
    const int ct0 = blockIdx.x * G;
    constexpr int ROWS = LEVELS * (K + 1);
    constexpr size_t ROW_ELEMS = (size_t)POLY_M * (K + 1);
    constexpr unsigned ROW_BYTES = (unsigned)(ROW_ELEMS * sizeof(cd));
    constexpr unsigned RING_BYTES = WS_RING * ROW_BYTES;

    // ---- prologue (all 512 threads) ---------------------------------------------------------------
    for (int i = tid; i < 256; i += WS_THREADS) sm.tw[i] = a.tw[i];
    for (int idx = tid; idx < G * np; idx += WS_THREADS) {
        const int g = idx / np, i = idx % np;
        const int ct = min(ct0 + g, a.count - 1);
        uint64_t x = a.lwe_in[(size_t)ct * np + i] * a.in_scale;
        if (i == n) x += a.pre_add_body;
        ahat[g * np + i] = (uint16_t)((x + (1ull << 53)) >> 54);  // modulus switch to 2N
    }
    if (tid == 0) {
        for (int r = 0; r <= K; r++) { mbar_init(&sm.full[r], G * 16); mbar_init(&sm.empty[r], WS_THREADS - WS_FFT_THREADS); }
        mbar_init(&sm.inv, WS_THREADS - WS_FFT_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int g = 0; g < G; g++) {
        const int rot = (2 * POLY_N - ahat[g * np + n]) & (2 * POLY_N - 1);
        for (int idx = tid; idx < (K + 1) * POLY_N; idx += WS_THREADS) {
            const int r = idx / POLY_N, j = idx % POLY_N;
            sm.acc[g][r][j] = (r == K) ? rotated_coef(a.lut, j, rot) : 0;
        }
    }
    __syncthreads();

    if (tid < WS_FFT_THREADS) {
        // ================================ FFT warps ================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int gid = tid >> 4, lane = tid & 15;
        const bool active = gid < G * (K + 1);
        const int ct = active ? gid / (K + 1) : 0, r = active ? gid % (K + 1) : 0;
        // Both 16-lane groups of a warp execute ONE instruction stream (a diverged half-warp would pay a full
        // issue slot and a full FP64 pipe pass for 16 lanes): waits are made warp-uniform by waiting for the
        // rows of both groups; an idle group (gid >= G*(K+1)) runs along on its own scratch slot.
        const int gid_a = (tid >> 5) * 2, gid_b = gid_a + 1;
        const int r_a = gid_a % (K + 1);
        const int r_b = (gid_b < G * (K + 1)) ? gid_b % (K + 1) : r_a;
        cd v[16];
        uint32_t st_re[16], st_im[16];
        unsigned produced = 0;                    // productions into this group's slot so far
        if (gid_a < G * (K + 1)) {
#pragma unroll 1
            for (int i = 0; i < n; i++) {
                load_decompose_rot<BASE_LOG, LEVELS>(sm.acc[ct][r], lane, ahat[ct * np + i], v, st_re, st_im);
#pragma unroll 1
                for (int lev = LEVELS; lev >= 1; lev--) {
                    if (lev != LEVELS) next_digits<BASE_LOG, LEVELS>(v, st_re, st_im, lev);
                    // the slots must have been consumed by the MAC warps (rows of the previous production)
                    if (produced > 0) {
                        mbar_wait(&sm.empty[r_a], (produced - 1) & 1);
                        if (r_b != r_a) mbar_wait(&sm.empty[r_b], (produced - 1) & 1);
                    }
                    __syncwarp();
                    fft256_fwd_pass1(v, lane, sm.tw, sm.xb[gid]);
                    __syncwarp();
                    fft256_fwd_pass2(v, lane, sm.xb[gid]);
                    __syncwarp();
#pragma unroll
                    for (int k2 = 0; k2 < 16; k2++) sm.xb[gid][lane + 16 * k2] = v[rev4(k2)];
                    if (active) mbar_arrive(&sm.full[r]);
                    produced++;
                }
                // inverse transform of the Fourier accumulator the MAC warps left in this group's slot
                mbar_wait(&sm.inv, i & 1);
                __syncwarp();
#pragma unroll
                for (int k2 = 0; k2 < 16; k2++) v[k2] = sm.xb[gid][lane + 16 * k2];
                fft256_inv_pass1_compute(v);
                __syncwarp();
                fft256_inv_pass1_store(v, lane, sm.tw, sm.xb[gid]);
                __syncwarp();
                fft256_inv_pass2(v, lane, sm.xb[gid]);
                if (active) {
                    uint64_t *poly = sm.acc[ct][r];
#pragma unroll
                    for (int n1 = 0; n1 < 16; n1++) {
                        const int j = 16 * n1 + lane;
                        poly[j] += f64_to_torus(v[n1].x);
                        poly[j + POLY_M] += f64_to_torus(v[n1].y);
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ================================ MAC warps ================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        const int p = tid - WS_FFT_THREADS;
        // bootstrap-key ring: thread-private slots, filled WS_RING rows ahead (key stored in consumption order)
        const cd *pf_src = a.bsk + p;
        long pf_left = (long)n * ROWS;
        const unsigned ring_u32 = (unsigned)__cvta_generic_to_shared(ring) + p * (unsigned)sizeof(cd);
        unsigned pf_off = 0, rd_off = 0;
        auto issue = [&]() {
            if (pf_left > 0) {
#pragma unroll
                for (int c = 0; c <= K; c++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(ring_u32 + pf_off + c * (POLY_M * (unsigned)sizeof(cd))),
                                 "l"(pf_src + c * POLY_M));
                pf_src += ROW_ELEMS;
                pf_left--;
                pf_off += ROW_BYTES;
                if (pf_off == RING_BYTES) pf_off = 0;
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
#pragma unroll
        for (int s = 0; s < WS_RING; s++) issue();
        cd facc[G][K + 1];
        unsigned level_count = 0;
#pragma unroll 1
        for (int i = 0; i < n; i++) {
#pragma unroll
            for (int g = 0; g < G; g++)
#pragma unroll
                for (int c = 0; c <= K; c++) facc[g][c] = cmk(0.0, 0.0);
#pragma unroll 1
            for (int lev = LEVELS; lev >= 1; lev--) {
                const unsigned parity = level_count & 1;
#pragma unroll
                for (int r = 0; r <= K; r++) {
                    asm volatile("cp.async.wait_group %0;\n" ::"n"(WS_RING - 1) : "memory");
                    const cd *w_ptr = reinterpret_cast<const cd *>(reinterpret_cast<const unsigned char *>(ring) + rd_off) + p;
                    mbar_wait(&sm.full[r], parity);
                    cd x[G];
#pragma unroll
                    for (int g = 0; g < G; g++) x[g] = sm.xb[g * (K + 1) + r][p];
#pragma unroll
                    for (int c = 0; c <= K; c++) {   // key values are streamed one at a time (register budget)
                        const cd w = w_ptr[c * POLY_M];
#pragma unroll
                        for (int g = 0; g < G; g++) cmac(facc[g][c], x[g], w);
                    }
                    if (lev > 1) mbar_arrive(&sm.empty[r]);   // after the last level the slot is released below
                    rd_off += ROW_BYTES;
                    if (rd_off == RING_BYTES) rd_off = 0;
                    issue();
                }
                level_count++;
            }
            // hand over through the slots: thread p only ever reads and writes element p of a slot, so no
            // MAC-side synchronisation is needed; the FFT lanes read after INV completes
#pragma unroll
            for (int g = 0; g < G; g++)
#pragma unroll
                for (int c = 0; c <= K; c++) sm.xb[g * (K + 1) + c][p] = facc[g][c];
            mbar_arrive(&sm.inv);
#pragma unroll
            for (int r = 0; r <= K; r++) mbar_arrive(&sm.empty[r]);  // the slots are free again once the FFT lanes have read them (they wait on INV first)
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    // sample extract of coefficient 0
    for (int g = 0; g < G; g++) {
        if (ct0 + g >= a.count) continue;
        uint64_t *out = a.out + (size_t)(ct0 + g) * (K * POLY_N + 1);
        for (int idx = tid; idx < K * POLY_N; idx += WS_THREADS) {
            const int r = idx / POLY_N, j = idx % POLY_N;
            out[idx] = (j == 0) ? sm.acc[g][r][0] : (uint64_t)0 - sm.acc[g][r][POLY_N - j];
        }
        if (tid == 0) out[K * POLY_N] = sm.acc[g][K][0] + a.post_add;
    }
}

#define LAUNCH_WS(k, g, bl, lv)                                                                         \
    if (K == k && G == g && base_log == bl && levels == lv) {                                           \
        size_t smem = sizeof(WsSmem<k, g>) + (size_t)WS_RING * POLY_M * (k + 1) * sizeof(cd) +              \
                      (size_t)g * (a.lwe_dim + 1) * sizeof(uint16_t);                                  \
        cudaError_t e = cudaFuncSetAttribute(pbs_ws_kernel<k, g, bl, lv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                 \
        pbs_ws_kernel<k, g, bl, lv><<<(a.count + g - 1) / g, WS_THREADS, smem, s>>>(a);                 \
        return cudaGetLastError();                                                                      \
    }
cudaError_t launch_pbs_ws(int K, int G, int base_log, int levels, const PbsArgs &a, cudaStream_t s) {
    LAUNCH_WS(4, 1, 8, 5) LAUNCH_WS(1, 1, 8, 5)
    return cudaErrorInvalidValue;
}
