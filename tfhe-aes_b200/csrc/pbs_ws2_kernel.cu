// pbs_ws2_kernel.cu — warp-specialised PBS with TWO ciphertext sets per CTA taking turns (sm_100a).
//
// Same arithmetic and the same two roles as pbs_ws_kernel.cu (FFT warps 8-15, MAC warps 0-7, key rows streamed by TMA into a ring of
// one level), but the CTA holds NS = 2 sets of G ciphertexts and every role works on the sets alternately, a whole CMux step at a
// time:
//     FFT role:  inverse A(i-1), rotate + decompose A(i), forward A(i) levels 5..1, inverse B(i-1), rotate + decompose B(i), ...
//     MAC role:                                   rows of A(i) levels 5..1, hand over A(i), rows of B(i) ...
// In pbs_ws_kernel the FFT role waits at the end of every step until the MAC role has consumed the last level and handed the
// Fourier accumulators over, and the MAC role waits while the FFT role runs inverse transform, decomposition and the first forward
// level (2 400 + 10 200 of 29 000 cycles per step).  Here the other set's work fills both: when the FFT role comes
// back to set A its hand-over was completed a third of a step ago, so the FFT role, which carries 2/3 of the FP64 work, never waits.
// The key stream delivers the 25 rows of every step twice (once per set): L2 -> SM traffic per ciphertext is unchanged.
//
// What makes the second set fit: the accumulators (G x 20 KB per set) live in TENSOR MEMORY, not in shared memory.  Each FFT lane
// owns the same 32 coefficients of its polynomial for the whole bootstrap (TMEM is lane-private: 64 columns per set and thread, next
// to the 32 columns of parked digits), adds the rounded inverse transform to them there (tcgen05.ld / tcgen05.st), and drops a copy
// into its own hand-over slot, which is idle between the inverse transform and the first forward level; the rotation X^a gathers
// from that copy.  Shared memory: 2 x 60 KB slots + 100 KB key ring + 4 KB twiddles.
#include "ws_common.cuh"

#ifndef PBS_WS2_NS
#define PBS_WS2_NS ws2_default
#endif
#ifndef PBS_WS2_LAUNCH_NAME
#define PBS_WS2_LAUNCH_NAME launch_pbs_ws2
#endif
#ifndef WS2_MAC_REGS
#define WS2_MAC_REGS 112
#endif
#ifndef WS2_DIAG
#define WS2_DIAG 0        // scratch/pbs_lab diagnostics (results WRONG): 1 = the roles never wait for each other's spectra / slots / hand-over
#endif
#ifndef WS2_TIMING
#define WS2_TIMING 0      // scratch/pbs_lab: per-activity clock64() sums of CTA 0 (FFT thread 0, MAC thread 0) into PbsArgs::dbg[0..9)
#endif
namespace PBS_WS2_NS {

template <int K, int G, int NS>
struct Ws2Smem {
    cd hs[NS][(K + 1) * G][XB_ELEMS];    // hand-over slots of set s: spectrum / Fourier accumulator / accumulator copy of (ct, r) at [r * G + ct]
    cd tw[256];                          // mid twiddles, swizzled (fft_core.cuh)
    cd ring[K + 1][K + 1][POLY_M];       // one level of the Fourier bootstrap key: [row][col][p]
    uint64_t rfull[NS][K + 1];
    uint64_t rempty[NS][K + 1];
    uint64_t inv[NS];
    uint64_t bfull[2];                   // key rows 0..BSPLIT-1 / BSPLIT..K of a level have landed
    uint64_t bempty[K + 1];
    uint32_t tmem_base;
    uint32_t pad_;
};

__device__ __forceinline__ void tmem_ld16(unsigned taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                 "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}

template <int G, int KP1>
__device__ __forceinline__ void cmac_cols2(cd (&facc)[G][KP1], int c, const cd (&x)[G], cd w) {
#pragma unroll
    for (int g = 0; g < G; g++) facc[g][c].x = fma(x[g].x, w.x, facc[g][c].x);
#pragma unroll
    for (int g = 0; g < G; g++) facc[g][c].y = fma(x[g].x, w.y, facc[g][c].y);
#pragma unroll
    for (int g = 0; g < G; g++) facc[g][c].x = fma(-x[g].y, w.y, facc[g][c].x);
#pragma unroll
    for (int g = 0; g < G; g++) facc[g][c].y = fma(x[g].y, w.x, facc[g][c].y);
}

// TMEM columns per FFT thread (two FFT warps share a lane quarter, 256 columns each): [0, 32) parked digits, [32 + 64 s, 96 + 64 s)
// the 32 accumulator coefficients of set s: coefficient 16 n1 + lane at columns 4 n1 (lo), 4 n1 + 1 (hi), + 256 at 4 n1 + 2, 4 n1 + 3
constexpr int TM_DIGITS = 0, TM_ACC = 32, TM_PER_WARP = 256;

template <int K, int G, int NS, int BASE_LOG, int LEVELS>
__global__ void __launch_bounds__(WS_THREADS, 1) pbs_ws2_kernel(PbsArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Ws2Smem<K, G, NS> &sm = *reinterpret_cast<Ws2Smem<K, G, NS> *>(smem_raw);
    const int tid = threadIdx.x;
    const int n = a.lwe_dim;
    const int ct0 = blockIdx.x * (G * NS);
    constexpr int RING = K + 1;
    constexpr int ROWS = LEVELS * (K + 1);
    constexpr int ROW_ELEMS = POLY_M * (K + 1);
    constexpr unsigned ROW_BYTES = (unsigned)(ROW_ELEMS * sizeof(cd));
    const int nrows = n * ROWS * NS;           // rows of the key stream: the ROWS rows of every step NS times
    constexpr int BSPLIT = (K + 2) / 2;
    constexpr int MAC_REGS = WS2_MAC_REGS;
    static_assert(LEVELS == 5 && BASE_LOG == 8, "digit parking packs five 8-bit levels");
    static_assert(TM_ACC + 64 * NS <= TM_PER_WARP, "tensor-memory columns");
    long long tacc[6] = {0, 0, 0, 0, 0, 0};
    long long tlast = WS2_TIMING ? clock64() : 0;
#define WT(k) do { if (WS2_TIMING) { const long long t_ = clock64(); tacc[(k) % 6] += t_ - tlast; tlast = t_; } } while (0)

    // ---- prologue (all 512 threads) ---------------------------------------------------------------
    for (int i = tid; i < 256; i += WS_THREADS) sm.tw[i] = a.tw[i];
    if (tid == 0) {
        for (int s = 0; s < NS; s++) {
            for (int r = 0; r <= K; r++) {
                ws_mbar_init(&sm.rfull[s][r], G * ((r | 1) <= K ? 2 : 1));   // shared by the row pairs (0,1), (2,3), (4)
                ws_mbar_init(&sm.rempty[s][r], WS_MAC_WARPS);
            }
            ws_mbar_init(&sm.inv[s], WS_MAC_WARPS);
        }
        for (int r = 0; r <= K; r++) ws_mbar_init(&sm.bempty[r], WS_MAC_WARPS);
        ws_mbar_init(&sm.bfull[0], BSPLIT);
        ws_mbar_init(&sm.bfull[1], K + 1 - BSPLIT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {   // all 512 columns: the CTA owns the SM (229 KB of shared memory)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(ws_smem_u32(&sm.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    if (tid >= WS_THREADS - WS_FFT_THREADS) {
        // ================================ FFT warps ================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(256 - MAC_REGS));
        const int ftid = tid - (WS_THREADS - WS_FFT_THREADS);
        const int gid = 15 - (ftid >> 4), lane = ftid & 15;        // groups in reverse warp order (the arbiter favours high warp ids)
        const bool active = gid < G * (K + 1);
        const int ct = active ? gid % G : 0, r = active ? gid / G : 0;
        const int gid_a = 14 - (ftid >> 5) * 2, gid_b = gid_a + 1;
        const int r_a = gid_a / G;
        const int r_b = (gid_b < G * (K + 1)) ? gid_b / G : r_a;
        (void)r_a;
        cd v[16];
        uint32_t st_re[16], st_im[16];
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned taddr = sm.tmem_base + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((ftid >> 7) * TM_PER_WARP);
        if (gid_a < G * (K + 1)) {
            // initial accumulators (0, ..., 0, lut * X^-b~): into tensor memory and, as the copy the first rotation reads, into the slot
#pragma unroll 1
            for (int s = 0; s < NS; s++) {
                const int my_ct = min(ct0 + s * G + ct, a.count - 1);
                const int rot = (2 * POLY_N - ws_mod_switch_2n(a, my_ct, n)) & (2 * POLY_N - 1);
                uint64_t *copy = reinterpret_cast<uint64_t *>(sm.hs[s][active ? gid : 0]);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t w[16];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int j = 16 * (4 * c + k) + lane;
                        const uint64_t a0 = (r == K) ? rotated_coef(a.lut, j, rot) : 0;
                        const uint64_t a1 = (r == K) ? rotated_coef(a.lut, j + POLY_M, rot) : 0;
                        w[4 * k] = (uint32_t)a0; w[4 * k + 1] = (uint32_t)(a0 >> 32);
                        w[4 * k + 2] = (uint32_t)a1; w[4 * k + 3] = (uint32_t)(a1 >> 32);
                        if (active) { copy[j] = a0; copy[j + POLY_M] = a1; }
                    }
                    tmem_st16(taddr + TM_ACC + 64 * s + 16 * c, w);
                }
            }
            ws_tmem_wait_st();
            __syncwarp();
#pragma unroll 1
            for (int i = 0; i <= n; i++) {
#pragma unroll 1
                for (int s = 0; s < NS; s++) {
                    cd *slot = sm.hs[s][active ? gid : 0];
                    const int my_ct = min(ct0 + s * G + ct, a.count - 1);
                    // raw mask element of step i: fetched before the inverse transform, mod-switched after it
                    const uint64_t raw = a.lwe_in[(size_t)my_ct * (n + 1) + min(i, n - 1)];
                    if (i > 0) {
                        // ---- finish step i-1: inverse transform of the Fourier accumulator the MAC warps left in this group's slot
                        WT(3);
                        if (!(WS2_DIAG & 1)) ws_mbar_wait(&sm.inv[s], (i - 1) & 1);
                        WT(4);
#pragma unroll
                        for (int k2 = 0; k2 < 16; k2++) v[k2] = slot[lane + 16 * k2];
                        fft256_inv_pass1_compute(v);
                        __syncwarp();
                        if (active) fft256_inv_pass1_store_b<8>(v, lane, sm.tw, slot);
                        __syncwarp();
                        fft256_inv_pass2(v, lane, slot);
                        __syncwarp();       // every lane has read the slot: it now takes the accumulator copy
                        uint64_t *copy = reinterpret_cast<uint64_t *>(slot);
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            uint32_t w[16];
                            tmem_ld16(taddr + TM_ACC + 64 * s + 16 * c, w);
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const int n1 = 4 * c + k, j = 16 * n1 + lane;
                                const uint64_t a0 = (((uint64_t)w[4 * k + 1] << 32) | w[4 * k]) + f64_to_torus(v[n1].x);
                                const uint64_t a1 = (((uint64_t)w[4 * k + 3] << 32) | w[4 * k + 2]) + f64_to_torus(v[n1].y);
                                w[4 * k] = (uint32_t)a0; w[4 * k + 1] = (uint32_t)(a0 >> 32);
                                w[4 * k + 2] = (uint32_t)a1; w[4 * k + 3] = (uint32_t)(a1 >> 32);
                                if (active) { copy[j] = a0; copy[j + POLY_M] = a1; }
                            }
                            tmem_st16(taddr + TM_ACC + 64 * s + 16 * c, w);
                        }
                        ws_tmem_wait_st();
                        __syncwarp();
                        WT(5);
                    }
                    if (i < n) {
                        // ---- start step i: rotate-subtract from the copy, decompose, park the digits of levels 4..1
                        const int rot = (int)((raw * a.in_scale + (1ull << 53)) >> 54) & (2 * POLY_N - 1);
                        load_decompose_rot<BASE_LOG, LEVELS>(reinterpret_cast<const uint64_t *>(slot), lane, rot, v, st_re, st_im);
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            uint32_t pk[8];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                pk[j] = ws_gather_byte(st_re[4 * j], st_re[4 * j + 1], st_re[4 * j + 2], st_re[4 * j + 3], b);
                                pk[4 + j] = ws_gather_byte(st_im[4 * j], st_im[4 * j + 1], st_im[4 * j + 2], st_im[4 * j + 3], b);
                            }
                            ws_tmem_st8(taddr + TM_DIGITS + b * 8, pk);
                        }
                        ws_tmem_wait_st();
                        __syncwarp();       // the copy has been gathered by every lane before the first spectrum overwrites it
                        WT(0);
#pragma unroll 1
                        for (int lev = LEVELS; lev >= 1; lev--) {
                            const unsigned produced = (unsigned)(i * LEVELS + (LEVELS - lev));   // levels of this set produced so far
                            if (lev != LEVELS) {
                                uint32_t pk[8];
                                ws_tmem_ld8(taddr + TM_DIGITS + (4 - lev) * 8, pk);
#pragma unroll
                                for (int n1 = 0; n1 < 16; n1++) v[n1] = cmk(digit85(pk[n1 >> 2], n1 & 3), digit85(pk[4 + (n1 >> 2)], n1 & 3));
                            }
                            fft256_fwd_pass1_compute(v, lane, sm.tw);
                            WT(1);
                            // the MAC warps have consumed the previous occupant of the slot (they release the rows of a level in order)
                            if (produced > 0 && !(WS2_DIAG & 1)) ws_mbar_wait(&sm.rempty[s][r_b], (produced - 1) & 1);
                            WT(2);
                            if (active) fft256_fwd_pass1_store(v, lane, slot);
                            __syncwarp();
                            fft256_fwd_pass2(v, lane, slot);
                            __syncwarp();
                            if (active) {
#pragma unroll
                                for (int k2 = 0; k2 < 16; k2++) slot[lane + 16 * k2] = v[rev4(k2)];
                            }
                            __syncwarp();
                            if (active && lane == 0) ws_mbar_arrive(&sm.rfull[s][r & ~1]);
                            WT(3);
                        }
                    }
                }
            }
            if (WS2_TIMING && a.dbg && blockIdx.x == 0 && ftid == 0)
                for (int k = 0; k < 6; k++) a.dbg[k] = (uint64_t)tacc[k];
        }
    } else {
        // ================================ MAC warps ================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MAC_REGS));
        const int p = tid;
        const int mwarp = p >> 5, mlane = p & 31;
        auto produce = [&](int q) {               // fetch row q of the stream into slot q % RING (the caller knows it is free)
            const int sl = q % RING;
            uint64_t *bar = &sm.bfull[sl < BSPLIT ? 0 : 1];
            const size_t src_row = (size_t)(q / (ROWS * NS)) * ROWS + (size_t)(q % ROWS);
            ws_mbar_arrive_expect_tx(bar, ROW_BYTES);
            ws_bulk_copy_g2s(&sm.ring[sl][0][0], a.bsk + src_row * ROW_ELEMS, ROW_BYTES, bar);
        };
        if (p == 0)
            for (int q = 0; q < RING; q++) produce(q);
        cd facc[G][K + 1];
        unsigned level_count = 0;                 // levels of the key stream consumed so far
        int q = 0;                                // next key row to consume
#pragma unroll 1
        for (int i = 0; i < n; i++) {
#pragma unroll 1
            for (int s = 0; s < NS; s++) {
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int c = 0; c <= K; c++) facc[g][c] = cmk(0.0, 0.0);
#pragma unroll 1
                for (int lev = LEVELS; lev >= 1; lev--) {
                    const unsigned parity = level_count & 1;                              // key ring
                    const unsigned pr = (unsigned)(i * LEVELS + (LEVELS - lev)) & 1;      // spectra of this set
#pragma unroll
                    for (int r = 0; r <= K; r++, q++) {
                        if (WS2_DIAG & 1) {
                            if (r == 0) ws_mbar_wait(&sm.bfull[0], parity);
                            else if (r == BSPLIT) ws_mbar_wait(&sm.bfull[1], parity);
                        } else if (r == 0) ws_mbar_wait2(&sm.bfull[0], parity, &sm.rfull[s][0], pr);
                        else if (r == BSPLIT && (r & 1) == 0) ws_mbar_wait2(&sm.bfull[1], parity, &sm.rfull[s][r], pr);
                        else if (r == BSPLIT) ws_mbar_wait(&sm.bfull[1], parity);
                        else if ((r & 1) == 0) ws_mbar_wait(&sm.rfull[s][r], pr);
                        WT(0);
                        cd x[G];
#pragma unroll
                        for (int g = 0; g < G; g++) x[g] = sm.hs[s][r * G + g][p];
#pragma unroll
                        for (int c = 0; c <= K; c++) {
                            const cd w = sm.ring[r][c][p];
                            cmac_cols2<G, K + 1>(facc, c, x, w);
                        }
                        __syncwarp();
                        if (mlane == 0) {
                            ws_mbar_arrive(&sm.bempty[r]);
                            if (lev > 1) ws_mbar_arrive(&sm.rempty[s][r]);   // after the last level the slot is released below
                            if (mwarp == (q & (WS_MAC_WARPS - 1)) && q >= 1 && q - 1 + RING < nrows) {
                                const int ps = (r + K) % RING;                              // slot of row q - 1
                                const unsigned pp = (r == 0) ? (parity ^ 1) : parity;       // its level
                                ws_mbar_wait(&sm.bempty[ps], pp);
                                produce(q - 1 + RING);
                            }
                        }
                        WT(1);
                    }
                    level_count++;
                }
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int c = 0; c <= K; c++) sm.hs[s][c * G + g][p] = facc[g][c];
                __syncwarp();
                if (mlane == 0) {
                    ws_mbar_arrive(&sm.inv[s]);
#pragma unroll
                    for (int r = 0; r <= K; r++) ws_mbar_arrive(&sm.rempty[s][r]);
                }
                WT(2);
            }
        }
        if (WS2_TIMING && a.dbg && blockIdx.x == 0 && p == 0)
            for (int k = 0; k < 3; k++) a.dbg[6 + k] = (uint64_t)tacc[k];
    }
#undef WT
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(sm.tmem_base) : "memory");
    // sample extract of coefficient 0 from the accumulator copies the last inverse transform left in the slots
    for (int s = 0; s < NS; s++)
        for (int g = 0; g < G; g++) {
            const int c = ct0 + s * G + g;
            if (c >= a.count) continue;
            uint64_t *out = a.out + (size_t)c * (K * POLY_N + 1);
            for (int idx = tid; idx < K * POLY_N; idx += WS_THREADS) {
                const int r = idx / POLY_N, j = idx % POLY_N;
                const uint64_t *poly = reinterpret_cast<const uint64_t *>(sm.hs[s][r * G + g]);
                out[idx] = (j == 0) ? poly[0] : (uint64_t)0 - poly[POLY_N - j];
            }
            if (tid == 0) out[K * POLY_N] = reinterpret_cast<const uint64_t *>(sm.hs[s][K * G + g])[0] + a.post_add;
        }
}

}  // namespace PBS_WS2_NS
using namespace PBS_WS2_NS;

// G ciphertexts per set, two sets per CTA; only the PARAM_OPT shape is instantiated (the caller falls back to pbs_ws_kernel otherwise)
cudaError_t PBS_WS2_LAUNCH_NAME(int K, int G, int base_log, int levels, const PbsArgs &a, cudaStream_t s) {
    if (K == 4 && G == 3 && base_log == 8 && levels == 5 && (WS2_TIMING || !a.dbg)) {
        const size_t smem = sizeof(Ws2Smem<4, 3, 2>);
        cudaError_t e = cudaFuncSetAttribute(pbs_ws2_kernel<4, 3, 2, 8, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        pbs_ws2_kernel<4, 3, 2, 8, 5><<<(a.count + 5) / 6, WS_THREADS, smem, s>>>(a);
        return cudaGetLastError();
    }
    return cudaErrorInvalidValue;
}
