// pbs_ws2_kernel.cu — warp-specialised PBS with TWO ciphertext sets per CTA taking turns (sm_100a).
//
// Same arithmetic and the same two roles as pbs_ws_kernel.cu (FFT warps 8-15, MAC warps 0-7, key rows streamed by TMA into a ring of
// one level), but the CTA holds NS = 2 sets of G = 3 ciphertexts and each role works on the sets alternately, a CMux step at a time:
//     MAC role:  rows of A(i) levels 5..1, hand over A(i), rows of B(i) levels 5..1, hand over B(i), rows of A(i+1) ...
//     FFT role:  while the MAC role is on A(i): forward levels 4..1 of A(i), then inverse transform of B(i-1), rotation +
//                decomposition of B(i), forward level 5 of B(i); then the same with A and B exchanged.
// In pbs_ws_kernel the FFT role waits at the end of every step until the MAC role has consumed the last level and handed the
// Fourier accumulators over (2 900 cycles per step), and the MAC role has nothing to do while the FFT role runs inverse transform,
// decomposition and the first forward level (10 200).  Here the hand-over of a set is complete long before the FFT role comes back
// to it, and the MAC role has the last level of the other set to work on during the first 4 800 cycles of that stretch.
// The key stream delivers the 25 rows of every step twice (once per set): L2 -> SM traffic per ciphertext is unchanged.
// Measured (scratch/pbs_lab, 888 ciphertexts = one wave, bit-identical outputs): 18.92 ms against 2 x 10.13 ms.
//
// What makes the second set fit: the accumulators (G x 20 KB per set) live in TENSOR MEMORY, not in shared memory.  Each FFT lane
// owns the same 32 coefficients of its polynomial for the whole bootstrap (TMEM is lane-private: 64 columns per set and thread, next
// to 40 columns of parked digits per set), adds the rounded inverse transform to them there (tcgen05.ld / tcgen05.st), and drops a
// copy into its own hand-over slot, which is idle between the inverse transform and the first forward level; the rotation X^a
// gathers from that copy.  Shared memory: 2 x 60 KB slots + 100 KB key ring + 4 KB twiddles; tensor memory: all 512 columns.
// The FFT role is a table-driven loop over three self-contained activities (forward level, end of step, start of step): no
// register state crosses from one to the next, every digit level comes back from tensor memory, and the order is a macro.
#include "ws_common.cuh"

#ifndef PBS_WS2_NS
#define PBS_WS2_NS ws2_default
#endif
#ifndef PBS_WS2_LAUNCH_NAME
#define PBS_WS2_LAUNCH_NAME launch_pbs_ws2
#endif
// Tuning switches; the defaults are what scratch/pbs_lab measured best on B200 (profiles/r2_pbs_lab_ws2.txt), every variant bit-identical.
#ifndef WS2_MAC_REGS
#define WS2_MAC_REGS 120      // registers per MAC thread (FFT threads get 256 - this): 112 spills in the MAC role (20.23 ms), 128 in the FFT role (19.04)
#endif
#ifndef WS2_ORDER
#define WS2_ORDER 0x5761234   // program of the FFT role per half-period, first activity in the low nibble (see the loop in the kernel):
                              // levels 4..1 of X, then end of step / start of step / level 5 of Y.  Measured per 888 ciphertexts: this order
                              // 19.18 ms; Y's work slotted between X's levels (0x1527364, 0x5712364, 0x5716234, 0x5176234) 19.4-20.7 ms
#endif
#ifndef WS2_EARLY_RELEASE
#define WS2_EARLY_RELEASE 0   // 1: spectra and key slot released right after the first FMA pass (all loads have returned by then): 19.26 ms
#endif                        //    against 18.94 (the barrier traffic in the middle of the FMA block costs more than the earlier refill gains)
#ifndef WS2_NAMED_BAR
#define WS2_NAMED_BAR 0       // 1: "spectra of a row pair ready" through hardware barriers 1-6 (FFT warps bar.arrive, MAC warps bar.sync) instead of
#endif                        //    polled mbarriers: no polling instructions, but 19.71 ms against 18.93 (bar.sync re-aligns the eight MAC warps)
#ifndef WS2_RFULL_PER_ROW
#define WS2_RFULL_PER_ROW 0   // 1: one spectra barrier per row instead of per row pair: 18.90 against 18.97 ms in the same run (noise)
#endif
#ifndef WS2_REVMAP
#define WS2_REVMAP 1          // FFT groups in reverse warp order: row 0, which the MAC role needs first, sits on the warp the issue arbiter favours
#endif                        //    (0: natural order, 20.07 ms against 18.96)
#ifndef WS2_MAC_WIDE
#define WS2_MAC_WIDE 1     // all K+1 key values of a row first, then four FMA passes over all (g, c): dependent FMAs 15 apart (19.17 -> 18.92 ms per 888)
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

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

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

// TMEM columns per FFT thread (two FFT warps share a lane quarter, 256 columns each): [64 s, 64 s + 40) the parked digits of set s
// (levels 4..1 at 8 (4 - lev), level 5 at 32), [128 + 64 s, 192 + 64 s) the 32 accumulator coefficients of set s: coefficient 16 n1 + lane at columns 4 n1 (lo), 4 n1 + 1 (hi), + 256 at 4 n1 + 2, 4 n1 + 3
constexpr int TM_DIGITS = 0, TM_ACC = 128, TM_PER_WARP = 256;

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
                ws_mbar_init(&sm.rfull[s][r], WS2_RFULL_PER_ROW ? G : G * ((r | 1) <= K ? 2 : 1));   // shared by the row pairs (0,1), (2,3), (4)
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
        const int gid = WS2_REVMAP ? 15 - (ftid >> 4) : (ftid >> 4), lane = ftid & 15;   // groups in reverse warp order (the arbiter favours high warp ids)
        const bool active = gid < G * (K + 1);
        const int ct = active ? gid % G : 0, r = active ? gid / G : 0;
        const int gid_a = WS2_REVMAP ? 14 - (ftid >> 5) * 2 : (ftid >> 5) * 2, gid_b = gid_a + 1;
        const int r_a = gid_a / G;
        const int r_b = (gid_b < G * (K + 1)) ? gid_b / G : r_a;
        (void)r_a;
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

            // ---- the three activities of the FFT role; each is self-contained (no register state crosses from one to the next)
            // forward transform of level lev of set s, step i: digits from tensor memory -> spectrum in the slot
            auto slot_of = [&](int s) {
                int off = s * (K + 1) * G * XB_ELEMS;
                asm volatile("" : "+r"(off));     // opaque: keeps per-set addresses from being precomputed (and spilled) outside the loop
                return &sm.hs[0][active ? gid : 0][0] + off;
            };
            auto forward_level = [&](int s, int lev, int i) {
                cd *slot = slot_of(s);
                cd v[16];
                {
                    uint32_t pk[8];
                    ws_tmem_ld8(taddr + TM_DIGITS + 64 * s + (lev == LEVELS ? 4 : 4 - lev) * 8, pk);
#pragma unroll
                    for (int n1 = 0; n1 < 16; n1++) v[n1] = cmk(digit85(pk[n1 >> 2], n1 & 3), digit85(pk[4 + (n1 >> 2)], n1 & 3));
                }
                fft256_fwd_pass1_compute(v, lane, sm.tw);
                WT(1);
                // the MAC warps have consumed the previous occupant of the slot (they release the rows of a level in order)
                const unsigned produced = (unsigned)(i * LEVELS + (LEVELS - lev));   // levels of this set produced so far
                if (produced > 0) ws_mbar_wait(&sm.rempty[s][r_b], (produced - 1) & 1);
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
#if WS2_NAMED_BAR
                // hardware barrier 1 + 3 s + pair: the FFT warps of the pair arrive, the eight MAC warps wait in bar.sync (asleep, no polling)
                named_arrive(1 + 3 * s + (r_a >> 1), WS_THREADS - WS_FFT_THREADS + 32 * ((r_a >> 1) == 2 ? 2 : 3));
#else
                if (active && lane == 0) ws_mbar_arrive(&sm.rfull[s][WS2_RFULL_PER_ROW ? r : (r & ~1)]);
#endif
                WT(3);
            };
            // end of step i of set s: inverse transform of the Fourier accumulator the MAC warps left in the slot, added to the
            // accumulator in tensor memory; the slot keeps a copy for the rotation of the next step
            auto finish_step = [&](int s, int i) {
                cd *slot = slot_of(s);
                cd v[16];
                ws_mbar_wait(&sm.inv[s], i & 1);
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
            };
            // start of step i of set s: (acc * X^a - acc) gathered from the copy, signed decomposition, all five digit levels parked
            // in tensor memory (one word per four coefficients and level)
            auto start_step = [&](int s, uint64_t raw) {
                const uint64_t *poly = reinterpret_cast<const uint64_t *>(slot_of(s));
                asm volatile("" : "+l"(raw)::"memory");   // the rotation arithmetic stays here (it was being hoisted to the fetch and spilled)
                const int rot = (int)((raw * a.in_scale + (1ull << 53)) >> 54) & (2 * POLY_N - 1);
                const int s0 = (lane - rot) & (2 * POLY_N - 1);
                uint32_t st_re[16], st_im[16], p5[8];
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++) {
                    uint32_t lo_re[4], lo_im[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int n1 = 4 * q4 + k, j = 16 * n1 + lane;
                        const int sh = (s0 + 16 * n1) & (2 * POLY_N - 1);
                        const int i0 = sh & (POLY_N - 1);
                        const uint64_t x0 = poly[i0], x1 = poly[i0 ^ POLY_M];
                        const uint64_t m0 = (uint64_t)0 - (uint64_t)((sh >> 9) & 1);
                        const uint64_t m1 = (uint64_t)0 - (uint64_t)(((sh >> 9) ^ (sh >> 8)) & 1);
                        const uint64_t y0 = ((x0 ^ m0) - m0) - poly[j] + DECOMP85_ADD;            // decomp85_first
                        const uint64_t y1 = ((x1 ^ m1) - m1) - poly[j + POLY_M] + DECOMP85_ADD;
                        lo_re[k] = (uint32_t)y0; st_re[n1] = (uint32_t)(y0 >> 32);
                        lo_im[k] = (uint32_t)y1; st_im[n1] = (uint32_t)(y1 >> 32);
                    }
                    p5[q4] = ws_gather_byte(lo_re[0], lo_re[1], lo_re[2], lo_re[3], 3);
                    p5[4 + q4] = ws_gather_byte(lo_im[0], lo_im[1], lo_im[2], lo_im[3], 3);
                }
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        pk[j] = ws_gather_byte(st_re[4 * j], st_re[4 * j + 1], st_re[4 * j + 2], st_re[4 * j + 3], b);
                        pk[4 + j] = ws_gather_byte(st_im[4 * j], st_im[4 * j + 1], st_im[4 * j + 2], st_im[4 * j + 3], b);
                    }
                    ws_tmem_st8(taddr + TM_DIGITS + 64 * s + b * 8, pk);
                }
                ws_tmem_st8(taddr + TM_DIGITS + 64 * s + 32, p5);
                ws_tmem_wait_st();
                __syncwarp();       // the copy has been gathered by every lane before the first spectrum overwrites it
                WT(0);
            };

            // Program of one half-period h (the MAC role consumes step i of set X = h & 1 meanwhile; Y is the other set): the forward
            // levels 4..1 of X, each due when the MAC role has consumed the level before it, with the end of Y's previous step, the
            // start of its next one and its first level slotted in between.  h = -1 and h = 2n are the ramps (only Y's part runs).
#pragma unroll 1
            for (int h = -1; h <= 2 * n; h++) {
                const int X = h & 1, Y = X ^ 1, i = h >> 1;
                const bool xlive = h >= 0 && h < 2 * n;
                const int y_fin = X == 0 ? i - 1 : i, y_new = y_fin + 1;
                // mask element of Y's next step: fetched a whole half-period's worth of work before it is used
                const uint64_t raw = a.lwe_in[(size_t)min(ct0 + Y * G + ct, a.count - 1) * (n + 1) + min(y_new, n - 1)];
#pragma unroll 1
                for (int op = 0; op < 7; op++) {
                    const int code = (WS2_ORDER >> (4 * op)) & 15;     // 1-4: forward level of X, 5: forward level 5 of Y, 6: finish Y, 7: start Y
                    if (code <= 5) {
                        const bool isx = code <= 4;
                        if (isx ? xlive : y_new < n) forward_level(isx ? X : Y, isx ? code : LEVELS, isx ? i : y_new);
                    } else if (code == 6) { if (y_fin >= 0) finish_step(Y, y_fin); }
                    else { if (y_new < n) start_step(Y, raw); }
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
#if WS2_NAMED_BAR
                        if ((r & 1) == 0) named_sync(1 + 3 * s + (r >> 1), WS_THREADS - WS_FFT_THREADS + 32 * ((r >> 1) == 2 ? 2 : 3));
                        if (r == 0) ws_mbar_wait(&sm.bfull[0], parity);
                        else if (r == BSPLIT) ws_mbar_wait(&sm.bfull[1], parity);
#elif WS2_RFULL_PER_ROW
                        if (r == 0) ws_mbar_wait2(&sm.bfull[0], parity, &sm.rfull[s][0], pr);
                        else if (r == BSPLIT) ws_mbar_wait2(&sm.bfull[1], parity, &sm.rfull[s][r], pr);
                        else ws_mbar_wait(&sm.rfull[s][r], pr);
#else
                        if (r == 0) ws_mbar_wait2(&sm.bfull[0], parity, &sm.rfull[s][0], pr);
                        else if (r == BSPLIT && (r & 1) == 0) ws_mbar_wait2(&sm.bfull[1], parity, &sm.rfull[s][r], pr);
                        else if (r == BSPLIT) ws_mbar_wait(&sm.bfull[1], parity);
                        else if ((r & 1) == 0) ws_mbar_wait(&sm.rfull[s][r], pr);
#endif
                        WT(0);
                        auto release_row = [&]() {
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
                        };
                        cd x[G];
#pragma unroll
                        for (int g = 0; g < G; g++) x[g] = sm.hs[s][r * G + g][p];
#if WS2_MAC_WIDE
                        {   // all K+1 key values first, then the four FMA passes over all (g, c): dependent FMAs are 15 apart
                            cd w[K + 1];
#pragma unroll
                            for (int c = 0; c <= K; c++) w[c] = sm.ring[r][c][p];
#pragma unroll
                            for (int c = 0; c <= K; c++)
#pragma unroll
                                for (int g = 0; g < G; g++) facc[g][c].x = fma(x[g].x, w[c].x, facc[g][c].x);
#if WS2_EARLY_RELEASE
                            // every 16-byte load of the row has fed an FMA above (in-order issue: they have all returned): the spectra
                            // slot and the key slot can go back to their producers three FMA passes before this warp is done with the row
                            release_row();
#endif
#pragma unroll
                            for (int c = 0; c <= K; c++)
#pragma unroll
                                for (int g = 0; g < G; g++) facc[g][c].y = fma(x[g].x, w[c].y, facc[g][c].y);
#pragma unroll
                            for (int c = 0; c <= K; c++)
#pragma unroll
                                for (int g = 0; g < G; g++) facc[g][c].x = fma(-x[g].y, w[c].y, facc[g][c].x);
#pragma unroll
                            for (int c = 0; c <= K; c++)
#pragma unroll
                                for (int g = 0; g < G; g++) facc[g][c].y = fma(x[g].y, w[c].x, facc[g][c].y);
                        }
#else
#pragma unroll
                        for (int c = 0; c <= K; c++) {
                            const cd w = sm.ring[r][c][p];
                            cmac_cols2<G, K + 1>(facc, c, x, w);
                        }
#endif
#if !(WS2_MAC_WIDE && WS2_EARLY_RELEASE)
                        release_row();
#endif
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
