// pbs_ws1_kernel.cu — warp-specialised PBS for ONE ciphertext per CTA with the decomposition levels transformed
// in parallel, sm_100a.  The latency kernel: waves of at most one ciphertext per SM (Server::aes_key_expansion,
// server.rs:107-167: 50 dependent stages of 32 bootstraps; the carry chain of the counter add; single-block calls).
//
// Same arithmetic as pbs_ws_kernel (cmux_core.cuh).  With one ciphertext only K+1 of the 16 FFT groups of
// pbs_ws_kernel have work, and the CMux step is a chain of LEVELS forward transforms per polynomial.  The byte-recoded
// decomposition yields the digits of all levels at once, so here LP = 3 groups share a polynomial:
//   group (r, ls), ls < LP: polynomial r, levels LEVELS - ls, LEVELS - ls - LP, ...     (5,2 / 4,1 / 3 at LEVELS = 5)
// Each of them rotates, subtracts and recodes polynomial r itself (cheap, integer) and transforms its own levels: the
// forward chain shrinks from LEVELS to ceil(LEVELS / LP) transforms.  Group (r, 0) also owns the inverse transform
// and the accumulator update of polynomial r and tells the other two (ACC[r]) when the accumulator is final.
//   MAC role: thread p owns Fourier point p; it consumes the rows in level order, waiting once per ROUND of
//   productions (ROUND[j]: every group that has a j-th level arrives once) instead of once per row, and hands each
//   slot back (REMPTY[slot]) as soon as its row is consumed.  Key ring and its barriers as in pbs_ws_kernel.
#include "../tfhe-aes_b200/csrc/ws_common.cuh"

#define WS1_LP 3

template <int K>
struct Ws1Smem {
    uint64_t acc[K + 1][POLY_N];                 // the accumulator (GLWE, standard domain)
    cd hs[(K + 1) * WS1_LP][XB_ELEMS];           // hand-over slot of group (r, ls) at [r * LP + ls]
    cd tw[256];
    cd ring[K + 1][K + 1][POLY_M];               // one level of the Fourier bootstrap key: [row][col][p]
    uint64_t round_full[2];                      // j-th productions of a step are all posted (j = 0, 1)
    uint64_t rempty[(K + 1) * WS1_LP];           // slot consumed (one arrival per MAC warp)
    uint64_t bfull[2];
    uint64_t bempty[K + 1];
    uint64_t inv;                                // Fourier accumulators handed over (one arrival per MAC warp)
    uint64_t acc_ready[K + 1];                   // accumulator polynomial r updated (one arrival)
};

template <int K, int BASE_LOG, int LEVELS>
__global__ void __launch_bounds__(WS_THREADS, 1) pbs_ws1_kernel(PbsArgs a) {
    static_assert(LEVELS <= 2 * WS1_LP, "two production rounds");
    static_assert(BASE_LOG == 8 && LEVELS == 5, "needs the byte-recoded decomposition: digits of any level on demand");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Ws1Smem<K> &sm = *reinterpret_cast<Ws1Smem<K> *>(smem_raw);
    constexpr int LP = WS1_LP;
    constexpr int NGROUPS = (K + 1) * LP;        // active 16-lane groups
    static_assert(NGROUPS <= 16, "FFT role has 16 groups");
    constexpr int RING = K + 1;
    constexpr int ROWS = LEVELS * (K + 1);
    constexpr int ROW_ELEMS = POLY_M * (K + 1);
    constexpr unsigned ROW_BYTES = (unsigned)(ROW_ELEMS * sizeof(cd));
    constexpr int BSPLIT = (K + 2) / 2;
    constexpr int MAC_REGS = 72;
    // productions per round: every group has a first level; groups with ls < LEVELS - LP have a second one
    constexpr int ROUND0 = NGROUPS;
    constexpr int ROUND1 = (K + 1) * (LEVELS - LP > 0 ? LEVELS - LP : 0);
    const int tid = threadIdx.x;
    const int n = a.lwe_dim;
    const int ct = min((int)blockIdx.x, a.count - 1);
    const int nrows = n * ROWS;

    for (int i = tid; i < 256; i += WS_THREADS) sm.tw[i] = a.tw[i];
    if (tid == 0) {
        ws_mbar_init(&sm.round_full[0], ROUND0);
        ws_mbar_init(&sm.round_full[1], ROUND1 > 0 ? ROUND1 : 1);
        for (int s = 0; s < NGROUPS; s++) ws_mbar_init(&sm.rempty[s], WS_MAC_WARPS);
        for (int r = 0; r <= K; r++) { ws_mbar_init(&sm.bempty[r], WS_MAC_WARPS); ws_mbar_init(&sm.acc_ready[r], 1); }
        ws_mbar_init(&sm.bfull[0], BSPLIT);
        ws_mbar_init(&sm.bfull[1], K + 1 - BSPLIT);
        ws_mbar_init(&sm.inv, WS_MAC_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const int rot = (2 * POLY_N - ws_mod_switch_2n(a, ct, n)) & (2 * POLY_N - 1);
        for (int idx = tid; idx < (K + 1) * POLY_N; idx += WS_THREADS) {
            const int r = idx / POLY_N, j = idx % POLY_N;
            sm.acc[r][j] = (r == K) ? rotated_coef(a.lut, j, rot) : 0;
        }
    }
    __syncthreads();

    if (tid >= WS_THREADS - WS_FFT_THREADS) {
        // ================================ FFT warps ================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(256 - MAC_REGS));
        const int ftid = tid - (WS_THREADS - WS_FFT_THREADS);
        const int gid = ftid >> 4, lane = ftid & 15;
        const bool active = gid < NGROUPS;
        const int r = active ? gid / LP : 0, ls = active ? gid % LP : 0;
        cd *slot = sm.hs[active ? gid : 0];
        cd v[16];
        uint32_t st_re[16], st_im[16];
        // warp-uniform view of the two groups (a, b) of this warp: all waits are issued warp-uniformly
        const int gid_a = (ftid >> 5) * 2, gid_b = gid_a + 1;
        const bool act_a = gid_a < NGROUPS, act_b = gid_b < NGROUPS;
        const int r_a = act_a ? gid_a / LP : 0, ls_a = act_a ? gid_a % LP : 0;
        const int r_b = act_b ? gid_b / LP : 0, ls_b = act_b ? gid_b % LP : 0;
        unsigned prod_a = 0, prod_b = 0;          // productions of the two groups so far
        const bool warp_has_owner = (act_a && ls_a == 0) || (act_b && ls_b == 0);
        auto wait_pair = [&](bool need_a, uint64_t *bar_a, unsigned par_a, bool need_b, uint64_t *bar_b, unsigned par_b) {
            if (need_a && need_b) ws_mbar_wait2(bar_a, par_a, bar_b, par_b);
            else if (need_a) ws_mbar_wait(bar_a, par_a);
            else if (need_b) ws_mbar_wait(bar_b, par_b);
        };
        // both 16-lane groups of a warp run ONE instruction stream; a group without work in the second round (or the
        // idle 16th group) runs along with its stores, waits and arrivals predicated off
        if ((ftid >> 5) * 2 < NGROUPS) {
            int rot = ws_mod_switch_2n(a, ct, 0);
#pragma unroll 1
            for (int i = 0; i < n; i++) {
                const int rot_next = (i + 1 < n) ? ws_mod_switch_2n(a, ct, i + 1) : 0;
                // polynomial r is final once its owner group (r, 0) has added the previous step's product
                if (i > 0) wait_pair(act_a && ls_a > 0, &sm.acc_ready[r_a], (i - 1) & 1, act_b && ls_b > 0, &sm.acc_ready[r_b], (i - 1) & 1);
                __syncwarp();
                load_decompose_rot<BASE_LOG, LEVELS>(sm.acc[r], lane, rot, v, st_re, st_im);
                rot = rot_next;
#pragma unroll 1
                for (int j = 0; j < 2; j++) {
                    const int lev = LEVELS - ls - LP * j;
                    const bool on = active && lev >= 1;
                    if (j == 1 && ROUND1 == 0) break;
                    if (lev != LEVELS) next_digits<BASE_LOG, LEVELS>(v, st_re, st_im, lev >= 1 ? lev : 1);
                    fft256_fwd_pass1_compute(v, lane, sm.tw);
                    const bool on_a = act_a && LEVELS - ls_a - LP * j >= 1, on_b = act_b && LEVELS - ls_b - LP * j >= 1;
                    wait_pair(on_a && prod_a > 0, &sm.rempty[gid_a], (prod_a - 1) & 1, on_b && prod_b > 0, &sm.rempty[act_b ? gid_b : gid_a], (prod_b - 1) & 1);
                    __syncwarp();
                    if (on) fft256_fwd_pass1_store(v, lane, slot);
                    __syncwarp();
                    fft256_fwd_pass2(v, lane, slot);
                    __syncwarp();
                    if (on) {
#pragma unroll
                        for (int k2 = 0; k2 < 16; k2++) slot[lane + 16 * k2] = v[rev4(k2)];
                    }
                    __syncwarp();
                    if (on && lane == 0) ws_mbar_arrive(&sm.round_full[j]);
                    prod_a += on_a; prod_b += on_b;
                }
                // inverse transform of the Fourier accumulator of polynomial r (owner groups only)
                const bool owner = active && ls == 0;
                if (!warp_has_owner) continue;                    // (warp-uniform) nothing to invert in this warp
                ws_mbar_wait(&sm.inv, i & 1);
                __syncwarp();
#pragma unroll
                for (int k2 = 0; k2 < 16; k2++) v[k2] = slot[lane + 16 * k2];
                fft256_inv_pass1_compute(v);
                __syncwarp();
                if (owner) fft256_inv_pass1_store(v, lane, sm.tw, slot);
                __syncwarp();
                fft256_inv_pass2(v, lane, slot);
                if (owner) {
                    uint64_t *poly = sm.acc[r];
#pragma unroll
                    for (int n1 = 0; n1 < 16; n1++) {
                        const int jj = 16 * n1 + lane;
                        poly[jj] += f64_to_torus(v[n1].x);
                        poly[jj + POLY_M] += f64_to_torus(v[n1].y);
                    }
                }
                __syncwarp();
                if (owner && lane == 0) ws_mbar_arrive(&sm.acc_ready[r]);
            }
        }
    } else {
        // ================================ MAC warps ================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MAC_REGS));
        const int p = tid;
        const int mwarp = p >> 5, mlane = p & 31;
        auto produce = [&](int q) {
            const int s = q % RING;
            uint64_t *bar = &sm.bfull[s < BSPLIT ? 0 : 1];
            ws_mbar_arrive_expect_tx(bar, ROW_BYTES);
            ws_bulk_copy_g2s(&sm.ring[s][0][0], a.bsk + (size_t)q * ROW_ELEMS, ROW_BYTES, bar);
        };
        if (p == 0)
            for (int q = 0; q < RING; q++) produce(q);
        cd facc[K + 1];
        unsigned level_count = 0;
        int q = 0;
#pragma unroll 1
        for (int i = 0; i < n; i++) {
#pragma unroll
            for (int c = 0; c <= K; c++) facc[c] = cmk(0.0, 0.0);
#pragma unroll 1
            for (int lev = LEVELS; lev >= 1; lev--) {
                const unsigned parity = level_count & 1;
                const int ls = (LEVELS - lev) % LP, round = (LEVELS - lev) / LP;
                // the first level of a round waits for all productions of that round
                if (ls == 0) ws_mbar_wait(&sm.round_full[round], i & 1);
#pragma unroll
                for (int r = 0; r <= K; r++, q++) {
                    if (r == 0) ws_mbar_wait(&sm.bfull[0], parity);
                    else if (r == BSPLIT) ws_mbar_wait(&sm.bfull[1], parity);
                    const int sl = r * LP + ls;
                    const cd x = sm.hs[sl][p];
#pragma unroll
                    for (int c = 0; c <= K; c++) cmac(facc[c], x, sm.ring[r][c][p]);
                    __syncwarp();
                    if (mlane == 0) {
                        ws_mbar_arrive(&sm.bempty[r]);
                        ws_mbar_arrive(&sm.rempty[sl]);
                        if (mwarp == (q & (WS_MAC_WARPS - 1)) && q >= 1 && q - 1 + RING < nrows) {
                            const int ps = (r + K) % RING;
                            const unsigned pp = (r == 0) ? (parity ^ 1) : parity;
                            ws_mbar_wait(&sm.bempty[ps], pp);
                            produce(q - 1 + RING);
                        }
                    }
                }
                level_count++;
            }
#pragma unroll
            for (int c = 0; c <= K; c++) sm.hs[c * LP][p] = facc[c];     // slot of the owner group (c, 0)
            __syncwarp();
            if (mlane == 0) ws_mbar_arrive(&sm.inv);
        }
    }
    __syncthreads();
    if ((int)blockIdx.x < a.count) {
        uint64_t *out = a.out + (size_t)blockIdx.x * (K * POLY_N + 1);
        for (int idx = tid; idx < K * POLY_N; idx += WS_THREADS) {
            const int r = idx / POLY_N, j = idx % POLY_N;
            out[idx] = (j == 0) ? sm.acc[r][0] : (uint64_t)0 - sm.acc[r][POLY_N - j];
        }
        if (tid == 0) out[K * POLY_N] = sm.acc[K][0] + a.post_add;
    }
}

#define LAUNCH_WS1(k, bl, lv)                                                                           \
    if (K == k && base_log == bl && levels == lv) {                                                     \
        const size_t smem = sizeof(Ws1Smem<k>);                                                         \
        cudaError_t e = cudaFuncSetAttribute(pbs_ws1_kernel<k, bl, lv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                 \
        pbs_ws1_kernel<k, bl, lv><<<a.count, WS_THREADS, smem, s>>>(a);                                 \
        return cudaGetLastError();                                                                      \
    }
cudaError_t launch_pbs_ws1(int K, int base_log, int levels, const PbsArgs &a, cudaStream_t s) {
    LAUNCH_WS1(4, 8, 5) LAUNCH_WS1(1, 8, 5)
    return cudaErrorInvalidValue;
}
