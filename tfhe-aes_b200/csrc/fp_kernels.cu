// fp_kernels.cu — the FP64 kernels of the WoPBS chain (sm_100a):
//   pbs_kernel            blind rotation + sample extract (PBS) — K3 of SURVEY.md §2.2
//                         (circuit_bootstrap_boolean -> bootstrap, many_wopbs.rs:253; extract_bits loop :194)
//   vp_kernel             blind-rotation half of vertical_packing + sample extract — K6 (many_wopbs.rs:277)
//   cmux_tree_kernel      one layer of the vertical-packing CMux tree (LUT larger than N) — K6 general form
//   fourier_convert_kernel standard GGSW -> Fourier layout — K5 (fill_with_forward_fourier, many_wopbs.rs:263)
//                         and the one-time bootstrap-key conversion at key load.
// All of them run the CTA-level phases of cmux_core.cuh.
#include "cmux_core.cuh"
#include "kernels.h"

template <int K, int G>
__device__ __forceinline__ CmuxSmem<K, G> &smem_view(unsigned char *raw) {
    return *reinterpret_cast<CmuxSmem<K, G> *>(raw);
}

template <int K, int G>
__device__ __forceinline__ void load_twiddles(CmuxSmem<K, G> &sm, const cd *tw, int tid) {
    cd *dst = sm.twf;  // twf and twi are contiguous
    for (int i = tid; i < 512; i += CMUX_THREADS) dst[i] = tw[i];
}

// One full CMux step on the resident accumulators (all barriers included).
template <int K, int G, int BASE_LOG, int LEVELS, int MODE>
__device__ __forceinline__ void cmux_step(int tid, CmuxSmem<K, G> &sm, CmuxRegs<K, G> &rg, const cd *__restrict__ ggsw,
                                          const uint64_t *const *ext) {
    phase_load_decompose<K, G, BASE_LOG, LEVELS, MODE>(tid, sm, rg, ext);
#pragma unroll 1
    for (int lev = LEVELS; lev >= 1; lev--) {
        if (lev != LEVELS) phase_next_digits<K, G, BASE_LOG>(tid, rg);
        phase_fwd1<K, G>(tid, sm, rg);
        __syncwarp();
        phase_fwd2<K, G>(tid, sm, rg);
        __syncwarp();
        phase_fwd3<K, G>(tid, sm, rg);
        __syncthreads();
        phase_mac<K, G>(tid, sm, rg, ggsw + (size_t)(LEVELS - lev) * (K + 1) * POLY_M * (K + 1));
        __syncthreads();
    }
    phase_inv0<K, G>(tid, sm, rg);
    __syncthreads();
    phase_inv1<K, G>(tid, sm, rg);
    __syncwarp();
    phase_inv2<K, G>(tid, sm, rg);
    __syncwarp();
    phase_inv3<K, G>(tid, sm, rg);
    __syncthreads();
}

// sample extract of coefficient 0 from the resident accumulator g
template <int K, int G>
__device__ __forceinline__ void sample_extract(int tid, const CmuxSmem<K, G> &sm, int g, uint64_t *out, uint64_t post_add) {
    for (int idx = tid; idx < K * POLY_N; idx += CMUX_THREADS) {
        const int r = idx / POLY_N, j = idx % POLY_N;
        out[idx] = (j == 0) ? sm.acc[g][r][0] : (uint64_t)0 - sm.acc[g][r][POLY_N - j];
    }
    if (tid == 0) out[K * POLY_N] = sm.acc[g][K][0] + post_add;
}

// ------------------------------------------------------------------------------------------------
// PBS: out[ct] = SampleExtract( BlindRotate(lut * X^-b~, a~, BSK) ) + post_add on the body
//
// Bootstrap-key streaming: the Fourier BSK slice of one CMux step is 25 rows x 20 KB (K=4).  Every
// thread owns Fourier point p = tid and needs, per row, the K+1 complex values [row][0..K][p]
// (each cp.async instruction of a warp moves 512 contiguous bytes).
// Those are prefetched with cp.async.cg (L2 -> shared memory, no register staging) into a ring of
// BSK_RING row-slots, thread-private (each thread only reads what it copied itself, so
// cp.async.wait_group is the only synchronisation), BSK_RING rows ahead of the multiply-accumulate
// that consumes them — across level and iteration boundaries, so the L2 latency is hidden behind the
// FFT phases.  All CTAs walk the key in the same order, so the slice is an L2 hit for all but the first.
// ------------------------------------------------------------------------------------------------
#define BSK_RING 4
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int K, int G, int BASE_LOG, int LEVELS>
__global__ void __launch_bounds__(CMUX_THREADS, 1) pbs_kernel(PbsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CmuxSmem<K, G> &sm = smem_view<K, G>(smem_raw);
    cd *ring = reinterpret_cast<cd *>(smem_raw + sizeof(CmuxSmem<K, G>));            // [BSK_RING][256][K+1]
    uint16_t *ahat = reinterpret_cast<uint16_t *>(ring + BSK_RING * POLY_M * (K + 1));
    CmuxRegs<K, G> rg;
    const int tid = threadIdx.x;
    const int n = a.lwe_dim, np = a.lwe_dim + 1;
    const int ct0 = blockIdx.x * G;
    constexpr int ROWS = LEVELS * (K + 1);               // GGSW rows per CMux step, stored in consumption order
    constexpr size_t ROW_ELEMS = (size_t)POLY_M * (K + 1);
    constexpr unsigned ROW_BYTES = (unsigned)(ROW_ELEMS * sizeof(cd));
    constexpr unsigned RING_BYTES = BSK_RING * ROW_BYTES;
    // prefetch head: the key is walked linearly, one row (K+1 polynomials of 256 points) at a time
    const cd *pf_src = a.bsk + tid;
    long pf_left = (long)n * ROWS;
    const unsigned ring_u32 = (unsigned)__cvta_generic_to_shared(ring) + tid * (unsigned)sizeof(cd);
    unsigned pf_off = 0;
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
        cp_async_commit();                               // always commit: keeps the group count uniform
    };
#pragma unroll
    for (int s = 0; s < BSK_RING; s++) issue();

    load_twiddles<K, G>(sm, a.tw, tid);
    // modulus switch to 2N: a~ = (a + 2^53) >> 54
    for (int idx = tid; idx < G * np; idx += CMUX_THREADS) {
        const int g = idx / np, i = idx % np;
        const int ct = min(ct0 + g, a.count - 1);
        uint64_t x = a.lwe_in[(size_t)ct * np + i] * a.in_scale;
        if (i == n) x += a.pre_add_body;
        ahat[g * np + i] = (uint16_t)((x + (1ull << 53)) >> 54);
    }
    __syncthreads();
    for (int g = 0; g < G; g++) {
        const int bhat = ahat[g * np + n];
        const int rot = (2 * POLY_N - bhat) & (2 * POLY_N - 1);
        for (int idx = tid; idx < (K + 1) * POLY_N; idx += CMUX_THREADS) {
            const int r = idx / POLY_N, j = idx % POLY_N;
            sm.acc[g][r][j] = (r == K) ? rotated_coef(a.lut, j, rot) : 0;
        }
    }
    if (tid < G) sm.rot[tid] = ahat[tid * np];
    __syncthreads();
    unsigned rd_off = 0;                                 // ring offset of the next row to consume
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        // a step whose rotations are all zero adds exactly zero (ct1 == 0): it is executed like any
        // other so that the key walk stays in lock-step with the prefetch ring
        phase_load_decompose<K, G, BASE_LOG, LEVELS, DIFF_ROTATE>(tid, sm, rg, nullptr);
#pragma unroll 1
        for (int lev = LEVELS; lev >= 1; lev--) {
            phase_fwd1<K, G>(tid, sm, rg);
            __syncwarp();
            phase_fwd2<K, G>(tid, sm, rg);
            __syncwarp();
            phase_fwd3<K, G>(tid, sm, rg);
            __syncthreads();
#pragma unroll
            for (int r = 0; r <= K; r++) {
                cp_async_wait<BSK_RING - 1>();
                phase_mac_row<K, G>(tid, sm, rg, r, reinterpret_cast<const cd *>(reinterpret_cast<const unsigned char *>(ring) + rd_off) + tid);
                rd_off += ROW_BYTES;
                if (rd_off == RING_BYTES) rd_off = 0;
                issue();
                // the digits of the next level (integer + conversion pipes) are extracted in the shadow of the
                // multiply-accumulate (FP64 pipe): v is free during this phase
                if (r == 0 && lev > 1) phase_next_digits<K, G, BASE_LOG>(tid, rg);
            }
            __syncthreads();
        }
        phase_inv0<K, G>(tid, sm, rg);
        __syncthreads();
        phase_inv1<K, G>(tid, sm, rg);
        __syncwarp();
        phase_inv2<K, G>(tid, sm, rg);
        __syncwarp();
        phase_inv3<K, G>(tid, sm, rg);
        if (tid < G && i + 1 < n) sm.rot[tid] = ahat[tid * np + i + 1];
        __syncthreads();
    }
    cp_async_wait<0>();
    for (int g = 0; g < G; g++)
        if (ct0 + g < a.count) sample_extract<K, G>(tid, sm, g, a.out + (size_t)(ct0 + g) * (K * POLY_N + 1), a.post_add);
}

// ------------------------------------------------------------------------------------------------
// Vertical packing, blind-rotation half: for GGSW bit t = 0 (LSB) .. nrot-1:
//   acc += GGSW_t (x) (acc * X^-(2^t) - acc);  then sample extract.
// blockIdx.y = job (one encrypted byte = one GGSW list), blockIdx.x = chunk of G outputs.
// ------------------------------------------------------------------------------------------------
template <int K, int G, int BASE_LOG, int LEVELS>
__global__ void __launch_bounds__(CMUX_THREADS, 1) vp_kernel(VpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CmuxSmem<K, G> &sm = smem_view<K, G>(smem_raw);
    CmuxRegs<K, G> rg;
    const int tid = threadIdx.x;
    const int job = blockIdx.y, o0 = blockIdx.x * G;
    load_twiddles<K, G>(sm, a.tw, tid);
    for (int g = 0; g < G; g++) {
        const int o = min(o0 + g, a.nouts - 1);
        for (int idx = tid; idx < (K + 1) * POLY_N; idx += CMUX_THREADS) {
            const int r = idx / POLY_N, j = idx % POLY_N;
            uint64_t v;
            if (a.glwe_init) v = a.glwe_init[((size_t)job * a.nouts + o) * (K + 1) * POLY_N + idx];
            else v = (r == K) ? a.lut[(size_t)job * a.lut_job_stride + (size_t)o * a.lut_out_stride + j] : 0;
            sm.acc[g][r][j] = v;
        }
    }
    __syncthreads();
    const size_t ggsw_size = (size_t)LEVELS * (K + 1) * POLY_M * (K + 1);
    const cd *ggsw_job = a.ggsw_f + (size_t)job * a.nbits * ggsw_size;
#pragma unroll 1
    for (int t = 0; t < a.nrot; t++) {
        const int deg = (t < 31 ? (1 << t) : 0) & (2 * POLY_N - 1);
        if (tid < G) sm.rot[tid] = (2 * POLY_N - deg) & (2 * POLY_N - 1);
        __syncthreads();
        cmux_step<K, G, BASE_LOG, LEVELS, DIFF_ROTATE>(tid, sm, rg, ggsw_job + (size_t)t * ggsw_size, nullptr);
    }
    for (int g = 0; g < G; g++)
        if (o0 + g < a.nouts)
            sample_extract<K, G>(tid, sm, g, a.out + ((size_t)job * a.nouts + o0 + g) * (K * POLY_N + 1), 0);
}

// ------------------------------------------------------------------------------------------------
// One layer of the CMux tree: out[q] = c0[q] + GGSW (x) (c1[q] - c0[q]),  c0 = in[2q], c1 = in[2q+1].
// blockIdx.y = job, blockIdx.x = chunk of G pairs (pairs enumerate (output, pair-in-output)).
// ------------------------------------------------------------------------------------------------
template <int K, int G, int BASE_LOG, int LEVELS>
__global__ void __launch_bounds__(CMUX_THREADS, 1) cmux_tree_kernel(TreeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CmuxSmem<K, G> &sm = smem_view<K, G>(smem_raw);
    __shared__ const uint64_t *ext[G];
    CmuxRegs<K, G> rg;
    const int tid = threadIdx.x;
    const int job = blockIdx.y, q0 = blockIdx.x * G;
    const size_t gsz = (size_t)(K + 1) * POLY_N;
    load_twiddles<K, G>(sm, a.tw, tid);
    for (int g = 0; g < G; g++) {
        const int q = min(q0 + g, a.npairs - 1);
        const uint64_t *c0 = a.in + ((size_t)job * a.npairs * 2 + 2 * q) * gsz;
        if (tid == 0) ext[g] = c0 + gsz;
        for (int idx = tid; idx < (int)gsz; idx += CMUX_THREADS) sm.acc[g][idx / POLY_N][idx % POLY_N] = c0[idx];
    }
    __syncthreads();
    const size_t ggsw_size = (size_t)LEVELS * (K + 1) * POLY_M * (K + 1);
    const cd *ggsw = a.ggsw_f + ((size_t)job * a.nbits + a.bit_index) * ggsw_size;
    cmux_step<K, G, BASE_LOG, LEVELS, DIFF_EXTERNAL>(tid, sm, rg, ggsw, ext);
    for (int g = 0; g < G; g++)
        if (q0 + g < a.npairs) {
            uint64_t *o = a.out + ((size_t)job * a.npairs + q0 + g) * gsz;
            for (int idx = tid; idx < (int)gsz; idx += CMUX_THREADS) o[idx] = sm.acc[g][idx / POLY_N][idx % POLY_N];
        }
}

// ------------------------------------------------------------------------------------------------
// standard-domain GGSW list [g][level][row][col][N] (u64) -> Fourier layout [g][level][row][col][p]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CMUX_THREADS, 2) fourier_convert_kernel(ConvertArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *xb_all = reinterpret_cast<cd *>(smem_raw);
    cd *tw = xb_all + CMUX_GROUPS * XB_ELEMS;
    const int tid = threadIdx.x, gid = tid >> 4, lane = tid & 15;
    for (int i = tid; i < 512; i += CMUX_THREADS) tw[i] = a.tw[i];
    __syncthreads();
    const long q = (long)blockIdx.x * CMUX_GROUPS + gid;
    const bool active = q < a.npoly;
    cd *xb = xb_all + gid * XB_ELEMS;
    cd v[16];
    if (active) {
        load_torus_poly(v, lane, a.in + (size_t)q * POLY_N);
        fft256_fwd_pass1(v, lane, tw, xb);
    }
    __syncwarp();
    if (active) {
        fft256_fwd_pass2(v, lane, xb);
        // [g][level][row][col] -> [g][level slot = levels-1-level][row][col][p]
        const long per_level = (long)(a.glwe_dim + 1) * (a.glwe_dim + 1);
        const long rc = q % per_level, gl = q / per_level;
        const long lev = gl % a.levels, g = gl / a.levels;
        cd *dst = a.out + (size_t)((g * a.levels + (a.levels - 1 - lev)) * per_level + rc) * POLY_M;
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) dst[lane + 16 * k2] = v[rev4(k2)];
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <typename F>
static cudaError_t set_smem(F f, size_t bytes) {
    return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#define LAUNCH_PBS(k, g, bl, lv)                                                               \
    if (K == k && G == g && base_log == bl && levels == lv) {                                  \
        size_t smem = sizeof(CmuxSmem<k, g>) + (size_t)BSK_RING * POLY_M * (k + 1) * sizeof(cd) + \
                      (size_t)g * (a.lwe_dim + 1) * sizeof(uint16_t);                            \
        cudaError_t e = set_smem(pbs_kernel<k, g, bl, lv>, smem);                              \
        if (e != cudaSuccess) return e;                                                        \
        pbs_kernel<k, g, bl, lv><<<(a.count + g - 1) / g, CMUX_THREADS, smem, s>>>(a);         \
        return cudaGetLastError();                                                             \
    }
cudaError_t launch_pbs(int K, int G, int base_log, int levels, const PbsArgs &a, cudaStream_t s) {
    LAUNCH_PBS(4, 1, 8, 5) LAUNCH_PBS(4, 2, 8, 5) LAUNCH_PBS(4, 3, 8, 5)
    LAUNCH_PBS(1, 1, 8, 5) LAUNCH_PBS(1, 4, 8, 5) LAUNCH_PBS(1, 8, 8, 5)
    return cudaErrorInvalidValue;
}
#define LAUNCH_VP(k, g, bl, lv)                                                      \
    if (K == k && G == g && base_log == bl && levels == lv) {                        \
        size_t smem = sizeof(CmuxSmem<k, g>);                                        \
        cudaError_t e = set_smem(vp_kernel<k, g, bl, lv>, smem);                     \
        if (e != cudaSuccess) return e;                                              \
        dim3 grid((a.nouts + g - 1) / g, a.njobs);                                   \
        vp_kernel<k, g, bl, lv><<<grid, CMUX_THREADS, smem, s>>>(a);                 \
        return cudaGetLastError();                                                   \
    }
cudaError_t launch_vp(int K, int G, int base_log, int levels, const VpArgs &a, cudaStream_t s) {
    LAUNCH_VP(4, 1, 15, 1) LAUNCH_VP(4, 2, 15, 1) LAUNCH_VP(4, 3, 15, 1)
    LAUNCH_VP(1, 1, 15, 1) LAUNCH_VP(1, 4, 15, 1) LAUNCH_VP(1, 8, 15, 1)
    return cudaErrorInvalidValue;
}
#define LAUNCH_TREE(k, g, bl, lv)                                                    \
    if (K == k && G == g && base_log == bl && levels == lv) {                        \
        size_t smem = sizeof(CmuxSmem<k, g>);                                        \
        cudaError_t e = set_smem(cmux_tree_kernel<k, g, bl, lv>, smem);              \
        if (e != cudaSuccess) return e;                                              \
        dim3 grid((a.npairs + g - 1) / g, a.njobs);                                  \
        cmux_tree_kernel<k, g, bl, lv><<<grid, CMUX_THREADS, smem, s>>>(a);          \
        return cudaGetLastError();                                                   \
    }
cudaError_t launch_cmux_tree(int K, int G, int base_log, int levels, const TreeArgs &a, cudaStream_t s) {
    LAUNCH_TREE(4, 3, 15, 1) LAUNCH_TREE(4, 1, 15, 1) LAUNCH_TREE(1, 1, 15, 1) LAUNCH_TREE(1, 8, 15, 1)
    return cudaErrorInvalidValue;
}
cudaError_t launch_fourier_convert(const ConvertArgs &a, cudaStream_t s) {
    size_t smem = (size_t)(CMUX_GROUPS * XB_ELEMS + 512) * sizeof(cd);
    cudaError_t e = set_smem(fourier_convert_kernel, smem);
    if (e != cudaSuccess) return e;
    long blocks = (a.npoly + CMUX_GROUPS - 1) / CMUX_GROUPS;
    fourier_convert_kernel<<<(unsigned)blocks, CMUX_THREADS, smem, s>>>(a);
    return cudaGetLastError();
}
