// imma_kernels.cu — LWE keyswitch (K1) and private functional packing keyswitch (K4) as an EXACT integer
// GEMM on the int8 tensor path (mma.sync.m16n8k32 s8 x u8 -> s32), sm_100a.
//
//   out[bit][col] -= sum_rows d[bit][row] * key[row][col]        (mod 2^64; SURVEY §9.4(1),(5))
//
// Batched over the bits of all resident ciphertexts this is a dense contraction
// [bits x rows] . [rows x cols] (PFKS: rows = 2049*3 = 6147, cols = 5*2560; KS: rows = 2048*6, cols = 670).
// It is evaluated limb-wise, bit-exactly:
//   key  = sum_{b<8} 2^(8b) * kb,  kb in [0,255]            (u8 limbs, pre-arranged once at key load)
//   d    = dl + 128*dh,  dl in [-64,63], dh in [-16,16]     (s8 limbs; KS digits fit in dl alone)
//   sum_rows d*key = sum_b 2^(8b) * ( sum_rows dl*kb  +  128 * sum_rows dh*kb )
// Each inner sum is an s32 tensor-core accumulation (|.| <= 6147*64*255 < 2^27: no overflow); the 16 partial
// sums per output are recombined with shifts in u64 in the epilogue.  Measured on this B200 the scalar
// formulation (IMAD.WIDE.U32 + IMAD) is multiplier-pipe bound at ~14 u64-MAC/clk/SM; this path has a
// ceiling of 1966/16 = 123 u64-MAC/clk/SM (legacy mma.sync rate measured by scratch/mb_imma.cu).
//
// Key layout (built by imma_prepare_key_kernel): [key][ntile = col/8][kchunk = row/32][2048 B], each
// 2 KB chunk = the 16 B-fragments of one k32 x n8 tile for all 8 limbs:
//   [q<4][lane<32][ limb 2q: b0 b1 | limb 2q+1: b0 b1 ]   with the mma B-fragment convention
//   b0 = bytes (k = 4t..4t+3, n = g), b1 = bytes (k = 16+4t.., n = g), g = lane>>2, t = lane&3.
#include "kernels.h"

#define IG_THREADS 256
#define IG_MTILE 128           // bits per CTA (8 warps x 16)
#define IG_CHUNK_BYTES 2048

__global__ void imma_prepare_key_kernel(const uint64_t *__restrict__ key, size_t key_stride, int rows, int ncols, int row_stride,
                                        int ntiles, int kchunks, uint8_t *__restrict__ kp) {
    // one thread per (key, ntile, kchunk, q, lane): writes 16 bytes
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per_key = (long)ntiles * kchunks * 128;
    const int keyi = blockIdx.y;
    if (idx >= per_key) return;
    const int lane = idx & 31, q = (idx >> 5) & 3;
    const long chunk = idx >> 7;
    const int kc = (int)(chunk % kchunks), nt = (int)(chunk / kchunks);
    const int g = lane >> 2, t = lane & 3;
    const int col = nt * 8 + g;
    uint64_t w[8];  // rows 4t..4t+3 and 16+4t..16+4t+3 of this chunk, column col
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int row = kc * 32 + (i < 4 ? 4 * t + i : 16 + 4 * t + (i - 4));
        w[i] = (row < rows && col < ncols) ? key[(size_t)keyi * key_stride + (size_t)row * row_stride + col] : 0;
    }
    uint32_t o[4];
#pragma unroll
    for (int h = 0; h < 2; h++) {       // limb 2q+h
        const int sh = 8 * (2 * q + h);
        uint32_t b0 = 0, b1 = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            b0 |= (uint32_t)((w[i] >> sh) & 0xFF) << (8 * i);
            b1 |= (uint32_t)((w[4 + i] >> sh) & 0xFF) << (8 * i);
        }
        o[2 * h] = b0; o[2 * h + 1] = b1;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(kp + ((size_t)keyi * ntiles * kchunks + chunk) * IG_CHUNK_BYTES) + q * 32 + lane;
    *dst = make_uint4(o[0], o[1], o[2], o[3]);
}
cudaError_t launch_imma_prepare_key(const uint64_t *key, size_t key_stride, int nkeys, int rows, int ncols, int row_stride, uint8_t *kp,
                                    cudaStream_t s) {
    const int ntiles = (ncols + 7) / 8, kchunks = (rows + 31) / 32;
    const long per_key = (long)ntiles * kchunks * 128;
    dim3 grid((unsigned)((per_key + 255) / 256), nkeys);
    imma_prepare_key_kernel<<<grid, 256, 0, s>>>(key, key_stride, rows, ncols, row_stride, ntiles, kchunks, kp);
    return cudaGetLastError();
}

// signed decomposition -> s8 limb planes dl, dh: [bit][rows_pad], row = j*levels + (level-1)
__global__ void imma_decompose_kernel(const uint64_t *__restrict__ in, int in_stride, int nelem, long total, int base_log, int levels,
                                      int rows_pad, int8_t *__restrict__ dl, int8_t *__restrict__ dh) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long b = idx / nelem;
    const int j = (int)(idx % nelem);
    const uint64_t x = in[b * in_stride + j];
    const int r = 64 - base_log * levels;
    uint64_t state = ((x >> r) + ((x >> (r - 1)) & 1)) & (~0ull >> r);
    const uint64_t mask = (1ull << base_log) - 1;
    int8_t *pl = dl + b * rows_pad + (size_t)j * levels;
    int8_t *ph = dh ? dh + b * rows_pad + (size_t)j * levels : nullptr;
    for (int l = levels; l >= 1; l--) {
        uint64_t res = state & mask;
        state >>= base_log;
        uint64_t carry = ((res - 1) | state) & res;
        carry >>= (base_log - 1);
        state += carry;
        const int d = (int)res - (int)(carry << base_log);
        const int lo = ((d + 64) & 127) - 64;     // in [-64, 63]
        pl[l - 1] = (int8_t)lo;
        if (ph) ph[l - 1] = (int8_t)((d - lo) >> 7);  // exact: d - lo is a multiple of 128
    }
}
cudaError_t launch_imma_decompose(const uint64_t *in, int in_stride, int nelem, int count, int base_log, int levels, int rows_pad,
                                  int8_t *dl, int8_t *dh, cudaStream_t s) {
    const long total = (long)count * nelem;
    imma_decompose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, in_stride, nelem, total, base_log, levels, rows_pad, dl, dh);
    return cudaGetLastError();
}

__device__ __forceinline__ void mma_s8u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}

// CTA: 128 bits x (8*NT) columns of one key, all rows.  DL = digit limbs (1: KS, 2: PFKS).
template <int NT, int DL>
__global__ void __launch_bounds__(IG_THREADS, (NT * DL >= 4) ? 1 : 2) imma_gemv_kernel(ImmaGemvArgs a) {
    __shared__ __align__(16) uint8_t kbuf[3][NT][IG_CHUNK_BYTES];   // 3-stage ring: one barrier per 32-row chunk
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * IG_MTILE + warp * 16;
    const int nt0 = blockIdx.y * NT;
    const int keyi = blockIdx.z;
    const uint8_t *kp = a.kp + ((size_t)keyi * a.ntiles + nt0) * a.kchunks * IG_CHUNK_BYTES;
    int acc[NT][8][DL][4];
#pragma unroll
    for (int n = 0; n < NT; n++)
#pragma unroll
        for (int b = 0; b < 8; b++)
#pragma unroll
            for (int l = 0; l < DL; l++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[n][b][l][e] = 0;
    // digit rows of this lane: bits m0+g and m0+g+8 (zero beyond count)
    const bool ok0 = (m0 + g) < a.count, ok1 = (m0 + g + 8) < a.count;
    const int8_t *dl0 = a.dl + (size_t)(ok0 ? m0 + g : 0) * a.rows_pad + 4 * t;
    const int8_t *dl1 = a.dl + (size_t)(ok1 ? m0 + g + 8 : 0) * a.rows_pad + 4 * t;
    const int8_t *dh0 = DL == 2 ? a.dh + (size_t)(ok0 ? m0 + g : 0) * a.rows_pad + 4 * t : nullptr;
    const int8_t *dh1 = DL == 2 ? a.dh + (size_t)(ok1 ? m0 + g + 8 : 0) * a.rows_pad + 4 * t : nullptr;
    auto fill = [&](int buf, int kc) {
        if (kc < a.kchunks) {
            for (int i = tid; i < NT * 128; i += IG_THREADS) {  // NT chunks of 2 KB = NT*128 transfers of 16 B
                const int n = i >> 7, off = (i & 127) * 16;
                const bool valid = (nt0 + n) < a.ntiles;
                cp_async_16(&kbuf[buf][n][off], kp + ((size_t)(valid ? n : 0) * a.kchunks + kc) * IG_CHUNK_BYTES + off);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    auto load_a = [&](uint32_t (&af)[DL][4], int kc) {   // A fragments (digits) of chunk kc straight from global / L2
        const int o = kc * 32;
        af[0][0] = ok0 ? __ldg(reinterpret_cast<const uint32_t *>(dl0 + o)) : 0;
        af[0][1] = ok1 ? __ldg(reinterpret_cast<const uint32_t *>(dl1 + o)) : 0;
        af[0][2] = ok0 ? __ldg(reinterpret_cast<const uint32_t *>(dl0 + o + 16)) : 0;
        af[0][3] = ok1 ? __ldg(reinterpret_cast<const uint32_t *>(dl1 + o + 16)) : 0;
        if (DL == 2) {
            af[DL - 1][0] = ok0 ? __ldg(reinterpret_cast<const uint32_t *>(dh0 + o)) : 0;
            af[DL - 1][1] = ok1 ? __ldg(reinterpret_cast<const uint32_t *>(dh1 + o)) : 0;
            af[DL - 1][2] = ok0 ? __ldg(reinterpret_cast<const uint32_t *>(dh0 + o + 16)) : 0;
            af[DL - 1][3] = ok1 ? __ldg(reinterpret_cast<const uint32_t *>(dh1 + o + 16)) : 0;
        }
    };
    fill(0, 0);
    fill(1, 1);
    uint32_t af[DL][4], af_next[DL][4];
    load_a(af_next, 0);
    for (int kc = 0; kc < a.kchunks; kc++) {
#pragma unroll
        for (int l = 0; l < DL; l++)
#pragma unroll
            for (int e = 0; e < 4; e++) af[l][e] = af_next[l][e];
        if (kc + 1 < a.kchunks) load_a(af_next, kc + 1);       // in flight while this chunk's MMAs run
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");   // chunk kc has landed
        __syncthreads();                                         // ... for everyone, and chunk kc-1 is consumed
        fill((kc + 2) % 3, kc + 2);
        const int buf = kc % 3;
#pragma unroll
        for (int n = 0; n < NT; n++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 bw = *reinterpret_cast<const uint4 *>(&kbuf[buf][n][(q * 32 + lane) * 16]);
#pragma unroll
                for (int l = 0; l < DL; l++) {
                    mma_s8u8(acc[n][2 * q][l], af[l], bw.x, bw.y);
                    mma_s8u8(acc[n][2 * q + 1][l], af[l], bw.z, bw.w);
                }
            }
        }
    }
    // epilogue: recombine the limb partial sums (u64 wrapping) and subtract from the output
#pragma unroll
    for (int n = 0; n < NT; n++) {
        if (nt0 + n >= a.ntiles) continue;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int bit = m0 + g + 8 * (e >> 1);
            const int col = (nt0 + n) * 8 + 2 * t + (e & 1);
            if (bit >= a.count || col >= a.ncols) continue;
            uint64_t v = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                int64_t p = (int64_t)acc[n][b][0][e];
                if (DL == 2) p += (int64_t)acc[n][b][DL - 1][e] * 128;
                v += (uint64_t)p << (8 * b);
            }
            a.out[(size_t)bit * a.out_stride + (size_t)keyi * a.ncols + col] -= v;
        }
    }
}
cudaError_t launch_imma_gemv(const ImmaGemvArgs &a, int digit_limbs, cudaStream_t s) {
    if (digit_limbs == 2) {
        // 128 bits x 8 columns per CTA, 2 CTAs per SM (a 16-column tile with one CTA per SM measured slower:
        // 18.5 vs 15.4 ms per 3072 bits)
        if (!a.dh) return cudaErrorInvalidValue;
        dim3 grid((a.count + IG_MTILE - 1) / IG_MTILE, a.ntiles, a.nkeys);
        imma_gemv_kernel<1, 2><<<grid, IG_THREADS, 0, s>>>(a);
    } else {
        dim3 grid((a.count + IG_MTILE - 1) / IG_MTILE, (a.ntiles + 1) / 2, a.nkeys);
        imma_gemv_kernel<2, 1><<<grid, IG_THREADS, 0, s>>>(a);
    }
    return cudaGetLastError();
}
