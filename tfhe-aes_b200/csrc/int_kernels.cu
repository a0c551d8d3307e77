// int_kernels.cu — integer (u64 wrapping) kernels of the WoPBS chain (sm_100a):
//   decompose + gemv   LWE keyswitch (K1, many_wopbs.rs:194-199 -> tfhe-rs keyswitch_lwe_ciphertext) and
//                      private functional packing keyswitch (K4, inside circuit_bootstrap_boolean, :253):
//                      out = init - sum_rows d[row] * key[row][:]   (SURVEY §9.4(1),(5))
//   lwe_sum            fused ShiftRows / MixColumns / AddRoundKey additions (K7; server.rs:278-282,
//                      mix_columns.rs:4-78, inv_mix_columns.rs:4-58)
// Digits are stored offset (u = d + beta/2 >= 0) so the inner product is an unsigned 32x64 multiply
// (IMAD.WIDE.U32 + IMAD); the offset is undone exactly with a pre-computed column sum of the key:
//   sum d*key = sum u*key - beta/2 * colsum(key)     (mod 2^64, bit-exact).
#include "kernels.h"

// ---- signed decomposition, digits stored as u = d + beta/2 at slot level-1 ----------
__global__ void decompose_kernel(const uint64_t *__restrict__ in, int in_stride, int nelem, long total, int base_log,
                                 int levels, uint16_t *__restrict__ digits) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long b = idx / nelem;
    const int j = (int)(idx % nelem);
    const uint64_t x = in[b * in_stride + j];
    const int r = 64 - base_log * levels;
    uint64_t state = ((x >> r) + ((x >> (r - 1)) & 1)) & (~0ull >> r);
    const uint64_t mask = (1ull << base_log) - 1;
    const uint32_t off = 1u << (base_log - 1);
    uint16_t *dst = digits + idx * levels;
    for (int l = levels; l >= 1; l--) {
        uint64_t res = state & mask;
        state >>= base_log;
        uint64_t carry = ((res - 1) | state) & res;
        carry >>= (base_log - 1);
        state += carry;
        int d = (int)res - (int)(carry << base_log);
        dst[l - 1] = (uint16_t)(d + (int)off);
    }
}
cudaError_t launch_decompose(const uint64_t *in, int in_stride, int nelem, int count, int base_log, int levels,
                             uint16_t *digits, cudaStream_t s) {
    long total = (long)count * nelem;
    decompose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, in_stride, nelem, total, base_log, levels, digits);
    return cudaGetLastError();
}

// ---- batched decomposed matrix-vector product -----------------------------------------------------
// out[ct][key][col] -= sum_rows u[ct][row] * key[row][col]          (u64 wrapping, exact)
// CTA tile: GEMV_TB ciphertexts x 256 columns (one per thread, coalesced 64-bit key loads), rows streamed
// in chunks of GEMV_RC with the tile's digits staged in shared memory.  One key element feeds GEMV_TB
// multiply-accumulates (IMAD.WIDE.U32 + IMAD each), so key traffic is 8 B per 32 MACs.
// Grid order: the ciphertext tile is the FASTEST index, so CTAs that are resident together read the
// same key chunk and all but the first hit L2; HBM sees each key byte about once per pass.
// The u64 product sum is kept as two independent accumulators so that each multiply-accumulate costs one
// IMAD.WIDE.U32 + one IMAD on the multiplier pipe (the 64-bit add of the low partial product goes to the
// ALU pipe as IADD3 / IADD3.X):
//   lo64 += u * k_lo          (u < 2^16, k_lo < 2^32, <= 2^13 rows: never overflows 64 bits)
//   hi32 += u * k_hi          (mod 2^32)
//   result = lo64 + (hi32 << 32)   (mod 2^64, exact)
#define GEMV_TB 32
#define GEMV_RC 256
#define GEMV_THREADS 256
__global__ void __launch_bounds__(GEMV_THREADS, 2) gemv_kernel(GemvArgs a) {
    __shared__ __align__(16) uint16_t sdig[GEMV_TB][GEMV_RC];
    const int tid = threadIdx.x;
    const int ct0 = blockIdx.x * GEMV_TB;
    const int col = blockIdx.y * GEMV_THREADS + tid;
    const int nsplit = (a.rows + a.rows_per_split - 1) / a.rows_per_split;
    const int keyi = blockIdx.z / nsplit;
    const int row_begin = (blockIdx.z % nsplit) * a.rows_per_split;
    const int row_end = min(a.rows, row_begin + a.rows_per_split);
    const bool col_ok = col < a.ncols;
    const uint64_t *key = a.key + (size_t)keyi * a.key_stride + (col_ok ? col : 0);
    uint64_t acc_lo[GEMV_TB];
    uint32_t acc_hi[GEMV_TB];
#pragma unroll
    for (int b = 0; b < GEMV_TB; b++) { acc_lo[b] = 0; acc_hi[b] = 0; }
    for (int r0 = row_begin; r0 < row_end; r0 += GEMV_RC) {
        const int nr = min(GEMV_RC, row_end - r0);
        __syncthreads();
        for (int i = tid; i < GEMV_TB * GEMV_RC / 4; i += GEMV_THREADS) {  // 4 digits (8 B) per access
            const int b = i / (GEMV_RC / 4), rr = (i % (GEMV_RC / 4)) * 4;
            const int ct = ct0 + b;
            uint2 v = make_uint2(0, 0);
            if (ct < a.count) {
                const uint16_t *src = a.digits + (size_t)ct * a.rows + r0 + rr;
                if (rr + 4 <= nr && ((reinterpret_cast<uintptr_t>(src) & 7) == 0)) v = *reinterpret_cast<const uint2 *>(src);
                else {
                    uint32_t d[4];
                    for (int t = 0; t < 4; t++) d[t] = (rr + t < nr) ? src[t] : 0;
                    v = make_uint2(d[0] | (d[1] << 16), d[2] | (d[3] << 16));
                }
            }
            *reinterpret_cast<uint2 *>(&sdig[b][rr]) = v;
        }
        __syncthreads();
        const uint64_t *kp = key + (size_t)r0 * a.key_row_stride;
        for (int rr = 0; rr < nr; rr += 4) {   // rows beyond nr carry zero digits; key reads are clamped
            uint32_t klo[4], khi[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint64_t k = __ldg(kp + (size_t)min(rr + t, nr - 1) * a.key_row_stride);
                klo[t] = (uint32_t)k;
                khi[t] = (uint32_t)(k >> 32);
            }
#pragma unroll
            for (int b = 0; b < GEMV_TB; b++) {
                const uint2 dd = *reinterpret_cast<const uint2 *>(&sdig[b][rr]);
                const uint32_t u0 = dd.x & 0xFFFF, u1 = dd.x >> 16, u2 = dd.y & 0xFFFF, u3 = dd.y >> 16;
                acc_lo[b] += (uint64_t)u0 * klo[0];
                acc_lo[b] += (uint64_t)u1 * klo[1];
                acc_lo[b] += (uint64_t)u2 * klo[2];
                acc_lo[b] += (uint64_t)u3 * klo[3];
                acc_hi[b] += u0 * khi[0] + u1 * khi[1] + u2 * khi[2] + u3 * khi[3];
            }
        }
    }
    if (col_ok) {
#pragma unroll
        for (int b = 0; b < GEMV_TB; b++) {
            const int ct = ct0 + b;
            if (ct < a.count)
                atomicAdd(reinterpret_cast<unsigned long long *>(a.out + (size_t)ct * a.out_stride + (size_t)keyi * a.ncols + col),
                          (unsigned long long)(0 - (acc_lo[b] + ((uint64_t)acc_hi[b] << 32))));
        }
    }
}
cudaError_t launch_gemv(const GemvArgs &a, cudaStream_t s) {
    const int coltiles = (a.ncols + GEMV_THREADS - 1) / GEMV_THREADS;
    const int cttiles = (a.count + GEMV_TB - 1) / GEMV_TB;
    const int splits = (a.rows + a.rows_per_split - 1) / a.rows_per_split;
    dim3 grid(cttiles, coltiles, a.nkeys * splits);
    gemv_kernel<<<grid, GEMV_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

// column sums of a key (once per key at load time)
__global__ void key_colsum_kernel(const uint64_t *__restrict__ key, int rows, int ncols, int row_stride, size_t key_stride,
                                  uint64_t *__restrict__ sums) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    const uint64_t *k = key + (size_t)blockIdx.y * key_stride + col;
    uint64_t s = 0;
    for (int r = 0; r < rows; r++) s += k[(size_t)r * row_stride];
    sums[(size_t)blockIdx.y * ncols + col] = s;
}
cudaError_t launch_key_colsum(const uint64_t *key, int rows, int ncols, int row_stride, int nkeys, size_t key_stride,
                              uint64_t *sums, cudaStream_t s) {
    dim3 grid((ncols + 127) / 128, nkeys);
    key_colsum_kernel<<<grid, 128, 0, s>>>(key, rows, ncols, row_stride, key_stride, sums);
    return cudaGetLastError();
}
__global__ void gemv_init_kernel(uint64_t *out, int out_stride, int total_cols, long total, const uint64_t *colsum, uint64_t offset,
                                 const uint64_t *body_src, int body_src_stride, int body_src_index, int body_dst_col) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long b = idx / total_cols;
    const int c = (int)(idx % total_cols);
    uint64_t v = offset * colsum[c];
    if (body_src && c == body_dst_col) v += body_src[b * body_src_stride + body_src_index];
    out[b * out_stride + c] = v;
}
cudaError_t launch_gemv_init(uint64_t *out, int out_stride, int total_cols, int count, const uint64_t *colsum, uint64_t offset,
                             const uint64_t *body_src, int body_src_stride, int body_src_index, int body_dst_col, cudaStream_t s) {
    long total = (long)count * total_cols;
    gemv_init_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(out, out_stride, total_cols, total, colsum, offset, body_src,
                                                                    body_src_stride, body_src_index, body_dst_col);
    return cudaGetLastError();
}

// ---- fused linear layer: dst = sum of up to 5 encrypted bytes (wrapping u64 adds), 128-bit accesses
__global__ void lwe_sum_kernel(const SumEntry *__restrict__ entries, int unit_words) {
    const SumEntry e = entries[blockIdx.y];
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < unit_words; w += gridDim.x * blockDim.x) {
        uint64_t s = e.src[0][w];
        for (int t = 1; t < e.nsrc; t++) s += e.src[t][w];
        e.dst[w] = s;
    }
}
cudaError_t launch_lwe_sum(const SumEntry *entries, int nentries, int unit_words, cudaStream_t s) {
    dim3 grid((unit_words + 1023) / 1024, nentries);
    lwe_sum_kernel<<<grid, 256, 0, s>>>(entries, unit_words);
    return cudaGetLastError();
}

// ---- small elementwise helpers used by the general extract_bits loop --------------------------------
__global__ void scale_kernel(const uint64_t *in, uint64_t *out, long n, uint64_t mul) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] * mul;
}
cudaError_t launch_scale_lwe(const uint64_t *in, uint64_t *out, long n, uint64_t mul, cudaStream_t s) {
    scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n, mul);
    return cudaGetLastError();
}
__global__ void sub_kernel(uint64_t *inout, const uint64_t *sub, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inout[i] -= sub[i];
}
cudaError_t launch_sub_lwe(uint64_t *inout, const uint64_t *sub, long n, cudaStream_t s) {
    sub_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(inout, sub, n);
    return cudaGetLastError();
}
__global__ void add_body_kernel(uint64_t *lwe, int words, int count, uint64_t add) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) lwe[(size_t)i * words + words - 1] += add;
}
cudaError_t launch_add_body(uint64_t *lwe, int words, int count, uint64_t add, cudaStream_t s) {
    add_body_kernel<<<(count + 255) / 256, 256, 0, s>>>(lwe, words, count, add);
    return cudaGetLastError();
}
__global__ void fill_kernel(uint64_t *dst, long n, uint64_t v) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}
cudaError_t launch_fill_u64(uint64_t *dst, long n, uint64_t v, cudaStream_t s) {
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, n, v);
    return cudaGetLastError();
}
__global__ void tree_leaves_kernel(const uint64_t *lut, size_t js, size_t os, int nouts, int nleaf, int kp1, uint64_t *out) {
    // grid: x over words of one GLWE, y = leaf, z = job*nouts + out
    const int job = blockIdx.z / nouts, o = blockIdx.z % nouts, leaf = blockIdx.y;
    const int gsz = kp1 * 512;
    uint64_t *dst = out + ((size_t)blockIdx.z * nleaf + leaf) * gsz;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < gsz; w += gridDim.x * blockDim.x)
        dst[w] = (w >= (kp1 - 1) * 512) ? lut[job * js + o * os + (size_t)leaf * 512 + (w - (kp1 - 1) * 512)] : 0;
}
cudaError_t launch_tree_leaves(const uint64_t *lut, size_t js, size_t os, int njobs, int nouts, int nleaf, int glwe_dim,
                               uint64_t *out, cudaStream_t s) {
    dim3 grid(((glwe_dim + 1) * 512 + 255) / 256, nleaf, njobs * nouts);
    tree_leaves_kernel<<<grid, 256, 0, s>>>(lut, js, os, nouts, nleaf, glwe_dim + 1, out);
    return cudaGetLastError();
}

// ---- add_scalar LUTs (server.rs:181-196 for the low byte, :225-248 for the others) ------------------
// luts: [nblk][nbits+1 polys][512]: polys 0..7 = bits of the sum, poly 8 = carry (bit 0 of the carry LUT);
// for nbits == 8 (low byte) the layout is the same with 8-bit inputs repeated over 512 entries.
__global__ void add_scalar_luts_kernel(const uint64_t *ctr, int byte_index, int nbits, uint64_t *luts) {
    const int blk = blockIdx.x;
    const uint64_t lo = ctr[2 * blk], hi = ctr[2 * blk + 1];
    const int sh = 8 * (15 - byte_index);
    const uint32_t ib = (uint32_t)((sh >= 64 ? hi >> (sh - 64) : lo >> sh) & 0xFF);
    for (int idx = threadIdx.x; idx < 512; idx += blockDim.x) {
        const uint32_t x = (nbits == 8) ? (idx & 0xFF) : idx;
        const uint32_t sum = (x & 0xFF) + ((nbits == 9) ? ((x >> 8) & 1) : 0) + ib;
        for (int b = 0; b < 8; b++) luts[((size_t)blk * 9 + b) * 512 + idx] = (uint64_t)((sum >> b) & 1) << 63;
        luts[((size_t)blk * 9 + 8) * 512 + idx] = (uint64_t)(sum > 255 ? 1 : 0) << 63;
    }
}
cudaError_t launch_add_scalar_luts(const uint64_t *ctr, int nblk, int byte_index, int nbits, uint64_t *luts, cudaStream_t s) {
    add_scalar_luts_kernel<<<nblk, 256, 0, s>>>(ctr, byte_index, nbits, luts);
    return cudaGetLastError();
}
