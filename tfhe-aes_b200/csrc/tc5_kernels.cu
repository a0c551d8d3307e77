// tc5_kernels.cu — private functional packing keyswitch (K4; circuit_bootstrap_boolean,
// many_wopbs.rs:253) as an EXACT integer GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8,
// accumulators in tensor memory, operands staged by TMA), sm_100a.
//
//   out[bit][key][col] = - sum_rows d[bit][row] * PFPKSK[key][row][col]          (mod 2^64)
//
// Batched over the bits of all resident ciphertexts: [bits x rows] . [rows x cols], rows = (kN+1)*l_pfks = 6147,
// cols = (k+1)*(k+1)*N = 12 800 at PARAM_OPT.  The arithmetic is the limb scheme of imma_kernels.cu:
//   key = sum_{b<8} 2^(8b) kb   (u8 limbs),   d = dl + 128 dh   (s8 limbs),
//   sum_rows d * key = sum_b 2^(8b) ( sum_rows dl kb  +  128 sum_rows dh kb ),
// every inner sum an s32 tensor-core accumulation (|.| <= 6147 * 64 * 255 < 2^27), recombined in u64 in the epilogue.
// What changes is the machine mapping:
//   * B operand = the key exactly as it lies in memory.  A u64 key matrix [row][col] IS a u8 matrix
//     [row][8*col + limb] with the N index contiguous, i.e. an "MN-major" B operand; no key re-layout, no second
//     copy of the 630 MB key.
//   * A operand = the digit limb planes [bit][rows_pad] (K contiguous, "K-major"), one plane per limb; the two
//     planes accumulate into two TMEM accumulators that share the TMEM lanes, so that the epilogue finds the
//     16 partial sums of one output in one thread.
//   * CTA tile: 128 bits x 256 byte-columns (= 32 u64 columns) x all rows; 2 accumulators x 256 columns = the
//     whole TMEM of the SM.  Per 128-row stage: A 2 x 16 KB + B 32 KB = 64 KB, 3 stages.
//   * warp 0: TMA producer (one lane), warp 1: tcgen05.mma issuer (one lane) + TMEM allocation,
//     warps 2-5: epilogue (tcgen05.ld 32 lanes each, limb recombination, 16-byte stores).
// Shared-memory operand layouts are the canonical 128-byte-swizzle layouts of the tensor-core descriptors
// (K-major: rows of 128 B, 8-row groups 1 KB apart; MN-major: 128 B of N per K row, 8-row groups 1 KB apart,
// next 128 B of N one stage-block further), written by TMA with CU_TENSOR_MAP_SWIZZLE_128B.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.h"

#define T5_M 128                 // bits per CTA
#define T5_NB 256                // byte-columns per CTA (32 u64 columns x 8 limbs)
#define T5_KS 128                // key rows per pipeline stage
#define T5_THREADS 192
#define T5_TILE_BYTES (T5_M * 128)                 // one 128-row x 128-byte operand block = 16 KB
// DL = digit limbs (2: PFKS, digits up to 2^11; 1: LWE keyswitch, |digit| <= 2).  Stage = DL A blocks + 2 B blocks.
#define T5_STAGE_BYTES(DL) (((DL) + 2) * T5_TILE_BYTES)
#define T5_STAGES(DL) ((DL) == 2 ? 3 : 4)

__device__ __forceinline__ unsigned t5_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t5_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t5_smem(bar)), "r"(count));
}
__device__ __forceinline__ void t5_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(t5_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t5_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "T5_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni T5_WAIT_DONE;\n\t"
        "bra.uni T5_WAIT_LOOP;\n\t"
        "T5_WAIT_DONE:\n\t}" ::"r"(t5_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void t5_tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     t5_smem(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(t5_smem(bar))
                 : "memory");
}
// shared-memory matrix descriptor (tcgen05): start address, leading / stride byte offsets (all >> 4), version 1, 128 B swizzle
__device__ __forceinline__ uint64_t t5_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void t5_mma_i8(unsigned tmem_d, uint64_t desc_a, uint64_t desc_b, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void t5_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t5_smem(bar)) : "memory");
}

struct Tc5Args {
    const uint64_t *body_src; // keyswitch only: [count][body_stride] input LWEs, word body_index is added to column body_col
    int body_stride, body_index, body_col;
    uint64_t *out;            // [count][out_stride]; the product is written negated (no read-modify-write)
    int out_stride;
    int count;                // bits
    int rows;                 // key rows per key (digit rows actually used)
    int ncols;                // u64 columns per key
    int ntiles_per_key;       // ncols * 8 / T5_NB
    int nk;                   // pipeline stages over the rows: ceil(rows / T5_KS)
};

// instruction descriptor: D = s32, A = s8 (K-major), B = u8 (MN-major), M = 128, N = 256, dense, no saturation
#define T5_IDESC ((2u << 4) | (1u << 7) | (0u << 10) | (0u << 15) | (1u << 16) | ((T5_NB >> 3) << 17) | ((T5_M >> 4) << 24))

template <int DL>
__global__ void __launch_bounds__(T5_THREADS, 1) tc5_keyswitch_kernel(const __grid_constant__ CUtensorMap map_dl, const __grid_constant__ CUtensorMap map_dh,
                                                                      const __grid_constant__ CUtensorMap map_key, Tc5Args a) {
    constexpr int STAGES = T5_STAGES(DL);
    constexpr unsigned STAGE_BYTES = T5_STAGE_BYTES(DL);
    extern __shared__ __align__(1024) unsigned char t5_smem_raw[];
    // [STAGES][STAGE_BYTES]; the swizzled operand blocks need 1 KB alignment in the shared address space
    unsigned char *stages = t5_smem_raw + ((1024u - (t5_smem(t5_smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(stages + STAGES * STAGE_BYTES);         // [STAGES]
    uint64_t *empty = full + STAGES;                                          // [STAGES]
    uint64_t *tmem_full = empty + STAGES;
    unsigned *tmem_ptr = reinterpret_cast<unsigned *>(tmem_full + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * T5_M;
    const int nt = blockIdx.y;                                                   // byte-column tile over all keys
    const int keyi = nt / a.ntiles_per_key;
    const int nb0 = (nt % a.ntiles_per_key) * T5_NB;                             // first byte-column inside the key
    const int row0 = keyi * a.rows;                                              // first key row of this key in the key tensor

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { t5_mbar_init(&full[s], 1); t5_mbar_init(&empty[s], 1); }
        t5_mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: all 512 columns (2 accumulators x 256)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(t5_smem(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (int kc = 0; kc < a.nk; kc++) {
                const int s = kc % STAGES;
                if (kc >= STAGES) t5_mbar_wait(&empty[s], ((kc / STAGES) - 1) & 1);
                unsigned char *st = stages + (size_t)s * STAGE_BYTES;
                t5_mbar_expect_tx(&full[s], STAGE_BYTES);
                t5_tma_load_2d(st, &map_dl, kc * T5_KS, m0, &full[s]);                                          // A_lo: [bit][128 rows]
                if (DL == 2) t5_tma_load_2d(st + T5_TILE_BYTES, &map_dh, kc * T5_KS, m0, &full[s]);             // A_hi
                t5_tma_load_2d(st + DL * T5_TILE_BYTES, &map_key, nb0, row0 + kc * T5_KS, &full[s]);              // B bytes nb0 .. +127
                t5_tma_load_2d(st + (DL + 1) * T5_TILE_BYTES, &map_key, nb0 + 128, row0 + kc * T5_KS, &full[s]);  // B bytes nb0+128 ..
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            for (int kc = 0; kc < a.nk; kc++) {
                const int s = kc % STAGES;
                t5_mbar_wait(&full[s], (kc / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned st = t5_smem(stages + (size_t)s * STAGE_BYTES);
#pragma unroll
                for (int ks = 0; ks < T5_KS / 32; ks++) {
                    // K-major A: one MMA consumes 32 bytes of every 128-byte row; MN-major B: 32 rows of 128 bytes
                    const uint64_t da_lo = t5_desc(st + ks * 32, 16, 1024);
                    const uint64_t db = t5_desc(st + DL * T5_TILE_BYTES + ks * 32 * 128, T5_TILE_BYTES, 1024);
                    const unsigned acc = (kc | ks) != 0;
                    t5_mma_i8(tmem_base, da_lo, db, T5_IDESC, acc);
                    if (DL == 2) t5_mma_i8(tmem_base + T5_NB, t5_desc(st + T5_TILE_BYTES + ks * 32, 16, 1024), db, T5_IDESC, acc);
                }
                t5_commit(&empty[s]);            // the stage may be refilled once these MMAs have read it
            }
            t5_commit(tmem_full);                // all accumulations complete
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> u64 recombination -> global ----------------
        t5_mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;                                    // TMEM lanes 32*quarter .. +31 are this warp's
        const int bit = m0 + quarter * 32 + lane;
        const unsigned taddr = tmem_base + ((unsigned)(quarter * 32) << 16);
        const int bitc = min(bit, a.count - 1);
        uint64_t *orow = a.out + (size_t)bitc * a.out_stride + (size_t)keyi * a.ncols + nb0 / 8;
        const uint64_t body = a.body_src ? a.body_src[(size_t)bitc * a.body_stride + a.body_index] : 0;
#pragma unroll 1
        for (int c = 0; c < T5_NB / 16; c++) {                           // 16 byte-columns = 2 u64 outputs per step
            unsigned lo[16], hi[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(lo[0]), "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]), "=r"(lo[4]), "=r"(lo[5]), "=r"(lo[6]), "=r"(lo[7]), "=r"(lo[8]),
                           "=r"(lo[9]), "=r"(lo[10]), "=r"(lo[11]), "=r"(lo[12]), "=r"(lo[13]), "=r"(lo[14]), "=r"(lo[15])
                         : "r"(taddr + c * 16));
            if (DL == 2) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(hi[0]), "=r"(hi[1]), "=r"(hi[2]), "=r"(hi[3]), "=r"(hi[4]), "=r"(hi[5]), "=r"(hi[6]), "=r"(hi[7]), "=r"(hi[8]),
                               "=r"(hi[9]), "=r"(hi[10]), "=r"(hi[11]), "=r"(hi[12]), "=r"(hi[13]), "=r"(hi[14]), "=r"(hi[15])
                             : "r"(taddr + T5_NB + c * 16));
            } else {
#pragma unroll
                for (int e = 0; e < 16; e++) hi[e] = 0;
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint64_t v[2];
#pragma unroll
            for (int o = 0; o < 2; o++) {
                uint64_t acc = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    const int64_t p = (int64_t)(int)lo[8 * o + b] + (int64_t)(int)hi[8 * o + b] * 128;
                    acc += (uint64_t)p << (8 * b);
                }
                const int col = nb0 / 8 + 2 * c + o;
                v[o] = ((a.body_src && col == a.body_col) ? body : (uint64_t)0) - acc;
            }
            // (ncols is even and tiles start at multiples of 32 columns, so a pair is either all inside or all outside)
            if (bit < a.count && nb0 / 8 + 2 * c < a.ncols) *reinterpret_cast<ulonglong2 *>(orow + 2 * c) = make_ulonglong2(v[0], v[1]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point resolved through the runtime, no -lcuda) and launcher
// ------------------------------------------------------------------------------------------------
typedef CUresult (*t5_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static t5_encode_fn t5_get_encode() {
    static t5_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<t5_encode_fn>(p);
    }
    return fn;
}
// u8 matrix [outer][inner_bytes] with row pitch pitch_bytes, boxes of 128 bytes x 128 rows, 128 B swizzle, zero fill
static bool t5_make_map(CUtensorMap *m, const void *base, uint64_t inner_bytes, uint64_t outer, uint64_t pitch_bytes) {
    t5_encode_fn enc = t5_get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {inner_bytes, outer};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {128, 128};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tc5_pfks_supported(int ncols, int rows_pad) { return (ncols * 8) % T5_NB == 0 && rows_pad % 16 == 0; }
bool tc5_ks_supported(int ncols, int key_pitch_cols, int rows_pad) { return ncols % 2 == 0 && key_pitch_cols % 2 == 0 && rows_pad % 16 == 0; }

template <int DL>
static cudaError_t t5_launch(const CUtensorMap &map_dl, const CUtensorMap &map_dh, const CUtensorMap &map_key, const Tc5Args &a, dim3 grid, cudaStream_t s) {
    const size_t smem = (size_t)T5_STAGES(DL) * T5_STAGE_BYTES(DL) + 1024 + 128;   // + alignment slack + barriers
    // per launch, as the PBS and vertical-packing launchers do: the attribute is per device and a process may hold contexts on several
    cudaError_t e = cudaFuncSetAttribute(tc5_keyswitch_kernel<DL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc5_keyswitch_kernel<DL><<<grid, T5_THREADS, smem, s>>>(map_dl, map_dh, map_key, a);
    return cudaGetLastError();
}

// out[bit][key*ncols + col] = - sum_row (dl + 128 dh)[bit][row] * key[key][row][col];  dl/dh [count][rows_pad] (zero padded rows)
cudaError_t launch_tc5_pfks(const int8_t *dl, const int8_t *dh, int rows_pad, const uint64_t *key, int nkeys, int rows, int ncols, int count,
                            uint64_t *out, int out_stride, cudaStream_t s) {
    if (!tc5_pfks_supported(ncols, rows_pad)) return cudaErrorInvalidValue;
    CUtensorMap map_dl, map_dh, map_key;
    if (!t5_make_map(&map_dl, dl, (uint64_t)rows_pad, (uint64_t)count, (uint64_t)rows_pad) ||
        !t5_make_map(&map_dh, dh, (uint64_t)rows_pad, (uint64_t)count, (uint64_t)rows_pad) ||
        !t5_make_map(&map_key, key, (uint64_t)ncols * 8, (uint64_t)nkeys * rows, (uint64_t)ncols * 8))
        return cudaErrorInvalidValue;
    Tc5Args a{};
    a.out = out; a.out_stride = out_stride; a.count = count; a.rows = rows; a.ncols = ncols;
    a.ntiles_per_key = ncols * 8 / T5_NB;
    a.nk = (rows + T5_KS - 1) / T5_KS;
    return t5_launch<2>(map_dl, map_dh, map_key, a, dim3((count + T5_M - 1) / T5_M, a.ntiles_per_key * nkeys), s);
}

// LWE keyswitch: out[bit][col] = (col == body_col ? in[bit][body_index] : 0) - sum_row dl[bit][row] * key[row][col];
// key [rows][key_pitch_cols] u64 (columns >= ncols are padding), one digit limb (|digit| <= 64)
cudaError_t launch_tc5_keyswitch(const int8_t *dl, int rows_pad, const uint64_t *key, int rows, int ncols, int key_pitch_cols, int count,
                                 const uint64_t *in, int in_stride, int body_index, uint64_t *out, int out_stride, cudaStream_t s) {
    if (!tc5_ks_supported(ncols, key_pitch_cols, rows_pad)) return cudaErrorInvalidValue;
    CUtensorMap map_dl, map_key;
    if (!t5_make_map(&map_dl, dl, (uint64_t)rows_pad, (uint64_t)count, (uint64_t)rows_pad) ||
        !t5_make_map(&map_key, key, (uint64_t)key_pitch_cols * 8, (uint64_t)rows, (uint64_t)key_pitch_cols * 8))
        return cudaErrorInvalidValue;
    Tc5Args a{};
    a.body_src = in; a.body_stride = in_stride; a.body_index = body_index; a.body_col = ncols - 1;
    a.out = out; a.out_stride = out_stride; a.count = count; a.rows = rows; a.ncols = ncols;
    a.ntiles_per_key = (ncols * 8 + T5_NB - 1) / T5_NB;
    a.nk = (rows + T5_KS - 1) / T5_KS;
    return t5_launch<1>(map_dl, map_dl, map_key, a, dim3((count + T5_M - 1) / T5_M, a.ntiles_per_key), s);
}
