// fft_core.cuh — register-level building blocks of the 256-point complex FFT that implements the
// negacyclic transform for N = 512 (reference call sites many_wopbs.rs:253,263,277).
//
// Convention (same as tfhe-fft): z[n] = p[n] + i p[n+256]; X[k] = sum_n z[n] theta^(n(4k+1)),
// theta = exp(2 pi i / 1024), i.e. twist exp(i pi n / 512) followed by a 256-point DFT with
// kernel exp(+2 pi i nk / 256).  Pointwise products of such spectra are negacyclic products.
//
// Decomposition ("four-step", 16 x 16): n = 16 n1 + n2, k = k1 + 16 k2.
//   pass 1 (lane n2):  T[k1] = sum_n1 z[16 n1 + n2] phi^(n1) W16^(n1 k1),  phi = exp(2 pi i/64)
//   mid twiddle:       T[k1] *= theta^(n2 (4 k1 + 1))
//   exchange through shared memory (16 x 16 transpose inside one half-warp, XOR-swizzled, no padding)
//   pass 2 (lane k1):  X[k1 + 16 k2] = sum_n2 T[k1][n2] W16^(n2 k2)
// Each of the 16 lanes holds 16 complex values in registers, so a transform costs one shared-memory
// round trip.  Everything here is __host__ __device__ so the CPU emulation in emu.cu runs the very
// same code the kernels run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef HD
#define HD __host__ __device__ __forceinline__
#endif

typedef double2 cd;

#include "w64_table.inc"

HD cd cmk(double x, double y) { cd r; r.x = x; r.y = y; return r; }
HD cd cadd(cd a, cd b) { return cmk(a.x + b.x, a.y + b.y); }
HD cd csub(cd a, cd b) { return cmk(a.x - b.x, a.y - b.y); }
HD cd cmul(cd a, cd b) { return cmk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// acc += a * b
HD void cmac(cd &acc, cd a, cd b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

// v * exp(SIGN * 2 pi i m / 64), m a compile-time constant after unrolling
template <int SIGN>
HD cd mul_w64(cd v, int m) {
    m &= 63;
    if (m == 0) return v;
    if (m == 16) return SIGN > 0 ? cmk(-v.y, v.x) : cmk(v.y, -v.x);
    if (m == 32) return cmk(-v.x, -v.y);
    if (m == 48) return SIGN > 0 ? cmk(v.y, -v.x) : cmk(-v.y, v.x);
    const double c = w64_cos(m), s = SIGN * w64_sin(m);
    return cmk(v.x * c - v.y * s, v.x * s + v.y * c);
}

// radix-4 butterfly, kernel exp(SIGN * 2 pi i /4): (a,b,c,d) -> (y0,y1,y2,y3)
template <int SIGN>
HD void radix4(cd &a, cd &b, cd &c, cd &d) {
    cd t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
    cd it3 = SIGN > 0 ? cmk(-t3.y, t3.x) : cmk(t3.y, -t3.x);
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = cadd(t1, it3);
    d = csub(t1, it3);
}

// position of natural output index k (0..15) inside the register array after fft16
HD constexpr int rev4(int k) { return ((k & 3) << 2) | (k >> 2); }

// 16-point DFT in registers, kernel exp(SIGN * 2 pi i nk / 16).  In: v[n].  Out: X[k] at v[rev4(k)].
template <int SIGN>
HD void fft16(cd (&v)[16]) {
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) radix4<SIGN>(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
#pragma unroll
    for (int n2 = 1; n2 < 4; n2++)
#pragma unroll
        for (int k1 = 1; k1 < 4; k1++) v[n2 + 4 * k1] = mul_w64<SIGN>(v[n2 + 4 * k1], 4 * n2 * k1);
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) radix4<SIGN>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

// a * conj(b)
HD cd cmul_conj(cd a, cd b) { return cmk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }

// Exchange buffer of one 16-lane group: a 16 x 16 matrix of complex values WITHOUT padding, element
// (row, col) stored at row*16 + (col ^ row).  Writing a fixed row from 16 lanes (col = lane) touches the 256
// contiguous bytes of that row; reading a fixed column from 16 lanes (row = lane) touches every 16-byte bank
// group exactly twice: both directions cost the minimum of two 128-byte wavefronts per half-warp.  The same
// storage, read linearly as 256 complex, is also the hand-over format to / from the multiply-accumulate phase.
#define XB_ELEMS 256
HD constexpr int xb_idx(int row, int col) { return row * 16 + (col ^ row); }

// Mid-twiddle table (filled once per CTA, make_twiddle_table): tw[xb_idx(k1, n2)] = theta^(n2 (4 k1 + 1)).
// The forward pass reads it along rows (k1 fixed, lane = n2), the inverse pass along columns (n2 fixed,
// lane = k1) and conjugates; the swizzle keeps both conflict-free, so one 4 KB table serves both.
HD void fft256_fwd_pass1(cd (&v)[16], int lane, const cd *tw, cd *xb) {
#pragma unroll
    for (int n1 = 1; n1 < 16; n1++) v[n1] = mul_w64<1>(v[n1], n1);
    fft16<1>(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) xb[xb_idx(k1, lane)] = cmul(v[rev4(k1)], tw[xb_idx(k1, lane)]);
}
// the same in two halves, so that the arithmetic can run before the exchange buffer is free
HD void fft256_fwd_pass1_compute(cd (&v)[16], int lane, const cd *tw) {
#pragma unroll
    for (int n1 = 1; n1 < 16; n1++) v[n1] = mul_w64<1>(v[n1], n1);
    fft16<1>(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) v[rev4(k1)] = cmul(v[rev4(k1)], tw[xb_idx(k1, lane)]);
}
HD void fft256_fwd_pass1_store(const cd (&v)[16], int lane, cd *xb) {
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) xb[xb_idx(k1, lane)] = v[rev4(k1)];
}
// lane = k1.  Out: X[lane + 16 k2] at v[rev4(k2)]
HD void fft256_fwd_pass2(cd (&v)[16], int lane, const cd *xb) {
#pragma unroll
    for (int n2 = 0; n2 < 16; n2++) v[n2] = xb[xb_idx(lane, n2)];
    fft16<1>(v);
}
// lane = k1.  In: v[k2] = X[lane + 16 k2].  Writes twiddled U[n2] to element (n2, lane).
HD void fft256_inv_pass1_compute(cd (&v)[16]) { fft16<-1>(v); }
HD void fft256_inv_pass1_store(cd (&v)[16], int lane, const cd *tw, cd *xb) {
#pragma unroll
    for (int n2 = 0; n2 < 16; n2++) xb[xb_idx(n2, lane)] = cmul_conj(v[rev4(n2)], tw[xb_idx(lane, n2)]);
}
// batched form (the stores to xb may alias the twiddle table as far as the compiler knows, so it never hoists a twiddle
// load above the previous store: explicit batches of TWB loads, then TWB products and stores)
template <int TWB>
HD void fft256_inv_pass1_store_b(cd (&v)[16], int lane, const cd *tw, cd *xb) {
#pragma unroll
    for (int h = 0; h < 16; h += TWB) {
        cd w[TWB];
#pragma unroll
        for (int j = 0; j < TWB; j++) w[j] = tw[xb_idx(lane, h + j)];
#pragma unroll
        for (int j = 0; j < TWB; j++) xb[xb_idx(h + j, lane)] = cmul_conj(v[rev4(h + j)], w[j]);
    }
}
// lane = n2.  Out: z[16 n1 + lane] * 256 at v[rev4(n1)] before the final untwist; this applies the
// untwist conj(phi^n1) and the 1/256 scale and returns natural order in v[n1].
HD void fft256_inv_pass2(cd (&v)[16], int lane, const cd *xb) {
    cd t[16];
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) t[k1] = xb[xb_idx(lane, k1)];
    fft16<-1>(t);
    // untwist by conj(phi^n1) with the 1/256 scale folded into the constants (a power of two: the products are
    // the same doubles as scaling afterwards)
    v[0] = cmk(t[rev4(0)].x * (1.0 / 256.0), t[rev4(0)].y * (1.0 / 256.0));
#pragma unroll
    for (int n1 = 1; n1 < 16; n1++) {
        const cd x = t[rev4(n1)];
        const double c = w64_cos(n1) * (1.0 / 256.0), s = -w64_sin(n1) * (1.0 / 256.0);
        v[n1] = cmk(x.x * c - x.y * s, x.x * s + x.y * c);
    }
}

// round-to-nearest(-even) f64 -> u64 modulo 2^64 (SURVEY §9.6 "round to nearest, reduce mod 2^64"), valid for
// |v| < 2^93 (bootstrap products stay below 2^84, the level-1 products of vertical packing below 2^89).  No
// conversion instructions (FRND / F2I issue at 1/4 of the FP64 rate on B200, scratch/mb_xu.cu) and four FP64
// operations: with M = 1.5 * 2^52,
//   t = v * 2^-42 + M      -> low word of t = h mod 2^32,  h = rint(v / 2^42)     (only h mod 2^22 is needed)
//   r = v - h * 2^42       -> exact, |r| <= 2^41
//   u = r + M              -> bits(u) - bits(M) = rint(r) as a signed 64-bit integer
//   result = (h << 42) + rint(r) = rint(v)  (mod 2^64)
HD uint64_t f64_to_torus(double v) {
    const double M = 6755399441055744.0;
    const double t = fma(v, 1.0 / 4398046511104.0, M);
    const double h = t - M;
    const double r = fma(-h, 4398046511104.0, v);
    const double u = r + M;
#ifdef __CUDA_ARCH__
    const uint32_t lo = (uint32_t)__double2loint(u);
    const uint32_t hi = ((uint32_t)__double2loint(t) << 10) + (uint32_t)__double2hiint(u) - 0x43380000u;
    return ((uint64_t)hi << 32) | lo;
#else
    uint64_t bt, bu;
    __builtin_memcpy(&bt, &t, 8);
    __builtin_memcpy(&bu, &u, 8);
    return (bt << 42) + (bu - 0x4338000000000000ull);
#endif
}
