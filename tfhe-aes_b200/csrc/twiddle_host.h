// twiddle_host.h — host-side construction of the 256-point mid-twiddle table used by fft_core.cuh.
#pragma once
#include <cmath>
#include "fft_core.cuh"
// tw[xb_idx(k1, n2)] = theta^(n2 (4 k1 + 1)),  theta = exp(2 pi i / 1024)   (256 entries, swizzled like the
// exchange buffers so that the forward and the inverse pass both read it without bank conflicts)
static inline void make_twiddle_table(cd *tw) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k1 = 0; k1 < 16; k1++)
        for (int n2 = 0; n2 < 16; n2++) {
            int e = (n2 * (4 * k1 + 1)) & 1023;
            long double a = two_pi * e / 1024.0L;
            double c = (double)cosl(a), s = (double)sinl(a);
            if (e == 0) { c = 1.0; s = 0.0; }
            if (e == 256) { c = 0.0; s = 1.0; }
            if (e == 512) { c = -1.0; s = 0.0; }
            if (e == 768) { c = 0.0; s = -1.0; }
            tw[xb_idx(k1, n2)] = cmk(c, s);
        }
}
