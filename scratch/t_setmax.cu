__global__ void __launch_bounds__(512, 1) k(double *out, const double *in, int n) {
    if (threadIdx.x < 256) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        double a[80];
#pragma unroll
        for (int i = 0; i < 80; i++) a[i] = in[threadIdx.x + i * 512];
        for (int it = 0; it < n; it++) {
#pragma unroll
            for (int i = 0; i < 80; i++) a[i] = a[i] * a[(i + 1) % 80] + 1.0;
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 80; i++) s += a[i];
        out[threadIdx.x] = s;
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        out[threadIdx.x] = in[threadIdx.x] * 2.0;
    }
}
