// device check of the (2^8,5) byte recoding of cmux_core.cuh against the host evaluation of the same code
#include <cstdio>
#include <random>
#include <vector>
#include "../tfhe-aes_b200/csrc/cmux_core.cuh"
__global__ void k(const uint64_t *x, double *d, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t st;
    d[5 * i] = decomp85_first(x[i], st) * 1.5;
    for (int lev = 4; lev >= 1; lev--) d[5 * i + 5 - lev] = decomp85_level(st, lev) * 1.5;
}
__global__ void kt(const double *v, uint64_t *o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = f64_to_torus(v[i]);
}
static int check_torus() {
    const int n = 1 << 18;
    std::mt19937_64 g(9);
    std::vector<double> v(n);
    for (int i = 0; i < n; i++) v[i] = ldexp((double)(int64_t)g() / 9223372036854775808.0, (int)(g() % 92));
    v[0] = 0.5; v[1] = -0.5; v[2] = 2147483648.0; v[3] = -2147483648.5; v[4] = 2147483647.5; v[5] = -0.0;
    double *dv; uint64_t *dout;
    cudaMalloc(&dv, n * 8); cudaMalloc(&dout, n * 8);
    cudaMemcpy(dv, v.data(), n * 8, cudaMemcpyHostToDevice);
    kt<<<n / 256, 256>>>(dv, dout, n);
    std::vector<uint64_t> o(n);
    cudaMemcpy(o.data(), dout, n * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; i++) if (o[i] != f64_to_torus(v[i])) bad++;
    printf("f64_to_torus device vs host: %d mismatches\n", bad);
    return bad;
}
int main() {
    if (check_torus()) return 1;
    const int n = 1 << 16;
    std::mt19937_64 g(3);
    std::vector<uint64_t> x(n);
    for (auto &v : x) v = g();
    uint64_t *dx; double *dd;
    cudaMalloc(&dx, n * 8); cudaMalloc(&dd, n * 40);
    cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, dd, n);
    std::vector<double> d(5 * n);
    cudaError_t e = cudaMemcpy(d.data(), dd, n * 40, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; i++) {
        uint32_t st; double h[5];
        h[0] = decomp85_first(x[i], st);
        for (int lev = 4; lev >= 1; lev--) h[5 - lev] = decomp85_level(st, lev);
        for (int j = 0; j < 5; j++) if (h[j] * 1.5 != d[5 * i + j]) { if (bad < 5) printf("x=%llx j=%d host %f dev %f\n", (unsigned long long)x[i], j, h[j], d[5 * i + j] / 1.5); bad++; }
    }
    printf("decomp85 device vs host: %d mismatches (%s)\n", bad, cudaGetErrorString(e));
    return bad != 0;
}
