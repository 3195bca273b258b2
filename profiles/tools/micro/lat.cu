// Dependent-issue latencies on sm_100a (cycles per op in a serial chain, one warp): DADD, DFMA, DMUL, FFMA, IADD,
// LDS (pointer chase), SHFL.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 4096;
template <int KIND>
__global__ void k(double *out, long long *cyc, double seed, int *chase) {
    __shared__ int sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = chase[i];
    __syncthreads();
    double a = seed, b = seed * 0.5 + 1.0, c = 1.0000001;
    float fa = (float)seed, fb = 1.0001f;
    int ia = (int)seed, p = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (KIND == 0) a = __dadd_rn(a, b);
        else if (KIND == 1) a = fma(a, c, b);
        else if (KIND == 2) a = __dmul_rn(a, c);
        else if (KIND == 3) fa = fmaf(fa, fb, fb);
        else if (KIND == 4) ia = ia * 3 + p;
        else if (KIND == 5) p = sm[p];
        else if (KIND == 6) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0;
        else if (KIND == 7) { a = fmin(a, b); b = b + 1e-9; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + fa + ia + p;
    if (threadIdx.x == 0) cyc[KIND] = t1 - t0;
}
int main() {
    double *out; long long *cyc; int *chase, h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = (i * 37 + 11) % 1024;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64 * 8); cudaMalloc(&chase, 4096);
    cudaMemcpy(chase, h, 4096, cudaMemcpyHostToDevice);
    k<0><<<1, 32>>>(out, cyc, 1.5, chase); k<1><<<1, 32>>>(out, cyc, 1.5, chase); k<2><<<1, 32>>>(out, cyc, 1.5, chase);
    k<3><<<1, 32>>>(out, cyc, 1.5, chase); k<4><<<1, 32>>>(out, cyc, 1.5, chase); k<5><<<1, 32>>>(out, cyc, 1.5, chase);
    k<6><<<1, 32>>>(out, cyc, 1.5, chase); k<7><<<1, 32>>>(out, cyc, 1.5, chase);
    long long hc[8];
    cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost);
    const char *names[] = {"DADD", "DFMA", "DMUL", "FFMA", "IMAD", "LDS chase", "SHFL+DADD", "DMNMX(+DADD indep)"};
    for (int i = 0; i < 8; ++i) printf("%-20s %.1f cycles/op\n", names[i], (double)hc[i] / N);
    return 0;
}
