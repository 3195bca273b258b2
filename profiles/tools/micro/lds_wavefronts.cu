// Micro-benchmark: shared-memory wavefronts per instruction for the access shapes of the pruning kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_wavefronts lds_wavefronts.cu
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,gpu__time_duration.sum ./lds_wavefronts
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int KIND>
__global__ void k(double *out, const int *idx) {
    __shared__ __align__(16) double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 0.5;
    __syncthreads();
    unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    const int lane = threadIdx.x & 31;
    const int sel = idx[lane];  // 0..4
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    unsigned off = 0;
#pragma unroll 4
    for (int it = 0; it < ITER; ++it) {
        off = (off + 128) & 16383;  // warp-uniform moving offset
        double a, b, c, d;
        if (KIND == 0) {  // uniform 128-bit
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(base + off));
            acc0 += a; acc1 += b;
        } else if (KIND == 1) {  // uniform 64-bit
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(base + off));
            acc0 += a;
        } else if (KIND == 2) {  // lane-dependent 128-bit among 5 rows of 32 B (tip table rows), first half
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(base + off + sel * 32));
            acc0 += a; acc1 += b;
        } else if (KIND == 3) {  // consecutive bytes
            unsigned r;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(base + off + lane));
            acc0 += r;
        } else if (KIND == 4) {  // lane-dependent 64-bit among 25 entries (cherry plane)
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(base + off + (sel * 5 + (lane & 3)) * 8));
            acc0 += a;
        } else if (KIND == 5) {  // per-lane 256-bit as two 128-bit (stage of 32 messages, 1 KB per warp)
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(base + ((off * 8 + lane * 32) & 32767)));
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(c), "=d"(d) : "r"(base + ((off * 8 + lane * 32 + 16) & 32767)));
            acc0 += a; acc1 += b; acc2 += c; acc3 += d;
        } else if (KIND == 6) {  // uniform 128-bit through a generic pointer
            const volatile double *p = reinterpret_cast<const volatile double *>(reinterpret_cast<const char *>(sm) + off);
            acc0 += p[0]; acc1 += p[1];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

int main() {
    double *out;
    int *idx, h[32];
    for (int i = 0; i < 32; ++i) h[i] = (i * 7 + 3) % 5;
    cudaMalloc(&out, 148 * 256 * 8);
    cudaMalloc(&idx, 128);
    cudaMemcpy(idx, h, 128, cudaMemcpyHostToDevice);
    k<0><<<148, 256>>>(out, idx);
    k<1><<<148, 256>>>(out, idx);
    k<2><<<148, 256>>>(out, idx);
    k<3><<<148, 256>>>(out, idx);
    k<4><<<148, 256>>>(out, idx);
    k<5><<<148, 256>>>(out, idx);
    k<6><<<148, 256>>>(out, idx);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
