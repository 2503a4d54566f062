// Dependent-chain latencies of the fp64 operations the Eikonal update is made of (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP> __global__ void k(double *out, long long *cyc, double x0, double y0) {
    double x = x0 + threadIdx.x, y = y0;
    unsigned long long xb = __double_as_longlong(x), yb = __double_as_longlong(y);
    __shared__ volatile double sm[64];
    sm[threadIdx.x & 63] = x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 8; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) x = __dadd_rn(x, y);
            if (OP == 1) x = __dmul_rn(x, y);
            if (OP == 2) x = __fma_rn(x, y, y);
            if (OP == 3) x = x < y ? x + 0.0 * 0 : y, y = __longlong_as_double(__double_as_longlong(y) ^ (__double_as_longlong(x) & 1));   // DSETP + select chain
            if (OP == 4) x = __dsqrt_rn(x) + y;
            if (OP == 5) { xb = xb < yb ? xb : yb; yb ^= (xb & 1); }                 // 64-bit integer min chain
            if (OP == 6) { sm[threadIdx.x & 63] = x; x = sm[(threadIdx.x + 1) & 63] + 0.0; }   // STS -> LDS neighbour round trip (+ DADD)
            if (OP == 7) x = __shfl_up_sync(0xffffffffu, x, 1);
            if (OP == 8) x = fmin(x, y) + 1e-300;
        }
    }
    long long t1 = clock64();
    if (OP == 5) x = __longlong_as_double(xb ^ yb);
    out[threadIdx.x] = x + y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char *name, double x0, double y0) {
    double *o; long long *c, h;
    cudaMalloc(&o, 8 * 64); cudaMalloc(&c, 8);
    k<OP><<<1, 32>>>(o, c, x0, y0); k<OP><<<1, 32>>>(o, c, x0, y0);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %7.1f cycles per op (1 warp)\n", name, (double)h / N);
    cudaFree(o); cudaFree(c);
}
int main() {
    run<0>("DADD", 1.0, 1e-9); run<1>("DMUL", 1.0, 1.0000001); run<2>("DFMA", 1.0, 0.5);
    run<3>("DSETP+sel", 5.0, 3.0); run<4>("DSQRT+DADD", 2.0, 1.0); run<5>("int64 min", 5.0, 3.0);
    run<6>("STS->LDS+DADD", 1.0, 1.0); run<7>("SHFL.UP (64-bit)", 1.0, 1.0); run<8>("fmin+DADD", 5.0, 3.0);
    return 0;
}
