// fp64 pipe throughput per SM on B200: W warps, each 8 independent DFMA / DSETP / MUFU.RSQ64H chains.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double *out, long long *cyc, double y) {
    double x[8];
    for (int u = 0; u < 8; ++u) x[u] = 1.0 + threadIdx.x * 1e-3 + u;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) x[u] = __fma_rn(x[u], y, y);
            if (OP == 1) x[u] = x[u] < y ? __longlong_as_double(__double_as_longlong(x[u]) + 1) : y;
            if (OP == 2) x[u] = __dsqrt_rn(x[u]) + y;
            if (OP == 3) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[u])); x[u] = r + y; }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int u = 0; u < 8; ++u) s += x[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char *name, int warps) {
    double *o; long long *c, h;
    cudaMalloc(&o, 8 * 1024 * 148); cudaMalloc(&c, 8);
    k<OP><<<148, warps * 32>>>(o, c, 1.0000001); k<OP><<<148, warps * 32>>>(o, c, 1.0000001);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-22s %2d warps/SM: %7.2f cycles per warp-op  -> %6.2f warp-ops/cycle/SM\n", name, warps, (double)h / (256 * 8), warps * 256.0 * 8 / h);
    cudaFree(o); cudaFree(c);
}
int main() {
    for (int w : {1, 4, 8, 16, 32}) run<0>("DFMA", w);
    for (int w : {1, 4, 8, 16, 32}) run<1>("DSETP+sel+IADD", w);
    for (int w : {1, 4, 16, 32}) run<2>("DSQRT+DADD", w);
    for (int w : {1, 4, 16, 32}) run<3>("MUFU.RSQ64H+DADD", w);
    return 0;
}
