// Development aid: dependent-chain latency of the fp64 operations the Gauss-Newton bookkeeping uses (one warp, one CTA).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double * out, long long * cyc, double a, double b)
{
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for(int i = 0; i < 64; i++) { x = __fma_rn(x, b, a); x = __fma_rn(x, b, a); x = __fma_rn(x, b, a); x = __fma_rn(x, b, a); }
    long long t1 = clock64();
    double y = a;
#pragma unroll 1
    for(int i = 0; i < 64; i++) { y = __dadd_rn(y, b); y = __dadd_rn(y, b); y = __dadd_rn(y, b); y = __dadd_rn(y, b); }
    long long t2 = clock64();
    double z = a;
#pragma unroll 1
    for(int i = 0; i < 64; i++) { z = __ddiv_rn(1.0, z); z = __dadd_rn(z, b); }
    long long t3 = clock64();
    double w = a;
#pragma unroll 1
    for(int i = 0; i < 64; i++) { w = sqrt(w) + b; }
    long long t4 = clock64();
    double s = a;
#pragma unroll 1
    for(int i = 0; i < 64; i++) { s = sin(s) + cos(s); }
    long long t5 = clock64();
    float f = (float)a;
#pragma unroll 1
    for(int i = 0; i < 64; i++) { f = __fmaf_rn(f, (float)b, (float)a); f = __fmaf_rn(f, (float)b, (float)a); f = __fmaf_rn(f, (float)b, (float)a); f = __fmaf_rn(f, (float)b, (float)a); }
    long long t6 = clock64();
    out[threadIdx.x] = x + y + z + w + s + f;
    if(threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; }
}
int main()
{
    double * out; long long * cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 64);
    for(int r = 0; r < 2; r++) k<<<1, 32>>>(out, cyc, 1.25, 0.75);
    long long h[6];
    cudaMemcpy(h, cyc, 48, cudaMemcpyDeviceToHost);
    printf("dfma %.1f  dadd %.1f  ddiv+dadd %.1f  dsqrt+dadd %.1f  sin+cos+dadd %.1f  ffma %.1f cycles per op\n", h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 256.0);
    return 0;
}
