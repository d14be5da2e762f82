// Development aid: cost of a loop body that does not fit the SM's instruction caches, for ONE warp per SM (the situation of the
// solving warp of the persistent Gauss-Newton kernel) and for 16 warps.  Body = N x 16-byte FFMA instructions, 4 independent chains.
#include <cstdio>
#include <cuda_runtime.h>
#define F4 asm volatile("fma.rn.f32 %0, %0, %4, %5;\n\tfma.rn.f32 %1, %1, %4, %5;\n\tfma.rn.f32 %2, %2, %4, %5;\n\tfma.rn.f32 %3, %3, %4, %5;" : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3) : "f"(a), "f"(b));
#define F16 F4 F4 F4 F4
#define F64 F16 F16 F16 F16
#define F256 F64 F64 F64 F64
#define F1K F256 F256 F256 F256
template <int KB>   // body size in KB (1 KB = 64 instructions)
__global__ void k_loop(float * out, long long * cyc, int iters, float a, float b)
{
    float x0 = 1.f, x1 = 2.f, x2 = 3.f, x3 = 4.f;
    long long t0 = 0;
    for(int it = 0; it < iters + 1; it++)
    {
        if(it == 1) t0 = clock64();   // first pass warms
#pragma unroll
        for(int k = 0; k < KB / 16; k++) { F1K }
        if(KB % 16 >= 8) { F256 F256 }
        if(KB % 8 >= 4) { F256 }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
    if(blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int KB> void run(float * out, long long * cyc)
{
    for(int warps : {1, 16})
    {
        const int iters = 20;
        k_loop<KB><<<148, 32 * warps>>>(out, cyc, iters, 0.999f, 0.001f);
        cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("body %3d KB, %2d warps/SM: %.3f cycles per instruction (%.0f cycles per pass)\n", KB, warps, (double)h / iters / (KB * 64), (double)h / iters);
    }
}
int main()
{
    float * out; long long * cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    run<4>(out, cyc); run<8>(out, cyc); run<16>(out, cyc); run<20>(out, cyc); run<24>(out, cyc); run<28>(out, cyc); run<32>(out, cyc); run<40>(out, cyc); run<48>(out, cyc); run<64>(out, cyc); run<96>(out, cyc); run<128>(out, cyc);
    return 0;
}
