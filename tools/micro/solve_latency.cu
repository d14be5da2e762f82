// Development aid: latency of the per-step solve (gn_fast_math.cuh: warp_update_fast) on one warp, cold and warm.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 --ftz=true --prec-div=false --prec-sqrt=false -I include -I slam_b200/csrc
//        -o build/micro/solve_latency tools/micro/solve_latency.cu
#include <cstdio>
#include <cstdlib>
#include "gn_fast_math.cuh"
namespace slam { void set_last_error(const std::string &) {} }
using namespace slam;

__global__ void __launch_bounds__(512, 1) k_solve(long long * cycles, int reps, int evict, const float * junk, float * sink, slam_step_record * recs)
{
    __shared__ GnShared sh;
    if(threadIdx.x == 0)
    {
        LevelGeom g = {480, 640, 481.2f, -480.f, 319.5f, 239.5f};
        float Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tp[3] = {0.1f, 0.2f, 0.3f};
        seq_begin_pose(sh, Rp, tp);
        level_begin(sh, g);
        for(int k = 0; k < 16; k++) sh.resultRt[k] = (k % 5 == 0) ? 1.0 : 0.0;
    }
    __syncthreads();
    if(threadIdx.x < 32) warp_prepare_fast(sh, true);
    __syncthreads();
    for(int r = 0; r < reps; r++)
    {
        if(threadIdx.x < 64)
        {
            // a well conditioned system: J^T J of a few synthetic rows
            const int t = threadIdx.x & 31;
            float v = 0.f;
            if(t < 27)
            {
                int i = 0, rem = t;
                while(rem >= 7 - i) { rem -= 7 - i; i++; }
                const int j = i + rem;
                for(int q = 0; q < 40; q++)
                {
                    const float a = __sinf(0.37f * q + 1.3f * i + 0.01f * r) + (i == (q % 6) ? 2.f : 0.f);
                    const float b = j == 6 ? 1e-3f * __cosf(0.11f * q) : __sinf(0.37f * q + 1.3f * j + 0.01f * r) + (j == (q % 6) ? 2.f : 0.f);
                    v += a * b;
                }
            }
            sh.total[threadIdx.x] = v * (threadIdx.x < 32 ? 1.f : 50.f);
        }
        __syncthreads();
        if(evict)
        {
            // run something else in between, as the map phases do in the real kernel
            float s = 0.f;
            for(int q = threadIdx.x; q < evict; q += blockDim.x) s += junk[q];
            sink[threadIdx.x] = s;
        }
        __syncthreads();
        if(threadIdx.x < 32)
        {
            const long long t0 = clock64();
            warp_update_fast(sh, true, true, 10.f, (recs && threadIdx.x == 0) ? recs + r : nullptr, t0);
            const long long t1 = clock64();
            if(threadIdx.x == 0) cycles[r] = t1 - t0;
        }
        __syncthreads();
    }
    if(threadIdx.x < 9) sink[threadIdx.x] = sh.krk[threadIdx.x];
    if(threadIdx.x < 3) { sink[9 + threadIdx.x] = sh.kt[threadIdx.x]; sink[12 + threadIdx.x] = sh.tcurr[threadIdx.x]; }
    if(threadIdx.x < 9) sink[15 + threadIdx.x] = sh.Rcurr[threadIdx.x];
}

int main()
{
    long long * cyc;
    float * junk, * sink;
    cudaMalloc(&cyc, 64 * 8);
    cudaMalloc(&junk, 1 << 20);
    cudaMalloc(&sink, 4096);
    cudaMemset(junk, 0, 1 << 20);
    slam_step_record * recs;
    cudaMalloc(&recs, sizeof(slam_step_record) * 12);
    for(int evict : {0, 100000})
    {
        k_solve<<<1, 512>>>(cyc, 12, evict, junk, sink, evict ? recs : nullptr);
        cudaError_t e = cudaGetLastError();
        if(e == cudaSuccess) e = cudaDeviceSynchronize();
        if(e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[12];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("warp_update_fast, cycles per call (evict %d):", evict);
        for(int r = 0; r < 12; r++) printf(" %lld", h[r]);
        printf("\n");
        if(evict)
        {
            slam_step_record hr[12];
            cudaMemcpy(hr, recs, sizeof(hr), cudaMemcpyDeviceToHost);
            const slam_step_record & q = hr[8];
            float hs[24];
            cudaMemcpy(hs, sink, sizeof(hs), cudaMemcpyDeviceToHost);
            printf("after 12 steps: krk %g %g %g | %g %g %g | %g %g %g; kt %g %g %g; tcurr %g %g %g; Rcurr[0..2] %g %g %g\n", hs[0], hs[1], hs[2], hs[3], hs[4], hs[5], hs[6], hs[7], hs[8], hs[9],
                   hs[10], hs[11], hs[12], hs[13], hs[14], hs[15], hs[16], hs[17]);
            printf("stages of call 8 (cycles since entry): combined %u, eliminated %u, rotation %u, resultRt %u, parameters %u; x = %g %g %g %g %g %g\n", q.t_solve[0], q.t_solve[1], q.t_solve[2],
                   q.t_solve[3], q.t_solve[4], q.x[0], q.x[1], q.x[2], q.x[3], q.x[4], q.x[5]);
        }
    }
    return 0;
}
