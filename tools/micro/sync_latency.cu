// Development aid: latency of the inter-CTA reduction / barrier mechanisms the persistent Gauss-Newton kernel can use
// (cooperative launch, G CTAs x 512 threads, one reduction of 58 values per iteration), and of thread-block clusters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/micro/sync_latency tools/micro/sync_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while(0)

constexpr int kThreads = 512;
constexpr int kCols = 58;
constexpr int kSlots = 3;
constexpr int kMaxSpin = 1 << 22;

struct Ctl
{
    unsigned long long bar;          // monotonic barrier word (modes 0/1/2)
    unsigned long long pad[31];
    float rows[2][256][64];          // per-CTA partial rows (mode 2)
    // ring of accumulator slots (modes 3..6): column c of slot s at acc[s][c * stride]
    double acc[kSlots][kCols * 160];
    unsigned arrive[kSlots][64];
    unsigned long long fx[kSlots][4 * 2 * kCols * 32];   // fixed-point words (modes 7/8): word w of slot s at fx[s][w * 32] (256 B apart)
    int errors;
};

__device__ __forceinline__ void red_release_add(unsigned * p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_relaxed_add(unsigned * p, unsigned v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned * p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned * p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double ld_relaxed_f64(const double * p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_f64(double * p, double v) { asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double atom_f64(double * p, double v) { double o; asm volatile("atom.relaxed.gpu.global.add.f64 %0, [%1], %2;" : "=d"(o) : "l"(p), "d"(v) : "memory"); return o; }

// mode 0: barrier only (red.release + ld.acquire, the round-1 group barrier)
// mode 1: barrier only, relaxed
// mode 2: publish a 64-float row, barrier (release/acquire), every CTA folds the G rows (round 1)
// mode 3: 58 fp64 red + fence + arrival + poll + load totals (stride 160 doubles = 1280 B between columns)
// mode 4: 58 fp64 atom (return = ack) + arrival + poll + load totals
// mode 5: mode 3 with the columns packed (stride 1)
// mode 6: mode 4 with the columns packed
__global__ void __launch_bounds__(kThreads, 1) k_sync(Ctl * ctl, int mode, int iters, long long * cycles, int step0)
{
    __shared__ float total[64];
    __shared__ float red[32 * 64];
    const int G = gridDim.x, rank = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned target = (unsigned)(ctl->bar >> 32);
    __syncthreads();
    cg::this_grid().sync();
    const int stride = (mode == 5 || mode == 6) ? 1 : 160;
    long long t0 = clock64();
    int errors = 0;
    for(int it = 0; it < iters; it++)
    {
        const int step = step0 + it;
        const float mine = (float)((rank + 1) * ((threadIdx.x & 63) + 1));
        if(mode <= 1)
        {
            __syncthreads();
            if(threadIdx.x == 0)
            {
                unsigned * hi = reinterpret_cast<unsigned *>(&ctl->bar) + 1;
                target += G;
                if(mode == 0)
                {
                    red_release_add(hi, 1u);
                    int spin = 0;
                    while((int)(ld_acquire(hi) - target) < 0 && ++spin < kMaxSpin) {}
                }
                else
                {
                    red_relaxed_add(hi, 1u);
                    int spin = 0;
                    while((int)(ld_relaxed(hi) - target) < 0 && ++spin < kMaxSpin) {}
                }
            }
            __syncthreads();
        }
        else if(mode == 2)
        {
            float * rows = &ctl->rows[step & 1][0][0];
            if(threadIdx.x < 64) rows[rank * 64 + threadIdx.x] = mine;
            __syncthreads();
            if(threadIdx.x == 0)
            {
                unsigned * hi = reinterpret_cast<unsigned *>(&ctl->bar) + 1;
                target += G;
                red_release_add(hi, 1u);
                int spin = 0;
                while((int)(ld_acquire(hi) - target) < 0 && ++spin < kMaxSpin) {}
            }
            __syncthreads();
            // fold: 16 float4 per row, every thread up to 8 loads in flight
            const int nvec = G * 16;
            float4 v[8];
#pragma unroll
            for(int m = 0; m < 8; m++)
            {
                const int q = threadIdx.x + m * kThreads;
                v[m] = q < nvec ? __ldcg(reinterpret_cast<const float4 *>(rows) + q) : make_float4(0, 0, 0, 0);
            }
            float4 a = v[0];
#pragma unroll
            for(int m = 1; m < 8; m++) { a.x += v[m].x; a.y += v[m].y; a.z += v[m].z; a.w += v[m].w; }
            reinterpret_cast<float4 *>(red)[threadIdx.x] = a;
            __syncthreads();
            if(threadIdx.x < 64)
            {
                float s = 0;
                for(int k = 0; k < 32; k++) s += red[k * 64 + threadIdx.x];
                total[threadIdx.x] = s;
            }
            __syncthreads();
            if(threadIdx.x < 58 && total[threadIdx.x] != (float)((threadIdx.x + 1) * (G * (G + 1) / 2))) errors++;
        }
        else if(mode >= 7)
        {
            // two 64-bit fixed-point words per column (integer part | fraction * 2^48), each carrying the arrival count in its low
            // 8 bits: ONE atomic per word publishes data and arrival together, the readers poll the words themselves
            const int slot = step % kSlots;
            const int nwords = (mode == 8 || mode == 9) ? kCols : 2 * kCols;
            const int wstride = mode == 9 ? 4 : 32;
            const int R = mode == 10 ? 4 : (mode == 11 ? 2 : 1);      // replicas of every word (CTA r adds to replica r % R)
            const int delay = mode == 12 ? 600 : 0;                  // cycles before the first poll
            unsigned long long * fx = ctl->fx[slot];
            if(rank == G - 1 && wid == 15)
            {
                const int nslot = (step + 1) % kSlots;
                for(int c = lane; c < nwords * R; c += 32) ctl->fx[nslot][c * wstride] = 0ull;
            }
            __syncthreads();   // stands for the CTA-level reduce
            if(threadIdx.x < nwords * R)
            {
                const int w = threadIdx.x % nwords, rep = threadIdx.x / nwords;
                const int col = w % kCols;
                if(rep == rank % R)
                {
                    const float p = (float)((rank + 1) * (col + 1)) + 0.25f;
                    const float pi = rintf(p);
                    const float pf = p - pi;
                    const long long q = w < kCols ? __float2ll_rn(pi) : __double2ll_rn((double)pf * 281474976710656.0);
                    const unsigned long long add = ((unsigned long long)q << 8) + 1ull;
                    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(fx + threadIdx.x * wstride), "l"(add) : "memory");
                }
                if(delay)
                {
                    const long long tw = clock64();
                    while(clock64() - tw < delay) {}
                }
                const unsigned want = (unsigned)((G - rep + R - 1) / R);
                unsigned long long v;
                int spin = 0;
                do
                {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(fx + threadIdx.x * wstride) : "memory");
                } while((unsigned)(v & 0xffull) != want && ++spin < kMaxSpin);
                const double val = (double)((long long)v >> 8) * (w < kCols ? 1.0 : 1.0 / 281474976710656.0);
                reinterpret_cast<double *>(red)[threadIdx.x] = val;
            }
            __syncthreads();
            if(threadIdx.x < kCols)
            {
                double t = 0;
                for(int rep = 0; rep < R; rep++)
                    t += reinterpret_cast<double *>(red)[rep * nwords + threadIdx.x] + (nwords == 2 * kCols ? reinterpret_cast<double *>(red)[rep * nwords + threadIdx.x + kCols] : 0.0);
                if(nwords == kCols) t += 0.25 * G;
                total[threadIdx.x] = (float)t;
            }
            __syncthreads();
            if(threadIdx.x < 58 && total[threadIdx.x] != (float)((threadIdx.x + 1) * (G * (G + 1) / 2) + 0.25 * G)) errors++;
        }
        else
        {
            const int slot = step % kSlots;
            double * acc = ctl->acc[slot];
            unsigned * arr = &ctl->arrive[slot][0];
            // housekeeping off the critical path: the last warp of the last CTA clears the slot of step + 1
            if(rank == G - 1 && wid == 15)
            {
                const int nslot = (step + 1) % kSlots;
                for(int c = lane; c < kCols; c += 32) ctl->acc[nslot][c * stride] = 0.0;
                if(lane == 0) ctl->arrive[nslot][0] = 0u;
                if(mode == 4 || mode == 6) __threadfence();
            }
            __syncthreads();   // stands for the CTA-level reduce
            if(wid == 0)
            {
                const double a0 = (double)((rank + 1) * (lane + 1));
                const double a1 = (double)((rank + 1) * (lane + 33));
                if(mode == 3 || mode == 5)
                {
                    red_f64(acc + lane * stride, a0);
                    if(lane + 32 < kCols) red_f64(acc + (lane + 32) * stride, a1);
                    __syncwarp();
                    if(lane == 0)
                    {
                        asm volatile("fence.acq_rel.gpu;" ::: "memory");
                        red_relaxed_add(arr, 1u);
                    }
                }
                else
                {
                    double o0 = atom_f64(acc + lane * stride, a0);
                    double o1 = 0;
                    if(lane + 32 < kCols) o1 = atom_f64(acc + (lane + 32) * stride, a1);
                    // consume the returns so the arrival cannot be issued before them
                    const bool odd = (__double_as_longlong(o0) ^ __double_as_longlong(o1)) == 0x7ff0dead00000001ll;
                    const bool any_odd = __any_sync(0xffffffffu, odd);   // every lane's returns have arrived
                    if(lane == 0) red_relaxed_add(arr, any_odd ? 2u : 1u);
                }
                if(lane == 0)
                {
                    int spin = 0;
                    while(ld_relaxed(arr) < (unsigned)G && ++spin < kMaxSpin) {}
                }
                __syncwarp();
                const double t0v = ld_relaxed_f64(acc + lane * stride);
                const double t1v = (lane + 32 < kCols) ? ld_relaxed_f64(acc + (lane + 32) * stride) : 0.0;
                total[lane] = (float)t0v;
                total[lane + 32] = (float)t1v;
            }
            __syncthreads();
            if(threadIdx.x < 58 && total[threadIdx.x] != (float)((threadIdx.x + 1) * (G * (G + 1) / 2))) errors++;
        }
    }
    long long t1 = clock64();
    if(errors) atomicAdd(&ctl->errors, errors);
    if(rank == 0 && threadIdx.x == 0)
    {
        cycles[0] = t1 - t0;
        if(mode <= 2) ctl->bar = (unsigned long long)target << 32;
    }
    if(mode <= 2 && rank != 0 && threadIdx.x == 0) { /* all CTAs end with the same target */ }
}

// ---- clusters: all-reduce of a 64-float row through distributed shared memory
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ float ld_dsmem(const float * local, unsigned rank)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(local), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

__global__ void __launch_bounds__(kThreads, 1) k_cluster(int iters, long long * cycles, int * errors)
{
    __shared__ float row[2][64];
    __shared__ float total[64];
    const unsigned cr = cluster_ctarank(), cn = cluster_nctarank();
    cluster_arrive();
    cluster_wait();
    int err = 0;
    long long t0 = clock64();
    for(int it = 0; it < iters; it++)
    {
        if(threadIdx.x < 64) row[it & 1][threadIdx.x] = (float)((cr + 1) * (threadIdx.x + 1));
        cluster_arrive();
        cluster_wait();
        if(threadIdx.x < 64)
        {
            float v[16];
#pragma unroll
            for(int r = 0; r < 16; r++) v[r] = r < (int)cn ? ld_dsmem(&row[it & 1][threadIdx.x], r) : 0.f;
            float s = 0;
#pragma unroll
            for(int r = 0; r < 16; r++) s += v[r];
            total[threadIdx.x] = s;
        }
        __syncthreads();
        if(threadIdx.x < 58 && total[threadIdx.x] != (float)((threadIdx.x + 1) * (cn * (cn + 1) / 2))) err++;
    }
    long long t1 = clock64();
    cluster_arrive();
    cluster_wait();
    if(err) atomicAdd(errors, err);
    if(blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

// dependent-chain latencies of fp64 operations (one warp)
__global__ void k_fp64(double * out, long long * cyc, double a, double b)
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
    double q = a;
#pragma unroll 1
    for(int i = 0; i < 64; i++) { q = __shfl_sync(0xffffffffu, q, (threadIdx.x + 1) & 31); q = __dadd_rn(q, b); }
    long long t7 = clock64();
    float g = (float)a;
#pragma unroll 1
    for(int i = 0; i < 64; i++) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(g)); g = r + (float)b; }
    long long t8 = clock64();
    // independent DFMAs (issue rate): 8 chains
    double c0 = a, c1 = a + 1, c2 = a + 2, c3 = a + 3, c4 = a + 4, c5 = a + 5, c6 = a + 6, c7 = a + 7;
#pragma unroll 1
    for(int i = 0; i < 64; i++)
    {
        c0 = __fma_rn(c0, b, a); c1 = __fma_rn(c1, b, a); c2 = __fma_rn(c2, b, a); c3 = __fma_rn(c3, b, a);
        c4 = __fma_rn(c4, b, a); c5 = __fma_rn(c5, b, a); c6 = __fma_rn(c6, b, a); c7 = __fma_rn(c7, b, a);
    }
    long long t9 = clock64();
    out[threadIdx.x] = x + y + z + w + s + f + q + g + c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
    if(threadIdx.x == 0)
    {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; cyc[7] = t8 - t7; cyc[8] = t9 - t8;
    }
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s: %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    Ctl * ctl;
    long long * cyc;
    CK(cudaMalloc(&ctl, sizeof(Ctl)));
    CK(cudaMemset(ctl, 0, sizeof(Ctl)));
    CK(cudaMalloc(&cyc, 256));
    const int iters = 200;
    const char * names[] = {"barrier release/acquire", "barrier relaxed", "rows + barrier + fold (round 1)", "fp64 red + fence + arrive + poll + load (spread)",
                            "fp64 atom-ack + arrive + poll + load (spread)", "fp64 red + fence ... (packed)", "fp64 atom-ack ... (packed)",
                            "fixed-point u64 x2 per column, count embedded (116 words)", "fixed-point u64, 58 words", "fixed-point u64, 58 words, 32 B apart", "fixed-point x2, 4 replicas", "fixed-point x2, 2 replicas", "fixed-point x2, first poll after 600 cycles"};
    for(int G : {148, 74, 38, 16, 8})
    {
        if(G > prop.multiProcessorCount) continue;
        for(int mode = 0; mode < 13; mode++)
        {
            CK(cudaMemset(ctl, 0, sizeof(Ctl)));
            long long best = 1ll << 60;
            int step0 = 0;
            for(int rep = 0; rep < 3; rep++)
            {
                void * args[] = {&ctl, &mode, (void *)&iters, &cyc, &step0};
                CK(cudaLaunchCooperativeKernel((const void *)k_sync, dim3(G), dim3(kThreads), args, 0, 0));
                CK(cudaDeviceSynchronize());
                long long h;
                CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
                if(h < best) best = h;
                step0 += iters;
            }
            int err;
            CK(cudaMemcpy(&err, &ctl->errors, 4, cudaMemcpyDeviceToHost));
            printf("G=%3d mode %d %-52s %8.0f cycles/iteration  errors %d\n", G, mode, names[mode], (double)best / iters, err);
        }
    }
    // cooperative launch + cluster dimension together (148 CTAs)
    for(int cs : {2, 4})
    {
        CK(cudaMemset(ctl, 0, sizeof(Ctl)));
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(kThreads);
        cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs);
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative;
        at[1].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        int mode = 3, it = iters, step0 = 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_sync, ctl, mode, it, cyc, step0);
        if(e == cudaSuccess) e = cudaDeviceSynchronize();
        long long h = 0;
        if(e == cudaSuccess) CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("cooperative + cluster %d, grid %d: %s, mode 3: %.0f cycles/iteration\n", cs, cfg.gridDim.x, cudaGetErrorString(e), (double)h / iters);
        cudaGetLastError();
    }
    // clusters
    int * d_err;
    CK(cudaMalloc(&d_err, 4));
    for(int cs : {2, 4, 8, 16})
    {
        CK(cudaMemset(d_err, 0, 4));
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(kThreads);
        cfg.gridDim = dim3(cs);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if(cs > 8)
        {
            cudaError_t e = cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if(e != cudaSuccess) { printf("cluster %d: non-portable size not allowed: %s\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        }
        int nclusters = 0;
        cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs);
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_cluster, &cfg);
        printf("cluster size %2d: max active clusters %d (%s)\n", cs, nclusters, cudaGetErrorString(e));
        cudaGetLastError();
        cfg.gridDim = dim3(cs);
        int it = iters;
        e = cudaLaunchKernelEx(&cfg, k_cluster, it, cyc, d_err);
        if(e != cudaSuccess) { printf("cluster %d: launch failed: %s\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        CK(cudaDeviceSynchronize());
        e = cudaLaunchKernelEx(&cfg, k_cluster, it, cyc, d_err);
        CK(cudaDeviceSynchronize());
        long long h;
        int err;
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
        printf("cluster size %2d: DSMEM all-reduce of 64 floats: %8.0f cycles/iteration  errors %d\n", cs, (double)h / iters, err);
    }
    // fp64
    double * out;
    CK(cudaMalloc(&out, 32 * 8));
    for(int r = 0; r < 2; r++) k_fp64<<<1, 32>>>(out, cyc, 1.25, 0.75);
    CK(cudaDeviceSynchronize());
    long long h[9];
    CK(cudaMemcpy(h, cyc, 72, cudaMemcpyDeviceToHost));
    printf("latency per op (cycles): dfma %.1f  dadd %.1f  ddiv+dadd %.1f  dsqrt+dadd %.1f  sin+cos+dadd %.1f  ffma %.1f  shfl64+dadd %.1f  rcp.approx+fadd %.1f  | 8 independent dfma: %.1f per dfma\n",
           h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 256.0, h[6] / 64.0, h[7] / 64.0, h[8] / 512.0);
    return 0;
}
