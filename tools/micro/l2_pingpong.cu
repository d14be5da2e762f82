// Development aid: the floor of any inter-CTA handshake through L2 on B200 -- single-thread round trips and a two-CTA ping-pong.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while(0)

__device__ __forceinline__ unsigned ld_relaxed(const unsigned * p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_acquire(const unsigned * p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_volatile(const unsigned * p) { unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed(unsigned * p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_release(unsigned * p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add(unsigned * p, unsigned v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned atom_add(unsigned * p, unsigned v) { unsigned o; asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o; }

// single thread: [0] ld.relaxed RTT, [1] atom RTT, [2] red then poll until visible, [3] st.relaxed then poll until visible, [4] fence.acq_rel.gpu after a store
__global__ void k_self(unsigned * w, long long * out)
{
    const int N = 64;
    long long t0 = clock64();
    unsigned s = 0;
    for(int i = 0; i < N; i++) s += ld_relaxed(w + ((s & 1) ? 0 : 0));
    long long t1 = clock64();
    for(int i = 0; i < N; i++) s += atom_add(w + 64 + (s & 0), 1u);
    long long t2 = clock64();
    unsigned base = ld_relaxed(w + 128);
    for(int i = 0; i < N; i++)
    {
        red_add(w + 128, 1u);
        base++;
        while(ld_relaxed(w + 128) != base) {}
    }
    long long t3 = clock64();
    for(int i = 0; i < N; i++)
    {
        st_relaxed(w + 192, (unsigned)i + 1000u);
        while(ld_relaxed(w + 192) != (unsigned)i + 1000u) {}
    }
    long long t4 = clock64();
    for(int i = 0; i < N; i++)
    {
        st_relaxed(w + 256, (unsigned)i);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    long long t5 = clock64();
    out[0] = (t1 - t0) / N; out[1] = (t2 - t1) / N; out[2] = (t3 - t2) / N; out[3] = (t4 - t3) / N; out[4] = (t5 - t4) / N;
    out[7] = s;
}

// two CTAs (on different SMs) bounce a counter: mode 0 st.relaxed + ld.relaxed, 1 red + ld.relaxed, 2 atom + ld.relaxed, 3 st.release + ld.acquire, 4 st.relaxed + ld.volatile
__global__ void k_pingpong(unsigned * w, int mode, int iters, long long * out)
{
    if(threadIdx.x != 0) return;
    unsigned * mine = w + (blockIdx.x == 0 ? 0 : 64);     // the word this CTA writes
    unsigned * theirs = w + (blockIdx.x == 0 ? 64 : 0);   // the word it polls
    long long t0 = clock64();
    for(int i = 1; i <= iters; i++)
    {
        if(blockIdx.x == 0)
        {
            if(mode == 0 || mode == 4) st_relaxed(mine, (unsigned)i);
            else if(mode == 1) red_add(mine, 1u);
            else if(mode == 2) atom_add(mine, 1u);
            else st_release(mine, (unsigned)i);
        }
        int spin = 0;
        if(mode == 3) { while(ld_acquire(theirs) < (unsigned)i && ++spin < (1 << 22)) {} }
        else if(mode == 4) { while(ld_volatile(theirs) < (unsigned)i && ++spin < (1 << 22)) {} }
        else { while(ld_relaxed(theirs) < (unsigned)i && ++spin < (1 << 22)) {} }
        if(blockIdx.x == 1)
        {
            if(mode == 0 || mode == 4) st_relaxed(mine, (unsigned)i);
            else if(mode == 1) red_add(mine, 1u);
            else if(mode == 2) atom_add(mine, 1u);
            else st_release(mine, (unsigned)i);
        }
    }
    long long t1 = clock64();
    if(blockIdx.x == 0) out[0] = (t1 - t0) / iters;
}

int main()
{
    unsigned * w;
    long long * out;
    CK(cudaMalloc(&w, 4096));
    CK(cudaMalloc(&out, 64));
    long long h[8];
    for(int r = 0; r < 2; r++)
    {
        CK(cudaMemset(w, 0, 4096));
        k_self<<<1, 1>>>(w, out);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost));
    printf("single thread (cycles): ld.relaxed RTT %lld, atom RTT %lld, red->visible to own poll %lld, st->visible to own poll %lld, st + fence.acq_rel.gpu %lld\n", h[0], h[1], h[2], h[3], h[4]);
    const char * names[] = {"st.relaxed / ld.relaxed", "red / ld.relaxed", "atom / ld.relaxed", "st.release / ld.acquire", "st.relaxed / ld.volatile"};
    for(int mode = 0; mode < 5; mode++)
    {
        for(int r = 0; r < 2; r++)
        {
            CK(cudaMemset(w, 0, 4096));
            k_pingpong<<<2, 32>>>(w, mode, 200, out);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
        printf("ping-pong %-28s %lld cycles per round trip (two one-way messages)\n", names[mode], h[0]);
    }
    return 0;
}
