// Microbenchmarks for the pivot kernel's building blocks on B200: grid barrier variants, L2 chase, cluster barrier.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d: %s\n",#x,__LINE__,cudaGetErrorString(e)); return 1;}}while(0)
__device__ __forceinline__ unsigned long long ldacq(const unsigned long long* p){unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0,[%1];":"=l"(v):"l"(p):"memory"); return v;}
__device__ __forceinline__ unsigned long long ldrlx(const unsigned long long* p){unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0,[%1];":"=l"(v):"l"(p):"memory"); return v;}
__device__ __forceinline__ void redrel(unsigned long long* p){asm volatile("red.release.gpu.global.add.u64 [%0],1;"::"l"(p):"memory");}
__device__ __forceinline__ void strel(unsigned long long* p, unsigned long long v){asm volatile("st.release.gpu.global.u64 [%0],%1;"::"l"(p),"l"(v):"memory");}
__device__ __forceinline__ unsigned long long gt(){unsigned long long t; asm volatile("mov.u64 %0,%%globaltimer;":"=l"(t)); return t;}

// mode 0: fence + red.release + ld.acquire poll + fence (current)   mode 1: red.release + ld.acquire poll
// mode 2: per-CTA slots, st.release epoch, warp-parallel poll        mode 3: cooperative groups grid.sync
// mode 4: like 1 but poll with relaxed loads then one fence
__global__ void __launch_bounds__(1024,1) bar_kernel(unsigned long long* ctr, unsigned long long* slots, int iters, int mode, unsigned long long* out, int* sink)
{
    cg::grid_group grid = cg::this_grid();
    const int G = gridDim.x;
    unsigned long long target = 0, t0 = 0;
    __shared__ int dummy;
    if (blockIdx.x==0 && threadIdx.x==0) t0 = gt();
    for (int it = 1; it <= iters; ++it) {
        if (mode == 3) { grid.sync(); continue; }
        __syncthreads();
        if (mode == 2) {
            if (threadIdx.x == 0) strel(&slots[blockIdx.x * 16], (unsigned long long)it);
            if (threadIdx.x < G) { while (ldacq(&slots[threadIdx.x * 16]) < (unsigned long long)it) ; }
        } else if (threadIdx.x == 0) {
            target += G;
            if (mode == 0) __threadfence();
            redrel(ctr);
            if (mode == 4) { while (ldrlx(ctr) < target) ; __threadfence(); }
            else while (ldacq(ctr) < target) ;
            if (mode == 0) __threadfence();
        }
        __syncthreads();
    }
    if (blockIdx.x==0 && threadIdx.x==0) { out[0] = gt() - t0; dummy = 1; sink[0] = dummy; }
}
__global__ void chase_kernel(const int* next, int steps, unsigned long long* out, int* sink)
{
    int p = 0; unsigned long long t0 = gt(); long long c0 = clock64();
    for (int i = 0; i < steps; ++i) p = __ldcg(next + p);
    out[0] = gt() - t0; out[1] = clock64() - c0; sink[0] = p;
}
__global__ void __cluster_dims__(16,1,1) __launch_bounds__(1024,1) cluster_kernel16(int iters, unsigned long long* out)
{
    cg::cluster_group cl = cg::this_cluster();
    unsigned long long t0 = gt();
    for (int i = 0; i < iters; ++i) cl.sync();
    if (blockIdx.x==0 && threadIdx.x==0) out[0] = gt() - t0;
}
__global__ void __cluster_dims__(8,1,1) __launch_bounds__(1024,1) cluster_kernel8(int iters, unsigned long long* out)
{
    cg::cluster_group cl = cg::this_cluster();
    unsigned long long t0 = gt();
    for (int i = 0; i < iters; ++i) cl.sync();
    if (blockIdx.x==0 && threadIdx.x==0) out[0] = gt() - t0;
}
// one-way signal latency: CTA0 writes epoch, CTA1 polls and answers (ping-pong) => RT/2 per hop
__global__ void pingpong_kernel(unsigned long long* a, unsigned long long* b, int iters, unsigned long long* out)
{
    if (threadIdx.x) return;
    unsigned long long t0 = gt();
    if (blockIdx.x == 0) { for (int i = 1; i <= iters; ++i) { strel(a, i); while (ldacq(b) < (unsigned long long)i) ; } out[0] = gt() - t0; }
    else if (blockIdx.x == 1) { for (int i = 1; i <= iters; ++i) { while (ldacq(a) < (unsigned long long)i) ; strel(b, i); } }
}
int main()
{
    unsigned long long *ctr, *slots, *out; int* sink;
    CK(cudaMalloc(&ctr, 64)); CK(cudaMalloc(&slots, 148*16*8)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 64));
    const int iters = 20000;
    for (int mode = 0; mode <= 4; ++mode) for (int G : {2, 4, 16, 37, 74, 148}) {
        CK(cudaMemset(ctr, 0, 64)); CK(cudaMemset(slots, 0, 148*16*8));
        int it = iters; void* args[] = {&ctr, &slots, &it, &mode, &out, &sink};
        CK(cudaLaunchCooperativeKernel((void*)bar_kernel, dim3(G), dim3(1024), args, 0, 0));
        CK(cudaDeviceSynchronize());
        unsigned long long ns; CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
        printf("barrier mode %d G=%3d : %.3f us/barrier\n", mode, G, ns / 1000.0 / iters);
    }
    for (size_t bytes : {size_t(1)<<20, size_t(8)<<20, size_t(64)<<20, size_t(512)<<20}) {
        size_t n = bytes / 4; std::vector<int> h(n);
        // random cyclic permutation with stride of at least a line
        std::vector<int> perm(n/32); for (size_t i=0;i<perm.size();++i) perm[i]=(int)i;
        unsigned long long s=88172645463325252ULL; for (size_t i=perm.size()-1;i>0;--i){ s^=s<<13; s^=s>>7; s^=s<<17; size_t j=s%(i+1); std::swap(perm[i],perm[j]); }
        for (size_t i=0;i<perm.size();++i) h[(size_t)perm[i]*32] = perm[(i+1)%perm.size()]*32;
        int* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
        int steps = 20000;
        for (int rep=0; rep<2; ++rep) { chase_kernel<<<1,1>>>(d, steps, out, sink); CK(cudaDeviceSynchronize()); }
        unsigned long long r[2]; CK(cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost));
        printf("chase %4zu MB: %.1f ns/load, %.1f cycles/load\n", bytes>>20, (double)r[0]/steps, (double)r[1]/steps);
        cudaFree(d);
    }
    { int it = iters; cluster_kernel16<<<16,1024>>>(it, out); cudaError_t e = cudaDeviceSynchronize();
      if (e==cudaSuccess) { unsigned long long ns; cudaMemcpy(&ns,out,8,cudaMemcpyDeviceToHost); printf("cluster16 sync: %.3f us\n", ns/1000.0/iters);} else { printf("cluster16 failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); } }
    { int it = iters; cluster_kernel8<<<8,1024>>>(it, out); cudaError_t e = cudaDeviceSynchronize();
      if (e==cudaSuccess) { unsigned long long ns; cudaMemcpy(&ns,out,8,cudaMemcpyDeviceToHost); printf("cluster8 sync: %.3f us\n", ns/1000.0/iters);} else printf("cluster8 failed: %s\n", cudaGetErrorString(e)); }
    { CK(cudaMemset(ctr,0,64)); CK(cudaMemset(slots,0,64)); int it=iters; pingpong_kernel<<<2,32>>>(ctr, slots, it, out); CK(cudaDeviceSynchronize());
      unsigned long long ns; CK(cudaMemcpy(&ns,out,8,cudaMemcpyDeviceToHost)); printf("pingpong round trip: %.3f us\n", ns/1000.0/iters); }
    return 0;
}
