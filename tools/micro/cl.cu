// Microbenchmark (round 2): communication floor of the cluster-based team engine on B200.
//   1. how many clusters of C CTAs (512 threads, ~200 KB dynamic shared memory) are co-resident, C = 1, 2, 4, 8, 16
//   2. one-way latency of a 16-byte flag-in-data record pushed into a peer CTA's shared memory (DSMEM) and polled locally
//   3. the two-hop skeleton of one pivot: CTA 0 broadcasts a 5-word record through L2 (R replicas), every CTA pushes a
//      5-word record into its cluster leader's shared memory, the leader reduces and posts one record per cluster through
//      L2, every CTA (or only the leaders, then DSMEM) collects all cluster records and CTA 0 starts the next round.
//   4. cost of one gpu-scope fence after a handful of global stores
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d: %s\n",#x,__LINE__,cudaGetErrorString(e)); return 1;}}while(0)
__device__ __forceinline__ unsigned long long gt(){unsigned long long t; asm volatile("mov.u64 %0,%%globaltimer;":"=l"(t)); return t;}
__device__ __forceinline__ int4 ld_g4(const int4* p){int4 v; asm volatile("ld.relaxed.gpu.global.v4.s32 {%0,%1,%2,%3},[%4];":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"l"(p):"memory"); return v;}
__device__ __forceinline__ void st_g4(int4* p, int4 v){asm volatile("st.relaxed.gpu.global.v4.s32 [%0],{%1,%2,%3,%4};"::"l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}
__device__ __forceinline__ int4 ld_s4(const int4* p){int4 v; asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3},[%4];":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"r"((unsigned)__cvta_generic_to_shared(p)):"memory"); return v;}
__device__ __forceinline__ unsigned mapa(const void* p, int rank){unsigned a=(unsigned)__cvta_generic_to_shared(p), r; asm volatile("mapa.shared::cluster.u32 %0,%1,%2;":"=r"(r):"r"(a),"r"(rank)); return r;}
__device__ __forceinline__ void st_c4(unsigned a, int4 v){asm volatile("st.volatile.shared::cluster.v4.s32 [%0],{%1,%2,%3,%4};"::"r"(a),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}
__device__ __forceinline__ int crank(){int r; asm volatile("mov.u32 %0,%%cluster_ctarank;":"=r"(r)); return r;}
__device__ __forceinline__ int csize(){int r; asm volatile("mov.u32 %0,%%cluster_nctarank;":"=r"(r)); return r;}
__device__ __forceinline__ void csync(){asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n":::"memory");}

constexpr int kT = 512;

__global__ void __launch_bounds__(kT,1) probe_kernel(int* out){ extern __shared__ int dyn[]; if(threadIdx.x==0){dyn[0]=blockIdx.x; atomicAdd(out,1);} }

// DSMEM ping-pong between rank 0 and rank 1 of every cluster
__global__ void __launch_bounds__(kT,1) dsmem_pp(int iters, unsigned long long* out)
{
    __shared__ int4 box;
    const int r = crank();
    if (threadIdx.x == 0) box = make_int4(0,0,0,0);
    csync();
    if (threadIdx.x == 0 && r < 2) {
        const unsigned peer = mapa(&box, r ^ 1);
        const unsigned long long t0 = gt();
        for (int i = 1; i <= iters; ++i) {
            if (r == 0) { st_c4(peer, make_int4(i,i,i,i)); int4 v; do v = ld_s4(&box); while (v.w != i); }
            else { int4 v; do v = ld_s4(&box); while (v.w != i); st_c4(peer, make_int4(i,i,i,i)); }
        }
        if (blockIdx.x == 0) out[0] = gt() - t0;
    }
    csync();
}

// the two-hop pivot skeleton.  mode 0: every CTA collects the cluster records from L2; mode 1: only leaders do and forward over DSMEM
// ent: [2][R][8] words; cyc: [2][R][NCL][8] words
__global__ void __launch_bounds__(kT,1) pivot_skel(int4* ent, int4* cyc, int R, int iters, int mode, int work, unsigned long long* out, int* sink)
{
    __shared__ int4 slot[16][5];        // leader: records pushed by the members
    __shared__ int4 rec[64][5];         // collected cluster records
    __shared__ int4 e_in[5];
    __shared__ volatile int wsink;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    const int C = csize(), r = crank(), NCL = gridDim.x / C, cid = cta / C;
    for (int i = tid; i < 16 * 5; i += kT) slot[i / 5][i % 5] = make_int4(0,0,0,0);
    for (int i = tid; i < 64 * 5; i += kT) rec[i / 5][i % 5] = make_int4(0,0,0,0);
    csync();
    const unsigned lead_slot = mapa(&slot[r][0], 0);
    unsigned long long t0 = 0, tA = 0, tB = 0, tm = 0, c0 = 0;
    if (cta == 0 && tid == 0) { t0 = gt(); c0 = tm = clock64(); }
    int acc = 0;
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        // ---- hop 1: CTA 0 -> all
        if (cta == 0 && tid < 5 * R) st_g4(ent + ((size_t)par * R + tid / 5) * 8 + tid % 5, make_int4(acc, it, tid, it));
        if (warp == 0 && lane < 5) { int4 v; const int4* p = ent + ((size_t)par * R + cta % R) * 8 + lane; do v = ld_g4(p); while (v.w != it); e_in[lane] = v; }
        __syncthreads();
        int x = e_in[0].x;
        if (cta == 0 && tid == 0) { const unsigned long long t = clock64(); tA += t - tm; tm = t; }
        // ---- "scan": some dependent work
        for (int i = 0; i < work; ++i) x = x * 1664525 + 1013904223;
        if (x == 0x7fffffff) wsink = x;
        __syncthreads();
        // ---- hop 2: member -> leader (DSMEM), leader -> L2, all collect
        if (mode == 2) {
            if (warp == 0 && lane < 5 * R) st_g4(cyc + (((size_t)par * R + lane / 5) * gridDim.x + cta) * 8 + lane % 5, make_int4(cta, lane, x & 0, it));
            int s2 = 0;
            for (int q = tid; q < (int)gridDim.x; q += kT) {
                const int4* p = cyc + (((size_t)par * R + cta % R) * gridDim.x + q) * 8; int4 v[5];
                for (;;) { bool ok = true; for (int w = 0; w < 5; ++w) { v[w] = ld_g4(p + w); ok = ok && v[w].w == it; } if (ok) break; }
                s2 += v[0].x;
            }
            for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            if (lane == 0 && s2) atomicAdd((int*)&wsink, s2);
            __syncthreads();
            acc += wsink;
            if (cta == 0 && tid == 0) { const unsigned long long t = clock64(); tB += t - tm; tm = t; }
            continue;
        }
        if (warp == 0 && lane < 5) st_c4(lead_slot + lane * 16, make_int4(cta, lane, x & 0, it));
        if (r == 0 && warp == 1) {
            int s = 0;
            for (int q = lane; q < C * 5; q += 32) { int4 v; do v = ld_s4(&slot[q / 5][q % 5]); while (v.w != it); if (q % 5 == 0) s += v.x; }
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane < 5 * R) st_g4(cyc + (((size_t)par * R + lane / 5) * NCL + cid) * 8 + lane % 5, make_int4(s, lane, 0, it));
        }
        if (mode == 3) {
            if (cta == 0) for (int q = tid; q < NCL * 5; q += kT) { int4 v; const int4* p = cyc + (((size_t)par * R) * NCL + q / 5) * 8 + q % 5; do v = ld_g4(p); while (v.w != it); rec[q / 5][q % 5] = v; }
            __syncthreads();
        } else if (mode == 0) {
            for (int q = tid; q < NCL * 5; q += kT) { int4 v; const int4* p = cyc + (((size_t)par * R + cta % R) * NCL + q / 5) * 8 + q % 5; do v = ld_g4(p); while (v.w != it); rec[q / 5][q % 5] = v; }
            __syncthreads();
        } else {
            if (r == 0) {
                for (int q = tid; q < NCL * 5; q += kT) {
                    int4 v; const int4* p = cyc + (((size_t)par * R + cid % R) * NCL + q / 5) * 8 + q % 5; do v = ld_g4(p); while (v.w != it);
                    rec[q / 5][q % 5] = v;
                    for (int m = 1; m < C; ++m) st_c4(mapa(&rec[q / 5][q % 5], m), v);
                }
            } else {
                for (int q = tid; q < NCL * 5; q += kT) { int4 v; do v = ld_s4(&rec[q / 5][q % 5]); while (v.w != it); }
            }
            __syncthreads();
        }
        int s = 0;
        for (int q = lane; q < NCL; q += 32) s += rec[q][0].x;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        acc += s;
        if (cta == 0 && tid == 0) { const unsigned long long t = clock64(); tB += t - tm; tm = t; }
    }
    if (tid == 0) sink[cta] = acc;
    if (cta == 0 && tid == 0) { out[0] = gt() - t0; const double f = (double)out[0] / (double)(clock64() - c0); out[1] = (unsigned long long)(tA * f); out[2] = (unsigned long long)(tB * f); }
    csync();
}

__global__ void fence_kernel(int* buf, int iters, int nst, unsigned long long* out)
{
    if (threadIdx.x) return;
    const long long c0 = clock64();
    for (int i = 0; i < iters; ++i) { for (int k = 0; k < nst; ++k) buf[(i * 37 + k * 1031) & 0xfffff] = i; __threadfence(); }
    out[0] = clock64() - c0;
}

static cudaError_t launch_cl(const void* fn, int grid, int C, size_t smem, void** args, bool coop)
{
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2]; int na = 0;
    at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = C; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na;
    if (coop) { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; ++na; }
    cfg.attrs = at; cfg.numAttrs = na;
    return cudaLaunchKernelExC(&cfg, fn, args);
}

int main()
{
    unsigned long long* out; int* sink; int4 *ent, *cyc;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 4096)); CK(cudaMalloc(&ent, 2 * 8 * 8 * 16)); CK(cudaMalloc(&cyc, (size_t)2 * 8 * 160 * 8 * 16)); 
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, smem optin %zu\n", prop.name, prop.multiProcessorCount, (size_t)prop.sharedMemPerBlockOptin);
    // 1. co-resident clusters
    for (size_t smem : {(size_t)100 * 1024, (size_t)200 * 1024, (size_t)224 * 1024}) {
        CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        for (int C : {1, 2, 4, 8, 16}) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(C * 200); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int ncl = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, probe_kernel, &cfg);
            // largest cooperative launch that is accepted
            int ok = 0;
            for (int g = ncl > 0 ? ncl : 1; g >= 1; --g) {
                CK(cudaMemset(sink, 0, 4)); void* args[] = {&sink};
                cudaError_t le = launch_cl((const void*)probe_kernel, g * C, C, smem, args, true);
                if (le == cudaSuccess) le = cudaDeviceSynchronize();
                if (le == cudaSuccess) { ok = g; break; }
                cudaGetLastError();
            }
            printf("smem %3zu KB cluster %2d: max active clusters %d (%s) = %d CTAs; cooperative launch ok up to %d clusters\n", smem >> 10, C, ncl, cudaGetErrorString(e), ncl * C, ok);
        }
    }
    const int iters = 20000;
    // 2. DSMEM one-way latency
    for (int C : {2, 4, 8}) {
        int it = iters; void* args[] = {&it, &out};
        CK(launch_cl((const void*)dsmem_pp, C, C, 0, args, false)); CK(cudaDeviceSynchronize());
        unsigned long long ns; CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
        printf("DSMEM ping-pong cluster %d: %.3f us round trip, %.3f us one way\n", C, ns / 1000.0 / iters, ns / 2000.0 / iters);
    }
    // 3. pivot skeleton
    for (int C : {2, 4, 8}) for (int ncl : {8, 15, 32, 33, 74}) for (int R : {1, 4}) for (int mode : {0, 1, 2, 3}) {
        const int G = C * ncl;
        if (G > 148 || ncl > 64) continue; if (mode == 3 && R != 1) continue;
        CK(cudaMemset(ent, 0, 2 * 8 * 8 * 16)); CK(cudaMemset(cyc, 0, (size_t)2 * 8 * 160 * 8 * 16));
        int it = iters, rr = R, md = mode, work = 0; void* args[] = {&ent, &cyc, &rr, &it, &md, &work, &out, &sink};
        cudaError_t e = launch_cl((const void*)pivot_skel, G, C, 0, args, true);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("skeleton C=%d ncl=%d R=%d mode=%d: %s\n", C, ncl, R, mode, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        unsigned long long ns[3]; CK(cudaMemcpy(ns, out, 24, cudaMemcpyDeviceToHost));
        printf("skeleton C=%d clusters=%2d (%3d CTAs) R=%d mode=%d: %.3f us/round (hop1 %.3f, hop2 %.3f)\n", C, ncl, G, R, mode, ns[0] / 1000.0 / iters, ns[1] / 1000.0 / iters, ns[2] / 1000.0 / iters);
    }
    // 4. fence
    { int* buf; CK(cudaMalloc(&buf, 4 << 20));
      for (int nst : {0, 1, 8}) { int it = 20000; fence_kernel<<<1, 32>>>(buf, it, nst, out); CK(cudaDeviceSynchronize());
        unsigned long long c; CK(cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost)); printf("fence.gpu after %d stores: %.0f cycles\n", nst, (double)c / it); } }
    return 0;
}
