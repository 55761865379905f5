// Microbenchmark: cost of one "hop" (small all-to-all / broadcast exchange between persistent CTAs) on B200 for the
// message-passing variants the pivot kernel can use.  Each record is 16 B = {payload[3], seq}; the sequence number
// travels in the same 128-bit store as the data, so no fence is needed (flag-in-data).
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d: %s\n",#x,__LINE__,cudaGetErrorString(e)); return 1;}}while(0)
__device__ __forceinline__ unsigned long long gt(){unsigned long long t; asm volatile("mov.u64 %0,%%globaltimer;":"=l"(t)); return t;}
__device__ __forceinline__ int4 ld_vol4(const int4* p){int4 v; asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3},[%4];":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"l"(p):"memory"); return v;}
__device__ __forceinline__ int4 ld_rlx4(const int4* p){int4 v; asm volatile("ld.relaxed.gpu.global.v4.s32 {%0,%1,%2,%3},[%4];":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"l"(p):"memory"); return v;}
__device__ __forceinline__ void st_vol4(int4* p, int4 v){asm volatile("st.volatile.global.v4.s32 [%0],{%1,%2,%3,%4};"::"l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}
__device__ __forceinline__ void st_rlx4(int4* p, int4 v){asm volatile("st.relaxed.gpu.global.v4.s32 [%0],{%1,%2,%3,%4};"::"l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}

// mode 0: all-to-all, volatile st/ld.  mode 1: all-to-all, relaxed.gpu st/ld.
// mode 2: gather to CTA 0 then broadcast from CTA 0 (2 hops, volatile).  mode 3: broadcast only: CTA 0 writes, all poll, (no return path: CTA0 paces by clock)
// slots: [2][G] records of 16 B, padded to `stride` int4 per record
__global__ void __launch_bounds__(1024,1) a2a_kernel(int4* slots, int stride, int iters, int mode, unsigned long long* out, int* sink)
{
    const int G = gridDim.x, tid = threadIdx.x, cta = blockIdx.x;
    __shared__ int s_acc;
    unsigned long long t0 = 0;
    if (cta==0 && tid==0) t0 = gt();
    int acc = 0;
    for (int it = 1; it <= iters; ++it) {
        int4* base = slots + (size_t)(it & 1) * G * stride;
        if (mode == 0 || mode == 1) {
            if (tid == 0) { int4 v = make_int4(cta, acc, it, it); if (mode==0) st_vol4(base + (size_t)cta*stride, v); else st_rlx4(base + (size_t)cta*stride, v); }
            if (tid < G) { int4 v; do { v = mode==0 ? ld_vol4(base + (size_t)tid*stride) : ld_rlx4(base + (size_t)tid*stride); } while (v.w != it); acc += v.x; }
            __syncthreads();
        } else if (mode == 2) {
            int4* up = base; int4* down = slots + (size_t)(2 + (it & 1)) * G * stride;
            if (cta != 0) {
                if (tid == 0) { st_vol4(up + (size_t)cta*stride, make_int4(cta, acc, it, it)); int4 v; do { v = ld_vol4(down); } while (v.w != it); acc += v.x; }
            } else {
                if (tid > 0 && tid < G) { int4 v; do { v = ld_vol4(up + (size_t)tid*stride); } while (v.w != it); acc += v.x; }
                __syncthreads();
                if (tid == 0) st_vol4(down, make_int4(acc, 0, it, it));
            }
            __syncthreads();
        }
    }
    if (tid == 0) { s_acc = acc; sink[cta] = s_acc; }
    if (cta==0 && tid==0) out[0] = gt() - t0;
}

// hierarchical: clusters of CS CTAs; members write their record into the leader's smem through DSMEM, cluster barrier,
// leaders exchange all-to-all through global memory (volatile), leader publishes result in its smem, cluster barrier, members read it.
template <int CS>
__global__ void __launch_bounds__(1024,1) hier_kernel(int4* slots, int stride, int iters, unsigned long long* out, int* sink)
{
    cg::cluster_group cl = cg::this_cluster();
    const int tid = threadIdx.x, cta = blockIdx.x, rank = cl.block_rank(), NC = gridDim.x / CS, cid = cta / CS;
    __shared__ int4 s_in[16]; __shared__ int4 s_out;
    int4* lead_in = cl.map_shared_rank(s_in, 0);
    int4* lead_out = cl.map_shared_rank(&s_out, 0);
    unsigned long long t0 = 0; if (cta==0 && tid==0) t0 = gt();
    int acc = 0;
    for (int it = 1; it <= iters; ++it) {
        int4* base = slots + (size_t)(it & 1) * NC * stride;
        if (tid == 0) lead_in[rank] = make_int4(cta, acc, it, it);
        cl.sync();
        if (rank == 0) {
            if (tid == 0) { int s = 0; for (int r = 0; r < CS; ++r) s += s_in[r].x; st_vol4(base + (size_t)cid*stride, make_int4(s, 0, it, it)); }
            int part = 0;
            if (tid < NC) { int4 v; do { v = ld_vol4(base + (size_t)tid*stride); } while (v.w != it); part = v.x; }
            for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (tid == 0) s_out = make_int4(part, 0, it, it);
        }
        cl.sync();
        acc += lead_out->x;
    }
    if (tid == 0) sink[cta] = acc;
    if (cta==0 && tid==0) out[0] = gt() - t0;
}

// one-way latency ping-pong, flag in data, volatile / relaxed
__global__ void pingpong_kernel(int4* a, int4* b, int iters, int mode, unsigned long long* out)
{
    if (threadIdx.x) return;
    unsigned long long t0 = gt();
    if (blockIdx.x == 0) { for (int i = 1; i <= iters; ++i) { if (mode==0) st_vol4(a, make_int4(i,i,i,i)); else st_rlx4(a, make_int4(i,i,i,i)); int4 v; do { v = mode==0 ? ld_vol4(b) : ld_rlx4(b); } while (v.w != i); } out[0] = gt() - t0; }
    else if (blockIdx.x == gridDim.x - 1) { for (int i = 1; i <= iters; ++i) { int4 v; do { v = mode==0 ? ld_vol4(a) : ld_rlx4(a); } while (v.w != i); if (mode==0) st_vol4(b, make_int4(i,i,i,i)); else st_rlx4(b, make_int4(i,i,i,i)); } }
}

// gather latency: 1024 threads each load K random 8-byte words (L2 resident array of `n` words), K independent loads in flight
template <int K>
__global__ void __launch_bounds__(1024,1) gather_kernel(const long long* arr, const int* idx, int iters, unsigned long long* out, long long* sink)
{
    long long acc = 0; unsigned long long t0 = gt();
    for (int it = 0; it < iters; ++it) {
        long long v[K];
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = __ldcg(arr + idx[((it * K + k) * 1024 + threadIdx.x) & ((1<<22)-1)]);
#pragma unroll
        for (int k = 0; k < K; ++k) acc += v[k];
        __syncthreads();
    }
    sink[threadIdx.x] = acc;
    if (threadIdx.x == 0) out[0] = gt() - t0;
}

int main()
{
    int4* slots; unsigned long long* out; int* sink;
    const int stride_max = 8;
    CK(cudaMalloc(&slots, (size_t)4 * 148 * stride_max * 16)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 4096));
    const int iters = 20000;
    for (int mode = 0; mode <= 2; ++mode) for (int stride : {1, 8}) for (int G : {2, 16, 74, 128, 148}) {
        CK(cudaMemset(slots, 0, (size_t)4 * 148 * stride_max * 16));
        int it = iters; int st = stride; void* args[] = {&slots, &st, &it, &mode, &out, &sink};
        CK(cudaLaunchCooperativeKernel((void*)a2a_kernel, dim3(G), dim3(1024), args, 0, 0));
        CK(cudaDeviceSynchronize());
        unsigned long long ns; CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
        printf("exchange mode %d stride %d G=%3d : %.3f us\n", mode, stride * 16, G, ns / 1000.0 / iters);
    }
    {   // hierarchical, clusters of 8 and 16
        for (int cs : {8, 16}) for (int nc : {2, 8, 16}) {
            if (cs * nc > 128) continue;
            CK(cudaMemset(slots, 0, (size_t)4 * 148 * stride_max * 16));
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs * nc); cfg.blockDim = dim3(1024);
            cudaLaunchAttribute at[2]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
            cfg.attrs = at; cfg.numAttrs = 2;
            int it = iters, st = 8; cudaError_t e;
            if (cs == 8) e = cudaLaunchKernelEx(&cfg, hier_kernel<8>, slots, st, it, out, sink);
            else { cudaFuncSetAttribute(hier_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); e = cudaLaunchKernelEx(&cfg, hier_kernel<16>, slots, st, it, out, sink); }
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("hier cs=%d nc=%d failed: %s\n", cs, nc, cudaGetErrorString(e)); cudaGetLastError(); continue; }
            unsigned long long ns; CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
            printf("hier cluster %2d x %2d clusters : %.3f us\n", cs, nc, ns / 1000.0 / iters);
        }
    }
    for (int mode = 0; mode < 2; ++mode) for (int G : {2, 148}) {
        CK(cudaMemset(slots, 0, 4096)); int it = iters;
        pingpong_kernel<<<G, 32>>>(slots, slots + 64, it, mode, out); CK(cudaDeviceSynchronize());
        unsigned long long ns; CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
        printf("pingpong mode %d (cta 0 <-> cta %d) round trip: %.3f us\n", mode, G - 1, ns / 1000.0 / iters);
    }
    {   // gather latency from an 8 MB (L2-resident) table
        const int n = 1 << 20; long long* arr; int* idx; long long* lsink;
        CK(cudaMalloc(&arr, (size_t)n * 8)); CK(cudaMalloc(&idx, (size_t)(1 << 22) * 4)); CK(cudaMalloc(&lsink, 1024 * 8));
        std::vector<int> h(1 << 22); unsigned long long s = 88172645463325252ULL;
        for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (int)(s % n); }
        CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(arr, 1, (size_t)n * 8));
        int it = 2000;
        gather_kernel<1><<<1, 1024>>>(arr, idx, it, out, lsink); CK(cudaDeviceSynchronize());
        gather_kernel<1><<<1, 1024>>>(arr, idx, it, out, lsink); CK(cudaDeviceSynchronize());
        unsigned long long ns; CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost)); printf("gather 1024 thr x 1 : %.3f us/round\n", ns / 1000.0 / it);
        gather_kernel<2><<<1, 1024>>>(arr, idx, it, out, lsink); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost)); printf("gather 1024 thr x 2 : %.3f us/round\n", ns / 1000.0 / it);
        gather_kernel<6><<<1, 1024>>>(arr, idx, it, out, lsink); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost)); printf("gather 1024 thr x 6 : %.3f us/round\n", ns / 1000.0 / it);
        gather_kernel<12><<<1, 1024>>>(arr, idx, it, out, lsink); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost)); printf("gather 1024 thr x 12 : %.3f us/round\n", ns / 1000.0 / it);
    }
    return 0;
}
