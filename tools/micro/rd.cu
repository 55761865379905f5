// Microbenchmark (round 2): how long does ONE CTA need to read 6144 16-byte records (96 KB) from L2 with ld.relaxed.gpu, as a
// function of who wrote them last?  (the pricing CTA of the team engine collects the node records the owner CTAs serve)
//   mode 0: nobody writes (clean lines)          mode 1: CTA 0 itself rewrites them every round
//   mode 2: ONE other CTA rewrites them          mode 3: all other CTAs, interleaved at 16-byte granularity (the engine's pattern)
//   mode 4: all other CTAs, each a contiguous range
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d: %s\n",#x,__LINE__,cudaGetErrorString(e)); return 1;}}while(0)
__device__ __forceinline__ int4 ld_rlx(const int4* p){int4 v; asm volatile("{\n .reg .b128 t;\n ld.relaxed.gpu.global.b128 t, [%4];\n mov.b128 {%0,%1,%2,%3}, t;\n}":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"l"(p):"memory"); return v;}
__device__ __forceinline__ void st_rlx(int4* p, int4 v){asm volatile("{\n .reg .b128 t;\n mov.b128 t, {%1,%2,%3,%4};\n st.relaxed.gpu.global.b128 [%0], t;\n}"::"l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}
__device__ __forceinline__ unsigned ld_flag(const unsigned* p){unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0,[%1];":"=r"(v):"l"(p):"memory"); return v;}
__device__ __forceinline__ void st_flag(unsigned* p, unsigned v){asm volatile("st.relaxed.gpu.global.u32 [%0],%1;"::"l"(p),"r"(v):"memory");}

constexpr int kT = 512, kRec = 6144;

__global__ void __launch_bounds__(kT, 1) rd_kernel(int4* buf, unsigned* go, unsigned* done, int iters, int mode, int inflight, unsigned long long* out, int* sink)
{
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    unsigned long long acc_clk = 0;
    int acc = 0;
    for (int it = 1; it <= iters; ++it) {
        // ---- writers
        if (mode == 1 && cta == 0) for (int r = tid; r < kRec; r += kT) st_rlx(buf + r, make_int4(r, it, 0, it));
        if (mode == 2 && cta == 1) for (int r = tid; r < kRec; r += kT) st_rlx(buf + r, make_int4(r, it, 0, it));
        if (mode == 3 && cta > 0) for (int r = (cta - 1) + tid * (G - 1); r < kRec; r += kT * (G - 1)) st_rlx(buf + r, make_int4(r, it, 0, it));
        if (mode == 4 && cta > 0) { const int per = (kRec + G - 2) / (G - 1), lo = (cta - 1) * per; for (int r = lo + tid; r < min(kRec, lo + per); r += kT) st_rlx(buf + r, make_int4(r, it, 0, it)); }
        __syncthreads();
        if (cta > 0) {
            if (tid == 0) { __threadfence(); st_flag(done + cta * 32, (unsigned)it); while (ld_flag(go) < (unsigned)it) ; }
            __syncthreads();
            continue;
        }
        // ---- the reader: wait until every writer is through (+ a little), then read everything
        if (tid < G && tid > 0) while (ld_flag(done + tid * 32) < (unsigned)it) ;
        __syncthreads();
        { const long long t0 = clock64(); while (clock64() - t0 < 4000) ; }
        __syncthreads();
        const long long c0 = clock64();
        if (inflight == 12) {
            int4 v[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) v[j] = ld_rlx(buf + tid + j * kT);
#pragma unroll
            for (int j = 0; j < 12; ++j) acc += v[j].x + (v[j].w == it);
        } else if (inflight == 6) {
#pragma unroll
            for (int jb = 0; jb < 12; jb += 6) {
                int4 v[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) v[j] = ld_rlx(buf + tid + (jb + j) * kT);
#pragma unroll
                for (int j = 0; j < 6; ++j) acc += v[j].x + (v[j].w == it);
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < 12; ++j) { const int4 v = ld_rlx(buf + tid + j * kT); acc += v.x + (v.w == it); }
        }
        __syncthreads();
        acc_clk += clock64() - c0;
        if (tid == 0) st_flag(go, (unsigned)it);
    }
    if (tid == 0) sink[cta] = acc;
    if (cta == 0 && tid == 0) out[0] = acc_clk;
}

int main()
{
    int4* buf; unsigned *go, *done; unsigned long long* out; int* sink;
    CK(cudaMalloc(&buf, kRec * 16)); CK(cudaMalloc(&go, 128)); CK(cudaMalloc(&done, 160 * 128)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 4096));
    const int iters = 2000;
    for (int mode = 0; mode <= 4; ++mode) for (int inflight : {12, 6, 1}) {
        CK(cudaMemset(buf, 0, kRec * 16)); CK(cudaMemset(go, 0, 128)); CK(cudaMemset(done, 0, 160 * 128));
        int it = iters, md = mode, inf = inflight; void* args[] = {&buf, &go, &done, &it, &md, &inf, &out, &sink};
        CK(cudaLaunchCooperativeKernel((void*)rd_kernel, dim3(148), dim3(kT), args, 0, 0));
        CK(cudaDeviceSynchronize());
        unsigned long long c; CK(cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost));
        printf("mode %d, %2d loads in flight per thread: %7.0f cycles = %.3f us to read 96 KB\n", mode, inflight, (double)c / iters, (double)c / iters / 1965.0);
    }
    return 0;
}
