// Micro-benchmark behind DESIGN.md 4.3: what a read-only stream of the Best Eligible sweep's size (151 MB, cold L2) can reach on
// B200 for several launch shapes, with and without one random 8-byte gather per 16 bytes streamed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/stream tools/micro/stream.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

template <int UNROLL, int GATHER>
__global__ void rd(const int4* __restrict__ a, size_t n4, const long long* __restrict__ pi, int mask, long long* out)
{
    long long acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
        int4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcg(a + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (GATHER) acc += __ldcg(pi + (v[u].x & mask)) + v[u].y + v[u].z + v[u].w;
            else acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    for (; i < n4; i += stride) { int4 v = __ldcg(a + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 0x7fffffffffffLL) out[0] = acc;
}

int g_flush_mode = 0;      // 0: memset 256 MB (dirty lines stay in L2); 1: memset, then read another 256 MB (dirty lines written back before timing); 2: read only
template <int UNROLL, int GATHER>
float run(int grid, int block, const int4* a, size_t n4, const long long* pi, int mask, long long* out, void* flush, size_t fb)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ms;
    for (int r = 0; r < 9; ++r) {
        if (g_flush_mode != 2) cudaMemsetAsync(flush, r, fb);
        if (g_flush_mode != 0) rd<4, 0><<<148 * 4, 512>>>((const int4*)((char*)flush + fb), fb / 16, pi, mask, out);
        cudaEventRecord(e0);
        rd<UNROLL, GATHER><<<grid, block>>>(a, n4, pi, mask, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1); ms.push_back(t);
    }
    std::sort(ms.begin(), ms.end());
    return ms[ms.size() / 2];
}

int main()
{
    const size_t n4 = 9437184;                         // 16-byte records = S of NETGEN-8 2^20
    const int nodes = 1 << 20;
    int4* a; long long* pi; long long* out; void* flush; const size_t fb = 256u << 20;
    cudaMalloc(&a, n4 * 16); cudaMalloc(&pi, (size_t)nodes * 8); cudaMalloc(&out, 8); cudaMalloc(&flush, 2 * fb); cudaMemset(flush, 1, 2 * fb);
    std::vector<int4> h(n4);
    unsigned long long s = 88172645463325252ULL;
    for (size_t i = 0; i < n4; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = make_int4((int)(s >> 20), 1, 2, 3); }
    cudaMemcpy(a, h.data(), n4 * 16, cudaMemcpyHostToDevice); cudaMemset(pi, 0, (size_t)nodes * 8);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double mb = n4 * 16 / 1e6;
#define ROW(U, Gt, grid, block) { float t = run<U, Gt>(grid, block, a, n4, pi, nodes - 1, out, flush, fb); \
        printf("flush %d unroll %d gather %d grid %5d x %4d : %7.1f us  %7.1f GB/s\n", g_flush_mode, U, Gt, grid, block, t * 1e3, mb / t); }
    for (g_flush_mode = 0; g_flush_mode < 3; ++g_flush_mode) {
        ROW(2, 0, sms, 1024)
        ROW(4, 0, sms, 1024)
        ROW(8, 0, sms, 1024)
        ROW(4, 0, sms * 4, 512)
        ROW(4, 0, sms * 16, 256)
        ROW(2, 1, sms, 1024)
        ROW(4, 1, sms, 1024)
        ROW(8, 1, sms, 1024)
        ROW(8, 1, sms * 2, 1024)
        ROW(8, 1, sms * 8, 256)
    }
    return 0;
}
