// Micro-benchmark behind DESIGN.md 4.3 (round 2): launch shapes, loads in flight and cache hints for the Best Eligible
// pricing sweep on NETGEN-shaped data (SoA src/tgt/cost/state, sources grouped in runs of 8, random targets, the last n arcs
// basis arcs with state 0), cold L2 (256 MB overwritten + 256 MB read between launches).  Every variant returns the same
// arg-min; the harness checks that.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/micro/sweep tools/micro/sweep.cu
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

struct Arrays { const int *src, *tgt, *cost, *state; const long long* pi; int S; };
struct Best { long long rc; int arc; };

__device__ __forceinline__ unsigned long long pol_first() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ unsigned long long pol_last() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }

template <int HINT> __device__ __forceinline__ int4 ld_stream4(const int* base, int q, unsigned long long pol)
{
    int4 v;
    const int4* p = reinterpret_cast<const int4*>(base) + q;
    if (HINT == 0) v = __ldg(p);
    else if (HINT == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
template <int HINT> __device__ __forceinline__ int ld_stream1(const int* base, int e, unsigned long long pol)
{
    int v;
    if (HINT == 0) v = __ldg(base + e);
    else if (HINT == 1) asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(base + e));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(base + e), "l"(pol));
    return v;
}
template <int HINT> __device__ __forceinline__ long long ld_pi(const long long* pi, int u, unsigned long long pol)
{
    long long v;
    if (HINT < 2) v = __ldcg(pi + u);
    else asm volatile("ld.global.cg.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(pi + u), "l"(pol));
    return v;
}

__device__ __forceinline__ void block_min_out(long long rc, int arc, Best* out)
{
    __shared__ long long s_rc[32];
    __shared__ int s_arc[32];
    for (int o = 16; o > 0; o >>= 1) {
        const long long r2 = __shfl_xor_sync(0xffffffffu, rc, o); const int a2 = __shfl_xor_sync(0xffffffffu, arc, o);
        if (r2 < rc || (r2 == rc && a2 < arc)) { rc = r2; arc = a2; }
    }
    if ((threadIdx.x & 31) == 0) { s_rc[threadIdx.x >> 5] = rc; s_arc[threadIdx.x >> 5] = arc; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        rc = threadIdx.x < nw ? s_rc[threadIdx.x] : 0; arc = threadIdx.x < nw ? s_arc[threadIdx.x] : INT_MAX;
        for (int o = 16; o > 0; o >>= 1) {
            const long long r2 = __shfl_xor_sync(0xffffffffu, rc, o); const int a2 = __shfl_xor_sync(0xffffffffu, arc, o);
            if (r2 < rc || (r2 == rc && a2 < arc)) { rc = r2; arc = a2; }
        }
        if (threadIdx.x == 0) { out[blockIdx.x].rc = rc; out[blockIdx.x].arc = rc < 0 ? arc : -1; }
    }
}

// ---- variant Q: one quad (4 consecutive arcs, 128-bit loads) per thread and iteration, NQ quads in flight
template <int NQ, int HINT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_quad(const Arrays A, Best* out)
{
    const unsigned long long pf = HINT >= 2 ? pol_first() : 0, pl = HINT >= 2 ? pol_last() : 0;
    long long brc = 0; int barc = INT_MAX;
    const int nquad = A.S >> 2, stride = gridDim.x * THREADS;
    for (int q0 = blockIdx.x * THREADS + threadIdx.x; q0 < nquad; q0 += NQ * stride) {
        int4 s[NQ], t[NQ], c[NQ], st[NQ];
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
            const int q = q0 + u * stride;
            if (q < nquad) { s[u] = ld_stream4<HINT>(A.src, q, pf); t[u] = ld_stream4<HINT>(A.tgt, q, pf); c[u] = ld_stream4<HINT>(A.cost, q, pf); st[u] = ld_stream4<HINT>(A.state, q, pf); }
            else st[u] = make_int4(0, 0, 0, 0), s[u] = t[u] = c[u] = st[u];
        }
        long long pt[NQ][4], ps[NQ][4];
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
            pt[u][0] = st[u].x ? ld_pi<HINT>(A.pi, t[u].x, pl) : 0; pt[u][1] = st[u].y ? ld_pi<HINT>(A.pi, t[u].y, pl) : 0;
            pt[u][2] = st[u].z ? ld_pi<HINT>(A.pi, t[u].z, pl) : 0; pt[u][3] = st[u].w ? ld_pi<HINT>(A.pi, t[u].w, pl) : 0;
            ps[u][0] = ld_pi<HINT>(A.pi, s[u].x, pl);
            ps[u][1] = s[u].y == s[u].x ? ps[u][0] : ld_pi<HINT>(A.pi, s[u].y, pl);
            ps[u][2] = s[u].z == s[u].y ? ps[u][1] : ld_pi<HINT>(A.pi, s[u].z, pl);
            ps[u][3] = s[u].w == s[u].z ? ps[u][2] : ld_pi<HINT>(A.pi, s[u].w, pl);
        }
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
            const int e = (q0 + u * stride) << 2;
            const long long r0 = (long long)st[u].x * ((long long)c[u].x + ps[u][0] - pt[u][0]);
            const long long r1 = (long long)st[u].y * ((long long)c[u].y + ps[u][1] - pt[u][1]);
            const long long r2 = (long long)st[u].z * ((long long)c[u].z + ps[u][2] - pt[u][2]);
            const long long r3 = (long long)st[u].w * ((long long)c[u].w + ps[u][3] - pt[u][3]);
            if (r0 < brc) { brc = r0; barc = e; }
            if (r1 < brc) { brc = r1; barc = e + 1; }
            if (r2 < brc) { brc = r2; barc = e + 2; }
            if (r3 < brc) { brc = r3; barc = e + 3; }
        }
    }
    block_min_out(brc, barc, out);
}

// ---- variant P: the r01 kernel's shape - one quad priced while the next quad's arc data is in flight
template <int HINT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_pipe(const Arrays A, Best* out)
{
    const unsigned long long pf = HINT >= 2 ? pol_first() : 0, pl = HINT >= 2 ? pol_last() : 0;
    long long brc = 0; int barc = INT_MAX;
    const int nquad = A.S >> 2, stride = gridDim.x * THREADS;
    int q = blockIdx.x * THREADS + threadIdx.x;
    if (q < nquad) {
        int4 s = ld_stream4<HINT>(A.src, q, pf), t = ld_stream4<HINT>(A.tgt, q, pf), c = ld_stream4<HINT>(A.cost, q, pf), st = ld_stream4<HINT>(A.state, q, pf);
        for (;;) {
            const int qn = q + stride;
            int4 s2 = s, t2 = t, c2 = c, st2 = st;
            if (qn < nquad) { s2 = ld_stream4<HINT>(A.src, qn, pf); t2 = ld_stream4<HINT>(A.tgt, qn, pf); c2 = ld_stream4<HINT>(A.cost, qn, pf); st2 = ld_stream4<HINT>(A.state, qn, pf); }
            const long long pt0 = st.x ? ld_pi<HINT>(A.pi, t.x, pl) : 0, pt1 = st.y ? ld_pi<HINT>(A.pi, t.y, pl) : 0;
            const long long pt2 = st.z ? ld_pi<HINT>(A.pi, t.z, pl) : 0, pt3 = st.w ? ld_pi<HINT>(A.pi, t.w, pl) : 0;
            const long long ps0 = ld_pi<HINT>(A.pi, s.x, pl);
            const long long ps1 = s.y == s.x ? ps0 : ld_pi<HINT>(A.pi, s.y, pl);
            const long long ps2 = s.z == s.y ? ps1 : ld_pi<HINT>(A.pi, s.z, pl);
            const long long ps3 = s.w == s.z ? ps2 : ld_pi<HINT>(A.pi, s.w, pl);
            const long long r0 = (long long)st.x * ((long long)c.x + ps0 - pt0), r1 = (long long)st.y * ((long long)c.y + ps1 - pt1);
            const long long r2 = (long long)st.z * ((long long)c.z + ps2 - pt2), r3 = (long long)st.w * ((long long)c.w + ps3 - pt3);
            const int e = q << 2;
            if (r0 < brc) { brc = r0; barc = e; }
            if (r1 < brc) { brc = r1; barc = e + 1; }
            if (r2 < brc) { brc = r2; barc = e + 2; }
            if (r3 < brc) { brc = r3; barc = e + 3; }
            if (qn >= nquad) break;
            s = s2; t = t2; c = c2; st = st2; q = qn;
        }
    }
    block_min_out(brc, barc, out);
}

// ---- variant A: one arc per thread and iteration (32-bit coalesced loads), U arcs in flight: few registers, many threads
template <int U, int HINT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_arc(const Arrays A, Best* out)
{
    const unsigned long long pf = HINT >= 2 ? pol_first() : 0, pl = HINT >= 2 ? pol_last() : 0;
    long long brc = 0; int barc = INT_MAX;
    const int S = A.S, stride = gridDim.x * THREADS;
    for (int e0 = blockIdx.x * THREADS + threadIdx.x; e0 < S; e0 += U * stride) {
        int s[U], t[U], c[U], st[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * stride;
            if (e < S) { s[u] = ld_stream1<HINT>(A.src, e, pf); t[u] = ld_stream1<HINT>(A.tgt, e, pf); c[u] = ld_stream1<HINT>(A.cost, e, pf); st[u] = ld_stream1<HINT>(A.state, e, pf); }
            else { s[u] = t[u] = c[u] = st[u] = 0; }
        }
        long long d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) d[u] = ld_pi<HINT>(A.pi, s[u], pl) - (st[u] ? ld_pi<HINT>(A.pi, t[u], pl) : 0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = (long long)st[u] * ((long long)c[u] + d[u]);
            if (r < brc) { brc = r; barc = e0 + u * stride; }
        }
    }
    block_min_out(brc, barc, out);
}

__global__ void init_arcs(int* src, int* tgt, int* cost, int* state, long long* pi, int m, int n)
{
    const int S = m + n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < S; e += gridDim.x * blockDim.x) {
        unsigned long long x = (unsigned long long)e * 0x9E3779B97F4A7C15ULL; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
        if (e < m) { src[e] = e >> 3; tgt[e] = (int)(x % (unsigned)n); cost[e] = 1 + (int)((x >> 33) % 10000u); state[e] = (x >> 50) & 1 ? 1 : -1; }
        else { src[e] = e - m; tgt[e] = n; cost[e] = 0; state[e] = 0; }
    }
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u <= n; u += gridDim.x * blockDim.x) {
        unsigned long long x = (unsigned long long)(u + 77) * 0xD6E8FEB86659FD93ULL; x ^= x >> 31;
        pi[u] = (long long)(x % 20000001ULL) - 10000000LL;
    }
}
__global__ void __launch_bounds__(512) l2_read(const int4* buf, size_t n4, long long* sink)
{
    long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { const int4 v = __ldcg(buf + i); acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x7f5a5a5a5a5a5a5aLL) *sink = acc;
}

static Arrays g_A; static Best* g_out; static void* g_flush; static const size_t kFb = 256u << 20; static long long* g_sink;
static long long g_ref_rc = 1; static int g_ref_arc = -2;

template <typename K> void bench(const char* name, K kern, int grid, int threads)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ms;
    for (int r = 0; r < 11; ++r) {
        cudaMemsetAsync(g_flush, r, kFb);
        l2_read<<<148 * 4, 512>>>((const int4*)((char*)g_flush + kFb), kFb / 16, g_sink);
        cudaEventRecord(e0);
        kern<<<grid, threads>>>(g_A, g_out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1); ms.push_back(t);
    }
    cudaError_t err = cudaGetLastError();
    std::vector<Best> h(grid);
    cudaMemcpy(h.data(), g_out, sizeof(Best) * grid, cudaMemcpyDeviceToHost);
    long long rc = 0; int arc = -1;
    for (int g = 0; g < grid; ++g) if (h[g].arc >= 0 && (h[g].rc < rc || (h[g].rc == rc && h[g].arc < arc))) { rc = h[g].rc; arc = h[g].arc; }
    if (g_ref_arc == -2) { g_ref_rc = rc; g_ref_arc = arc; }
    std::sort(ms.begin(), ms.end());
    const float t = ms[ms.size() / 2];
    printf("%-34s grid %5d x %4d : median %6.1f us  min %6.1f us  %7.1f GB/s  frac-of-6545 %.3f  %s%s\n", name, grid, threads, t * 1e3, ms[0] * 1e3,
           16.0 * g_A.S / 1e6 / t, 16.0 * g_A.S / 1e6 / t / 6545.3, (rc == g_ref_rc && arc == g_ref_arc) ? "ok" : "MISMATCH", err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
}

int main()
{
    const int n = 1 << 20, m = 1 << 23, S = m + n;
    int *src, *tgt, *cost, *state; long long* pi;
    cudaMalloc(&src, 4 * (size_t)S); cudaMalloc(&tgt, 4 * (size_t)S); cudaMalloc(&cost, 4 * (size_t)S); cudaMalloc(&state, 4 * (size_t)S); cudaMalloc(&pi, 8 * (size_t)(n + 1));
    cudaMalloc(&g_out, sizeof(Best) * 148 * 32); cudaMalloc(&g_flush, 2 * kFb); cudaMemset(g_flush, 1, 2 * kFb); cudaMalloc(&g_sink, 8);
    init_arcs<<<148 * 8, 256>>>(src, tgt, cost, state, pi, m, n);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("init failed\n"); return 1; }
    g_A.src = src; g_A.tgt = tgt; g_A.cost = cost; g_A.state = state; g_A.pi = pi; g_A.S = S;
    const int sms = 148;
#define RUN(expr, grid, thr) bench(#expr, expr, grid, thr)
    RUN((k_pipe<0, 1024, 1>), sms, 1024);                    // the r01 kernel's shape
    RUN((k_pipe<1, 1024, 1>), sms, 1024);
    RUN((k_pipe<2, 1024, 1>), sms, 1024);
    RUN((k_pipe<0, 512, 2>), sms * 2, 512);
    RUN((k_pipe<0, 256, 4>), sms * 4, 256);
    RUN((k_quad<1, 0, 1024, 1>), sms, 1024);
    RUN((k_quad<2, 0, 1024, 1>), sms, 1024);
    RUN((k_quad<2, 1, 1024, 1>), sms, 1024);
    RUN((k_quad<2, 2, 1024, 1>), sms, 1024);
    RUN((k_quad<2, 0, 512, 2>), sms * 2, 512);
    RUN((k_quad<2, 0, 256, 4>), sms * 4, 256);
    RUN((k_quad<1, 0, 1024, 2>), sms * 2, 1024);             // 32 registers: 2048 threads per SM
    RUN((k_quad<1, 0, 512, 4>), sms * 4, 512);
    RUN((k_quad<1, 0, 256, 8>), sms * 8, 256);
    RUN((k_quad<1, 2, 256, 8>), sms * 8, 256);
    RUN((k_arc<4, 0, 1024, 2>), sms * 2, 1024);
    RUN((k_arc<4, 0, 256, 8>), sms * 8, 256);
    RUN((k_arc<8, 0, 256, 8>), sms * 8, 256);
    RUN((k_arc<8, 0, 512, 4>), sms * 4, 512);
    RUN((k_arc<8, 1, 256, 8>), sms * 8, 256);
    RUN((k_arc<8, 2, 256, 8>), sms * 8, 256);
    RUN((k_arc<8, 0, 256, 6>), sms * 6, 256);
    RUN((k_arc<16, 0, 256, 4>), sms * 4, 256);
    RUN((k_arc<4, 0, 256, 8>), sms * 16, 256);
    RUN((k_arc<8, 0, 128, 16>), sms * 16, 128);
    return 0;
}
