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
__device__ __forceinline__ long long ld_pi_l1(const long long* pi, int u) { return __ldg(pi + u); }
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
            if (HINT == 3) {
                ps[u][0] = ld_pi_l1(A.pi, s[u].x);
                ps[u][1] = s[u].y == s[u].x ? ps[u][0] : ld_pi_l1(A.pi, s[u].y);
                ps[u][2] = s[u].z == s[u].y ? ps[u][1] : ld_pi_l1(A.pi, s[u].z);
                ps[u][3] = s[u].w == s[u].z ? ps[u][2] : ld_pi_l1(A.pi, s[u].w);
            } else {
            ps[u][0] = ld_pi<HINT>(A.pi, s[u].x, pl);
            ps[u][1] = s[u].y == s[u].x ? ps[u][0] : ld_pi<HINT>(A.pi, s[u].y, pl);
            ps[u][2] = s[u].z == s[u].y ? ps[u][1] : ld_pi<HINT>(A.pi, s[u].z, pl);
            ps[u][3] = s[u].w == s[u].z ? ps[u][2] : ld_pi<HINT>(A.pi, s[u].w, pl);
            }
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


// ---- variant B: the four arc streams go global -> shared by bulk async copies (cp.async.bulk / UBLKCP: the TMA engine, no L1TEX
// sectors, no registers), a ring of NS chunks of CH arcs per CTA; the last warp out of a chunk refills its stage.  A thread
// handles CH / THREADS arcs of a chunk (positions t, t + THREADS, ...: conflict-free shared-memory reads).  GATHER 0 = stream only.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* b, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
template <int THREADS, int MINB, int CH, int NS, int GATHER>
__global__ void __launch_bounds__(THREADS, MINB) k_bulk(const Arrays A, Best* out)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    int* const ring = reinterpret_cast<int*>(dyn);                                   // [NS][4][CH]
    unsigned long long* const full = reinterpret_cast<unsigned long long*>(dyn + (size_t)NS * 4 * CH * 4);
    int* const outc = reinterpret_cast<int*>(full + NS);
    constexpr int PER = CH / THREADS, NW = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, S = A.S;
    const int nchunk = (S + CH - 1) / CH;
    const int mine = blockIdx.x < nchunk ? (nchunk - 1 - blockIdx.x) / gridDim.x + 1 : 0;   // chunks blockIdx.x, + gridDim.x, ...
    auto issue = [&](int i) {
        const int c = blockIdx.x + i * gridDim.x, lo = c * CH, len = S - lo < CH ? S - lo : CH, s = i % NS;
        const unsigned bytes = (unsigned)len * 4u;                                  // S is a multiple of 4 here: 16-byte granules
        mbar_expect_tx(full + s, 4u * bytes);
        int* const st = ring + (size_t)s * 4 * CH;
        bulk_g2s(st, A.src + lo, bytes, full + s); bulk_g2s(st + CH, A.tgt + lo, bytes, full + s);
        bulk_g2s(st + 2 * CH, A.cost + lo, bytes, full + s); bulk_g2s(st + 3 * CH, A.state + lo, bytes, full + s);
    };
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); outc[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < NS && i < mine; ++i) issue(i);
    }
    __syncthreads();
    long long brc = 0; int barc = INT_MAX;
    for (int i = 0; i < mine; ++i) {
        const int s = i % NS;
        const unsigned parity = (unsigned)(i / NS) & 1u;
        while (!mbar_try_wait(full + s, parity)) { }
        const int c = blockIdx.x + i * gridDim.x, lo = c * CH;
        const int* const st = ring + (size_t)s * 4 * CH;
        int vs[PER], vt[PER], vc[PER], vst[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int q = tid + u * THREADS;
            const bool in = lo + q < S;
            vs[u] = in ? st[q] : 0; vt[u] = in ? st[CH + q] : 0; vc[u] = in ? st[2 * CH + q] : 0; vst[u] = in ? st[3 * CH + q] : 0;
        }
        __syncwarp();
        if (lane == 0 && atomicAdd(outc + s, 1) == NW - 1) {          // every warp has its values in registers: refill the stage
            outc[s] = 0;
            if (i + NS < mine) { asm volatile("fence.proxy.async;" ::: "memory"); issue(i + NS); }
        }
        if (GATHER) {
            long long d[PER];
#pragma unroll
            for (int u = 0; u < PER; ++u) d[u] = __ldcg(A.pi + vs[u]) - (vst[u] ? __ldcg(A.pi + vt[u]) : 0);
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const long long r = (long long)vst[u] * ((long long)vc[u] + d[u]);
                if (r < brc) { brc = r; barc = lo + tid + u * THREADS; }
            }
        } else {
#pragma unroll
            for (int u = 0; u < PER; ++u) { const long long r = -(long long)(((vs[u] ^ vt[u] ^ vc[u] ^ vst[u]) & 0x3fffff)) - 1; if (r < brc) { brc = r; barc = lo + tid + u * THREADS; } }
        }
    }
    block_min_out(brc, barc, out);
}

// ---- variant W: variant B with a rolling window - a thread keeps the arcs of W chunks in registers with their gathers in flight
// (slot k is retired right before it is refilled, W chunks later), so the gathers never drain at a chunk boundary while the ring
// stays small (the shared memory comes out of the L1 that tracks the gathers' misses).
template <int THREADS, int MINB, int CH, int NS, int W>
__global__ void __launch_bounds__(THREADS, MINB) k_bulkw(const Arrays A, Best* out)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    int* const ring = reinterpret_cast<int*>(dyn);                                   // [NS][4][CH]
    unsigned long long* const full = reinterpret_cast<unsigned long long*>(dyn + (size_t)NS * 4 * CH * 4);
    int* const outc = reinterpret_cast<int*>(full + NS);
    constexpr int PER = CH / THREADS, NW = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, S = A.S;
    const int nchunk = (S + CH - 1) / CH;
    const int mine = blockIdx.x < nchunk ? (nchunk - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    auto issue = [&](int i) {
        const int c = blockIdx.x + i * gridDim.x, lo = c * CH, len = S - lo < CH ? S - lo : CH, s = i % NS;
        const unsigned bytes = (unsigned)len * 4u;
        mbar_expect_tx(full + s, 4u * bytes);
        int* const st = ring + (size_t)s * 4 * CH;
        bulk_g2s(st, A.src + lo, bytes, full + s); bulk_g2s(st + CH, A.tgt + lo, bytes, full + s);
        bulk_g2s(st + 2 * CH, A.cost + lo, bytes, full + s); bulk_g2s(st + 3 * CH, A.state + lo, bytes, full + s);
    };
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); outc[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < NS && i < mine; ++i) issue(i);
    }
    __syncthreads();
    long long brc = 0; int barc = INT_MAX;
    int wc[W][PER], wst[W][PER], wlo[W]; long long wps[W][PER], wpt[W][PER];
#pragma unroll
    for (int k = 0; k < W; ++k) { wlo[k] = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) { wc[k][u] = 0; wst[k][u] = 0; wps[k][u] = 0; wpt[k][u] = 0; } }
    for (int i0 = 0; i0 < mine + W; i0 += W) {
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const int i = i0 + k;
#pragma unroll
            for (int u = 0; u < PER; ++u) {                                         // retire what slot k holds (state 0 = nothing)
                const long long r = (long long)wst[k][u] * ((long long)wc[k][u] + wps[k][u] - wpt[k][u]);
                if (r < brc) { brc = r; barc = wlo[k] + tid + u * THREADS; }
                wst[k][u] = 0;
            }
            if (i < mine) {
                const int s = i % NS;
                const unsigned parity = (unsigned)(i / NS) & 1u;
                while (!mbar_try_wait(full + s, parity)) { }
                const int lo = (blockIdx.x + i * gridDim.x) * CH;
                const int* const st = ring + (size_t)s * 4 * CH;
                int vs[PER], vt[PER];
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int q = tid + u * THREADS;
                    const bool in = lo + q < S;
                    vs[u] = in ? st[q] : 0; vt[u] = in ? st[CH + q] : 0; wc[k][u] = in ? st[2 * CH + q] : 0; wst[k][u] = in ? st[3 * CH + q] : 0;
                }
                wlo[k] = lo;
                __syncwarp();
                if (lane == 0 && atomicAdd(outc + s, 1) == NW - 1) {
                    outc[s] = 0;
                    if (i + NS < mine) { asm volatile("fence.proxy.async;" ::: "memory"); issue(i + NS); }
                }
#pragma unroll
                for (int u = 0; u < PER; ++u) { wps[k][u] = __ldcg(A.pi + vs[u]); wpt[k][u] = wst[k][u] ? __ldcg(A.pi + vt[u]) : 0; }
            }
        }
    }
    block_min_out(brc, barc, out);
}

__global__ void init_arcs(int* src, int* tgt, int* cost, int* state, long long* pi, int m, int n)
{
    const int S = m + n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < S; e += gridDim.x * blockDim.x) {
        unsigned long long x = (unsigned long long)e * 0x9E3779B97F4A7C15ULL; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
        if (e < m) { src[e] = (int)(((long long)(e >> 3) * 5) % n);   /* runs of 8 arcs per source, every source in its own 32-byte sector (as NETGEN-8: 16.5 source sectors per 128 arcs) */ tgt[e] = (int)(x % (unsigned)n); cost[e] = 1 + (int)((x >> 33) % 10000u); state[e] = (x >> 50) & 1 ? 1 : -1; }
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

static int g_check = 1;
template <typename K> void bench(const char* name, K kern, int grid, int threads, size_t smem = 0)
{
    if (smem) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ms;
    for (int r = 0; r < 11; ++r) {
        cudaMemsetAsync(g_flush, r, kFb);
        l2_read<<<148 * 4, 512>>>((const int4*)((char*)g_flush + kFb), kFb / 16, g_sink);
        cudaEventRecord(e0);
        kern<<<grid, threads, smem>>>(g_A, g_out);
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
           16.0 * g_A.S / 1e6 / t, 16.0 * g_A.S / 1e6 / t / 6545.3, !g_check ? "(stream only)" : (rc == g_ref_rc && arc == g_ref_arc) ? "ok" : "MISMATCH", err == cudaSuccess ? "" : cudaGetErrorString(err));
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
#define RUNB(T, MB, CH, NS, Gt) { g_check = Gt; bench("k_bulk<" #T "," #MB "," #CH "," #NS "," #Gt ">", k_bulk<T, MB, CH, NS, Gt>, sms * MB, T, (size_t)NS * 16 * CH + 16 * NS + 64); g_check = 1; }
    RUN((k_quad<1, 0, 1024, 1>), sms, 1024);                 // reference result + the shipped shape
    RUN((k_quad<1, 3, 1024, 1>), sms, 1024);                 // pi[source] through L1
    RUN((k_quad<1, 3, 512, 4>), sms * 4, 512);
    RUNB(256, 4, 1024, 2, 0) RUNB(256, 4, 1024, 3, 0) RUNB(512, 2, 2048, 2, 0) RUNB(1024, 1, 4096, 2, 0) RUNB(256, 4, 512, 4, 0) RUNB(128, 8, 512, 2, 0) RUNB(256, 2, 1024, 4, 0)
    RUNB(256, 4, 1024, 2, 1) RUNB(256, 4, 1024, 3, 1) RUNB(512, 2, 2048, 2, 1) RUNB(1024, 1, 4096, 2, 1) RUNB(256, 4, 512, 4, 1) RUNB(128, 8, 512, 2, 1) RUNB(256, 2, 1024, 4, 1)
    RUNB(256, 8, 512, 2, 1) RUNB(256, 6, 1024, 2, 1) RUNB(512, 3, 2048, 2, 1) RUNB(512, 4, 1024, 2, 1)
#define RUNW(T, MB, CH, NS, W) bench("k_bulkw<" #T "," #MB "," #CH "," #NS "," #W ">", k_bulkw<T, MB, CH, NS, W>, sms * MB, T, (size_t)NS * 16 * CH + 16 * NS + 64);
    RUNW(512, 2, 1024, 2, 3) RUNW(512, 2, 1024, 2, 4) RUNW(256, 4, 512, 2, 3) RUNW(1024, 1, 2048, 2, 3) RUNW(1024, 1, 2048, 2, 4) RUNW(512, 2, 1024, 3, 3)
    RUNW(256, 4, 1024, 2, 2) RUNW(512, 2, 2048, 2, 2) RUNW(512, 2, 512, 2, 6) RUNW(512, 2, 512, 4, 6) RUNW(1024, 1, 1024, 2, 6) RUNW(1024, 1, 1024, 4, 6) RUNW(1024, 1, 1024, 4, 8)
    RUNW(512, 3, 1024, 2, 3) RUNW(512, 4, 512, 2, 4) RUNW(256, 8, 256, 2, 4)
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
