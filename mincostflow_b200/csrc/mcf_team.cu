// mcf_team.cu - the "team" pivot engine of libmcfgpu (sm_100a): Block Search network simplex as one persistent
// cooperative kernel in which the basis tree never leaves the chip.
//
// Why: a pivot of NetworkSimplex.Solve() (NS.cs:282-341) is a chain of pointer walks over parent/pred/thread/
// succ_num/last_succ (NS.cs:925-1209).  On B200 a dependent L2 load costs ~150 ns and a grid-wide barrier ~1.3 us
// (profiles/r01_micro_latency.txt), so the walks are replaced by flat passes over an interval labelling (in[u] = DFS
// index, sz[u] = subtree size; see mcf_device.cuh) and the labelling is kept in SHARED MEMORY, sliced by node id over
// the CTAs of the team ("owners").  What has to cross between CTAs per pivot is then tiny, and it crosses as 16-byte
// words that carry their own sequence number (pivot index) in the same 128-bit store - no fence, no barrier:
//
//   hop 1  ENTER   pricing CTA -> all    entering arc, its endpoints' (pi, in), cost, state, capacity
//   hop 2  CYC     every owner -> all    its best leaving-arc candidate per side of the cycle (+ counts)
//  (hop 2b STEM    every owner -> all    only when the re-hung stem is longer than one node: the stem entries)
//   hop 3  DONE    every owner -> pricer "my pi / in / flow updates of this pivot are globally visible" (after a fence)
//
// CTA 0 ("pricer") runs BlockSearchPivot.FindEnteringArc (NS.cs:1339-1441) over the arc arrays and the global node
// mirror {pi, in}; owners run FindJoinNode + FindLeavingArc as an interval test over their slice, every CTA reduces
// the candidates redundantly to the same decision (strict '<' on the first walk, '<=' on the second, NS.cs:958-998),
// owners apply ChangeFlow / UpdateTreeStructure / UpdatePotentials (NS.cs:1012-1209) to the nodes they own.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "mcf_device.cuh"

namespace mcf {

namespace {

__device__ __forceinline__ int4 ld_vol4(const int4* p)
{
    int4 v;
    asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol4(int4* p, int4 v)
{
    asm volatile("st.volatile.global.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned ld_vol_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol_u32(unsigned* p, unsigned v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int lo32(long long v) { return (int)(unsigned)(unsigned long long)v; }
__device__ __forceinline__ int hi32(long long v) { return (int)(unsigned)((unsigned long long)v >> 32); }
__device__ __forceinline__ long long mk64(int lo, int hi) { return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo); }

struct Key { long long a; int b; int idx; };
__device__ __forceinline__ bool key_less(const Key& x, const Key& y) { return x.a < y.a || (x.a == y.a && x.b < y.b); }
__device__ __forceinline__ Key key_none() { Key k; k.a = LLONG_MAX; k.b = INT_MAX; k.idx = -1; return k; }
__device__ __forceinline__ Key warp_min(Key k)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Key t;
        t.a = __shfl_xor_sync(0xffffffffu, k.a, o);
        t.b = __shfl_xor_sync(0xffffffffu, k.b, o);
        t.idx = __shfl_xor_sync(0xffffffffu, k.idx, o);
        if (key_less(t, k)) k = t;
    }
    return k;
}

// pricing candidate ordered by (block, reduced cost, scan offset): the first block in scan order that holds a negative
// reduced cost wins, inside it the smallest reduced cost, among equals the first in scan order (NS.cs:1349-1395).
struct PKey { long long rc; int blk; int off; int idx; };
__device__ __forceinline__ bool pkey_less(const PKey& x, const PKey& y)
{
    if (x.blk != y.blk) return x.blk < y.blk;
    if (x.rc != y.rc) return x.rc < y.rc;
    return x.off < y.off;
}
__device__ __forceinline__ PKey warp_pmin(PKey k)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        PKey t;
        t.rc = __shfl_xor_sync(0xffffffffu, k.rc, o);
        t.blk = __shfl_xor_sync(0xffffffffu, k.blk, o);
        t.off = __shfl_xor_sync(0xffffffffu, k.off, o);
        t.idx = __shfl_xor_sync(0xffffffffu, k.idx, o);
        if (pkey_less(t, k)) k = t;
    }
    return k;
}

struct Cand {                       // leaving-arc candidate of one side of the cycle
    long long d;                    // residual in cycle direction
    int in, sz, pd;                 // labels and pred word of the node below the candidate arc
    int zero;                       // flow on the arc is 0 after the augmentation (-> STATE_LOWER), else STATE_UPPER
    int valid;
};

struct TeamShared {
    Key red[2][kWarps];
    PKey pred[kWarps];
    int4 ent[kMailWords];           // ENTER record of this pivot
    Cand c1, c2;                    // winners of the two sides (owner-local, then team-wide)
    int ncyc, nstem, abort, cnt1, cnt2, found;
    int pre[kTeamMax + 1];
    // pricing winner
    int w_arc, w_src, w_tgt, w_cost, w_state, w_in_s, w_in_t;
    long long w_pi_s, w_pi_t, w_upper;
};

__device__ Key block_min(Key k, Key* s_red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    k = warp_min(k);
    if (lane == 0) s_red[warp] = k;
    __syncthreads();
    if (warp == 0) {
        k = lane < kWarps ? s_red[lane] : key_none();
        k = warp_min(k);
        if (lane == 0) s_red[0] = k;
    }
    __syncthreads();
    k = s_red[0];
    __syncthreads();
    return k;
}

// poll one self-validating word until its sequence number matches; false = abandoned (abort flag or time-out)
__device__ __forceinline__ bool poll_word(const int4* p, int seq, int4& out, const TeamParams& P)
{
    int4 v = ld_vol4(p);
    if (v.w == seq) { out = v; return true; }
    const long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
        v = ld_vol4(p);
        if (v.w == seq) { out = v; return true; }
        if ((++spins & 255u) == 0) {
            if (*(volatile int*)&P.ctl->abort) return false;
            if ((unsigned long long)(clock64() - t0) > P.timeout_cycles) { *(volatile int*)&P.ctl->abort = 1; return false; }
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1) ns_team_kernel(const TeamParams P)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ TeamShared sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = P.team, cta = blockIdx.x, nown = G - 1;
    const int n = P.n, S = P.S;
    const bool pricer = cta == 0;
    const int own = cta - 1;
    const int lo = pricer ? 0 : own * P.slice;
    const int cntn = pricer ? 0 : max(0, min(n + 1, lo + P.slice) - lo);

    // dynamic shared memory: stem staging (sorted + unsorted), then the resident slice: pi (int64), in / sz / pd (int),
    // and the list of this slice's cycle nodes (local indices, u16)
    int* const st_in = reinterpret_cast<int*>(dyn_smem);
    int* const st_z = st_in + kStemCap;
    int* const st_pd = st_z + kStemCap;
    int* const tmp_in = st_pd + kStemCap;
    int* const tmp_z = tmp_in + kStemCap;
    int* const tmp_pd = tmp_z + kStemCap;
    long long* pi_s = reinterpret_cast<long long*>(tmp_pd + kStemCap);
    int* in_s = reinterpret_cast<int*>(pi_s + P.slice);
    int* sz_s = in_s + P.slice;
    int* pd_s = sz_s + P.slice;
    unsigned short* list = reinterpret_cast<unsigned short*>(pd_s + P.slice);

    for (int j = tid; j < cntn; j += kThreads) {
        const NodeRec r = P.node[lo + j];
        pi_s[j] = r.pi; in_s[j] = r.in; sz_s[j] = P.sz0[lo + j]; pd_s[j] = P.pd0[lo + j];
    }
    if (tid == 0) sh.abort = 0;
    __syncthreads();

    // pricer state (BlockSearchPivot fields, NS.cs:1294-1302)
    int next_arc = 0, B = P.block_size, cons_low = 0, cons_high = 0;
    long long arcs_checked = 0, rounds_total = 0;
    // replicated state
    long long iterations = 0, degenerate = 0, cycle_nodes = 0, moved_nodes = 0, max_cycle = 0, max_stem = 0, stem_x = 0;
    int status = ST_NOT_SOLVED;
    unsigned long long t_price = 0, t_cycle = 0, t_update = 0, t_wdone = 0, t_wcyc = 0, t_stem = 0, t_mark = 0, t_begin = 0;
    if (pricer && tid == 0) t_begin = t_mark = gtimer();
#define TICK(acc) do { if (pricer && tid == 0) { const unsigned long long t__ = gtimer(); acc += t__ - t_mark; t_mark = t__; } } while (0)

    for (;;) {
        const long long k = iterations + 1;
        const int seq = (int)(unsigned)k;
        const int par = (int)(k & 1);

        // ================================================================ pricer: wait DONE(k-1), price, post ENTER(k)
        if (pricer) {
            if (k > 1) {
                if (tid < nown) {
                    const unsigned want = (unsigned)(k - 1);
                    const unsigned* p = P.done + (size_t)(tid + 1) * 32;
                    if (ld_vol_u32(p) != want) {
                        const long long t0 = clock64();
                        unsigned spins = 0;
                        while (ld_vol_u32(p) != want) {
                            if ((++spins & 255u) == 0) {
                                if (*(volatile int*)&P.ctl->abort) { sh.abort = 1; break; }
                                if ((unsigned long long)(clock64() - t0) > P.timeout_cycles) { *(volatile int*)&P.ctl->abort = 1; sh.abort = 1; break; }
                            }
                        }
                    }
                }
                __syncthreads();
                if (sh.abort) { status = ST_ERR_BARRIER_TIMEOUT; break; }
            }
            TICK(t_wdone);

            // ---- BlockSearchPivot.FindEnteringArc, NS.cs:1339-1397
            bool found = false;
            int arcs_this = 0;
            long long off0 = 0;
            bool first_group = true;
            for (;;) {
                const int nb = first_group ? 1 : 8;
                long long hi = off0 + (long long)nb * B; if (hi > S) hi = S;
                PKey best; best.rc = 0; best.blk = INT_MAX; best.off = INT_MAX; best.idx = tid;
                int b_src = 0, b_tgt = 0, b_cost = 0, b_state = 0, b_in_s = 0, b_in_t = 0; long long b_pi_s = 0, b_pi_t = 0;
                for (long long off = off0 + tid; off < hi; off += kThreads) {
                    int idx = next_arc + (int)off; if (idx >= S) idx -= S;
                    const int s = __ldg(P.src + idx), t = __ldg(P.tgt + idx), c = __ldg(P.cost + idx);
                    const int st = __ldcg(P.state + idx);
                    const int4 rs = __ldcg(reinterpret_cast<const int4*>(P.node + s));
                    const int4 rt = __ldcg(reinterpret_cast<const int4*>(P.node + t));
                    const long long ps = mk64(rs.x, rs.y), pt = mk64(rt.x, rt.y);
                    const long long rc = (long long)st * ((long long)c + ps - pt);
                    if (rc < 0) {
                        const int blk = first_group ? 0 : (int)((off - off0) / B);
                        if (blk < best.blk || (blk == best.blk && rc < best.rc)) {
                            best.blk = blk; best.rc = rc; best.off = (int)off;
                            b_src = s; b_tgt = t; b_cost = c; b_state = st; b_in_s = rs.z; b_in_t = rt.z; b_pi_s = ps; b_pi_t = pt;
                        }
                    }
                }
                best = warp_pmin(best);
                if (lane == 0) sh.pred[warp] = best;
                __syncthreads();
                if (warp == 0) {
                    PKey q = sh.pred[lane];
                    q = warp_pmin(q);
                    if (lane == 0) sh.pred[0] = q;
                }
                __syncthreads();
                const PKey win = sh.pred[0];
                __syncthreads();
                rounds_total++;
                if (win.blk != INT_MAX) {
                    if (win.idx == tid) {
                        int idx = next_arc + win.off; if (idx >= S) idx -= S;
                        sh.w_arc = idx; sh.w_src = b_src; sh.w_tgt = b_tgt; sh.w_cost = b_cost; sh.w_state = b_state;
                        sh.w_in_s = b_in_s; sh.w_in_t = b_in_t; sh.w_pi_s = b_pi_s; sh.w_pi_t = b_pi_t;
                        sh.w_upper = __ldg(P.upper + idx);
                    }
                    long long end = off0 + (long long)(win.blk + 1) * B; if (end > S) end = S;
                    arcs_this = (int)end;
                    // `_nextArc = e` (NS.cs:1397): the last arc examined, or unchanged after a full sweep that ended inside a block
                    if (end < S || (long long)S % B == 0) { int e = next_arc + (int)end - 1; if (e >= S) e -= S; next_arc = e; }
                    found = true;
                    break;
                }
                off0 = hi;
                if (off0 >= S) { arcs_this = S; break; }
                first_group = false;
            }
            arcs_checked += arcs_this;
            if (found && P.adaptive) {              // NS.cs:1399-1438
                const double hit = arcs_this > 0 ? 1.0 / arcs_this : 0;
                if (hit < P.low_thr) {
                    cons_high = 0; cons_low++;
                    if (cons_low >= P.consecutive) { const int ns = (int)(B * P.shrink); B = P.dyn_min_block > ns ? P.dyn_min_block : ns; cons_low = 0; }
                } else if (hit > P.high_thr) {
                    cons_low = 0; cons_high++;
                    if (cons_high >= P.consecutive) { const int ns = (int)(B * P.grow); B = P.max_block_size < ns ? P.max_block_size : ns; cons_high = 0; }
                } else { cons_low = 0; cons_high = 0; }
            }
            __syncthreads();
            if (tid < 5) {
                int4 w;
                if (tid == 0) w = make_int4(found ? sh.w_arc : -1, sh.w_src, sh.w_tgt, seq);
                else if (tid == 1) w = make_int4(sh.w_cost, sh.w_state, found ? 1 : 0, seq);
                else if (tid == 2) w = make_int4(lo32(sh.w_pi_s), hi32(sh.w_pi_s), sh.w_in_s, seq);
                else if (tid == 3) w = make_int4(lo32(sh.w_pi_t), hi32(sh.w_pi_t), sh.w_in_t, seq);
                else w = make_int4(lo32(sh.w_upper), hi32(sh.w_upper), 0, seq);
                sh.ent[tid] = w;
                st_vol4(P.enter + (size_t)par * kMailWords + tid, w);
            }
            __syncthreads();
            TICK(t_price);
        } else {
            // ============================================================ owners: hop 1, wait ENTER(k)
            if (tid < 5) {
                int4 w;
                if (!poll_word(P.enter + (size_t)par * kMailWords + tid, seq, w, P)) sh.abort = 1;
                sh.ent[tid] = w;
            }
            __syncthreads();
            if (sh.abort) { status = ST_ERR_BARRIER_TIMEOUT; break; }
        }

        const int in_arc = sh.ent[0].x, a_src = sh.ent[0].y, a_tgt = sh.ent[0].z;
        const int a_cost = sh.ent[1].x, a_state = sh.ent[1].y, a_found = sh.ent[1].z;
        if (!a_found) { status = ST_OPTIMAL; break; }
        iterations = k;
        if (iterations > P.max_iterations) { status = ST_INFEASIBLE; break; }          // NS.cs:311-317
        const long long pi_src = mk64(sh.ent[2].x, sh.ent[2].y), pi_tgt = mk64(sh.ent[3].x, sh.ent[3].y);
        const int in_src = sh.ent[2].z, in_tgt = sh.ent[3].z;
        const long long upper_in = mk64(sh.ent[4].x, sh.ent[4].y);
        const bool lower_state = a_state == STATE_LOWER;
        const int first = lower_state ? a_src : a_tgt;                                  // NS.cs:948-957
        const int inF = lower_state ? in_src : in_tgt, inS = lower_state ? in_tgt : in_src;
        const long long piF = lower_state ? pi_src : pi_tgt, piS = lower_state ? pi_tgt : pi_src;

        // ================================================================ owners: cycle discovery over the slice, post CYC(k)
        int nc = 0;
        if (!pricer) {
            if (tid == 0) { sh.ncyc = 0; sh.cnt1 = 0; sh.cnt2 = 0; }
            __syncthreads();
            for (int j = tid; j < cntn; j += kThreads) {
                const int in_u = in_s[j], sz_u = sz_s[j];
                const bool hasF = (unsigned)(inF - in_u) < (unsigned)sz_u;
                const bool hasS = (unsigned)(inS - in_u) < (unsigned)sz_u;
                if (hasF != hasS) { const int p = atomicAdd(&sh.ncyc, 1); list[p] = (unsigned short)j; }
            }
            __syncthreads();
            nc = sh.ncyc;
            Cand m1, m2; m1.valid = 0; m2.valid = 0; m1.d = m2.d = 0; m1.in = m2.in = m1.sz = m2.sz = m1.pd = m2.pd = m1.zero = m2.zero = 0;
            if (nc > 0) {
                Key k1 = key_none(), k2 = key_none();
                Cand b1 = m1, b2 = m2;
                int c1 = 0, c2 = 0;
                for (int p = tid; p < nc; p += kThreads) {
                    const int j = list[p];
                    const int in_u = in_s[j], sz_u = sz_s[j], pd = pd_s[j];
                    const int e = pd >> 1;
                    const long long fl = __ldcg(P.flow + e), up = __ldg(P.upper + e);
                    const long long res = up == LLONG_MAX ? (LLONG_MAX / 2) : up - fl;      // NS.cs:970-971
                    const bool side1 = (unsigned)(inF - in_u) < (unsigned)sz_u;
                    const bool dir_up = pd & 1;
                    // first walk: residual capacity when pred_dir == DOWN, else the flow; second walk mirrored (NS.cs:968, :986)
                    const bool increase = side1 ? !dir_up : dir_up;
                    Key kk; kk.a = increase ? res : fl; kk.idx = p;
                    Cand c; c.d = kk.a; c.in = in_u; c.sz = sz_u; c.pd = pd; c.zero = (!increase) || up == 0; c.valid = 1;
                    if (side1) { c1++; kk.b = -in_u; if (key_less(kk, k1)) { k1 = kk; b1 = c; } }   // strict '<' walking up: deepest minimum
                    else       { c2++; kk.b = in_u;  if (key_less(kk, k2)) { k2 = kk; b2 = c; } }   // '<=' walking up: shallowest minimum
                }
                if (c1) atomicAdd(&sh.cnt1, c1);
                if (c2) atomicAdd(&sh.cnt2, c2);
                k1 = block_min(k1, sh.red[0]);
                k2 = block_min(k2, sh.red[1]);
                if (k1.idx >= 0 && (k1.idx % kThreads) == tid) sh.c1 = b1;
                if (k2.idx >= 0 && (k2.idx % kThreads) == tid) sh.c2 = b2;
                __syncthreads();
                if (k1.idx >= 0) m1 = sh.c1;
                if (k2.idx >= 0) m2 = sh.c2;
            }
            if (tid < 5) {
                int4 w;
                if (tid == 0) w = make_int4(nc > 0 ? sh.cnt1 : 0, nc > 0 ? sh.cnt2 : 0, (m1.zero ? 1 : 0) | (m2.zero ? 2 : 0), seq);
                else if (tid == 1) w = make_int4(lo32(m1.d), hi32(m1.d), m1.in, seq);
                else if (tid == 2) w = make_int4(m1.sz, m1.pd, m1.valid, seq);
                else if (tid == 3) w = make_int4(lo32(m2.d), hi32(m2.d), m2.in, seq);
                else w = make_int4(m2.sz, m2.pd, m2.valid, seq);
                st_vol4(P.cyc + ((size_t)par * G + cta) * kMailWords + tid, w);
            }
        }

        // ================================================================ all: hop 2, gather CYC(k) and decide
        {
            Key k1 = key_none(), k2 = key_none();
            Cand b1, b2; b1.valid = b2.valid = 0; b1.d = b2.d = 0; b1.in = b2.in = b1.sz = b2.sz = b1.pd = b2.pd = b1.zero = b2.zero = 0;
            int c1 = 0, c2 = 0;
            if (tid == 0) { sh.cnt1 = 0; sh.cnt2 = 0; }
            if (tid < nown) {
                const int4* rec = P.cyc + ((size_t)par * G + tid + 1) * kMailWords;
                int4 w0, w1, w2, w3, w4;
                bool ok = poll_word(rec + 0, seq, w0, P);
                ok = ok && poll_word(rec + 1, seq, w1, P) && poll_word(rec + 2, seq, w2, P) && poll_word(rec + 3, seq, w3, P) && poll_word(rec + 4, seq, w4, P);
                if (!ok) sh.abort = 1;
                else {
                    c1 = w0.x; c2 = w0.y;
                    if (w2.z) { b1.d = mk64(w1.x, w1.y); b1.in = w1.z; b1.sz = w2.x; b1.pd = w2.y; b1.zero = w0.z & 1; b1.valid = 1; k1.a = b1.d; k1.b = -b1.in; k1.idx = tid; }
                    if (w4.z) { b2.d = mk64(w3.x, w3.y); b2.in = w3.z; b2.sz = w4.x; b2.pd = w4.y; b2.zero = (w0.z >> 1) & 1; b2.valid = 1; k2.a = b2.d; k2.b = b2.in; k2.idx = tid; }
                }
            }
            __syncthreads();
            if (c1) atomicAdd(&sh.cnt1, c1);
            if (c2) atomicAdd(&sh.cnt2, c2);
            k1 = block_min(k1, sh.red[0]);
            k2 = block_min(k2, sh.red[1]);
            if (k1.idx == tid) sh.c1 = b1;
            if (k2.idx == tid) sh.c2 = b2;
            __syncthreads();
            if (sh.abort) { status = ST_ERR_BARRIER_TIMEOUT; break; }
            if (pricer && tid == 0) { const unsigned long long t = gtimer(); t_cycle += t - t_mark; t_wcyc += 0; t_mark = t; }
            const bool has1 = k1.idx >= 0, has2 = k2.idx >= 0;
            const Cand w1 = sh.c1, w2 = sh.c2;
            const int cnt = sh.cnt1 + sh.cnt2;

            long long delta = upper_in;                                                 // NS.cs:958
            int result = 0;
            if (has1 && w1.d < delta) { delta = w1.d; result = 1; }
            if (has2 && w2.d <= delta) { delta = w2.d; result = 2; }
            const bool change = result != 0;
            if (!change && delta == 0) { status = ST_UNBOUNDED; break; }                // NS.cs:321-325
            if (delta == 0) degenerate++;
            cycle_nodes += cnt; if (cnt > max_cycle) max_cycle = cnt;
            const Cand out = result == 1 ? w1 : w2;
            const long long val = (long long)a_state * delta;                          // NS.cs:1017
            const bool in_side1 = result == 1;
            const int u_in = in_side1 ? first : (lower_state ? a_tgt : a_src);          // NS.cs:999-1008
            const int a = out.in, s = out.sz;                                           // old interval of the re-hung subtree
            const int in_uin = in_side1 ? inF : inS;
            const int b = in_side1 ? inS : inF;                                         // in[v_in]
            int ns = 1;

            // ---- stem = cycle nodes on u_in's side from u_in up to u_out.  One node: it is the candidate itself.
            if (change) {
                if (a == in_uin) {
                    if (tid == 0) { st_in[0] = a; st_z[0] = s; st_pd[0] = out.pd; }
                } else {
                    // hop 2b: owners publish their stem entries, everybody collects and sorts them (deepest first)
                    stem_x++;
                    if (!pricer) {
                        if (tid == 0) sh.nstem = 0;
                        __syncthreads();
                        for (int p = tid; p < nc; p += kThreads) {
                            const int j = list[p];
                            const int in_u = in_s[j], sz_u = sz_s[j];
                            const bool side1 = (unsigned)(inF - in_u) < (unsigned)sz_u;
                            if (side1 == in_side1 && in_u >= a) {
                                const int q = atomicAdd(&sh.nstem, 1);
                                st_vol4(P.stemseg + (size_t)par * (n + 1) + lo + q, make_int4(in_u, sz_u, pd_s[j], seq));
                            }
                        }
                        __syncthreads();
                        if (tid == 0) st_vol4(P.stemhdr + ((size_t)par * G + cta) * kMailWords, make_int4(sh.nstem, 0, 0, seq));
                    }
                    int mycnt = 0;
                    if (tid < nown) {
                        int4 w;
                        if (!poll_word(P.stemhdr + ((size_t)par * G + tid + 1) * kMailWords, seq, w, P)) sh.abort = 1;
                        else mycnt = w.x;
                    }
                    if (tid < kTeamMax) sh.pre[tid + 1] = tid < nown ? mycnt : 0;
                    if (tid == 0) sh.pre[0] = 0;
                    __syncthreads();
                    if (sh.abort) { status = ST_ERR_BARRIER_TIMEOUT; break; }
                    if (warp == 0) {                                                    // inclusive scan of pre[1..nown]
                        int carry = 0;
                        for (int base = 1; base <= nown; base += 32) {
                            const int i = base + lane;
                            int v = i <= nown ? sh.pre[i] : 0;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
                            if (i <= nown) sh.pre[i] = v + carry;
                            carry += __shfl_sync(0xffffffffu, v, 31);
                        }
                    }
                    __syncthreads();
                    ns = sh.pre[nown];
                    if (ns > kStemCap || ns < 2) { status = ST_ERR_STEM_TOO_LONG; break; }
                    for (int q = tid; q < ns; q += kThreads) {
                        int l = 0, r = nown;                                            // owner o (0-based) with pre[o] <= q < pre[o+1]
                        while (r - l > 1) { const int mid = (l + r) >> 1; if (sh.pre[mid] <= q) l = mid; else r = mid; }
                        int4 w;
                        if (!poll_word(P.stemseg + (size_t)par * (n + 1) + (size_t)l * P.slice + (q - sh.pre[l]), seq, w, P)) sh.abort = 1;
                        tmp_in[q] = w.x; tmp_z[q] = w.y; tmp_pd[q] = w.z;
                    }
                    __syncthreads();
                    if (sh.abort) { status = ST_ERR_BARRIER_TIMEOUT; break; }
                    for (int q = tid; q < ns; q += kThreads) {                          // rank by counting (in[] values are distinct)
                        const int x = tmp_in[q];
                        int rank = 0;
                        for (int r = 0; r < ns; ++r) rank += tmp_in[r] > x;
                        st_in[rank] = x; st_z[rank] = tmp_z[q]; st_pd[rank] = tmp_pd[q];
                    }
                    if (pricer && tid == 0) { const unsigned long long t = gtimer(); t_stem += t - t_mark; t_mark = t; }
                }
                __syncthreads();
                if (ns > max_stem) max_stem = ns;
                moved_nodes += s;
            }

            // ================================================================ updates
            const bool dir_new_up = u_in == a_src;                                      // NS.cs:1143
            if (pricer) {
                // arc states (ChangeFlow, NS.cs:1031-1039): only the pricing scan reads them
                if (tid == 0) {
                    if (change) { P.state[in_arc] = STATE_TREE; P.state[out.pd >> 1] = out.zero ? STATE_LOWER : STATE_UPPER; }
                    else P.state[in_arc] = -a_state;
                }
            } else {
                const bool src_side1 = lower_state;                                     // is `first` the source of the entering arc?
                if (delta > 0 && tid == 0 && first >= lo && first < lo + cntn)
                    P.flow[in_arc] = (lower_state ? 0 : upper_in) + val;                // NS.cs:1018 (a non-tree arc sits at a bound)
                for (int p = tid; p < nc; p += kThreads) {
                    const int j = list[p];
                    const int in_u = in_s[j], sz_u = sz_s[j], pd = pd_s[j];
                    const bool side1 = (unsigned)(inF - in_u) < (unsigned)sz_u;
                    if (delta > 0) {                                                    // NS.cs:1020-1029
                        const bool on_src_side = side1 == src_side1;
                        const long long dv = (pd & 1) ? val : -val;                     // pred_dir * val
                        atomicAdd(reinterpret_cast<unsigned long long*>(P.flow + (pd >> 1)), (unsigned long long)(on_src_side ? -dv : dv));
                    }
                    if (change) {
                        if (side1 != in_side1) sz_s[j] = sz_u + s;                      // v_in .. join (NS.cs:1174-1177)
                        else if (in_u < a) sz_s[j] = sz_u - s;                          // v_out .. join (NS.cs:1179-1182)
                        else {                                                          // stem node k (NS.cs:1095-1146)
                            int kk = 0;
                            if (ns > 1) { int l = 0, r = ns - 1; while (l < r) { const int mid = (l + r) >> 1; if (st_in[mid] <= in_u) r = mid; else l = mid + 1; } kk = l; }
                            if (kk == 0) { pd_s[j] = in_arc * 2 + (dir_new_up ? 1 : 0); sz_s[j] = s; }
                            else { pd_s[j] = st_pd[kk - 1] ^ 1; sz_s[j] = s - st_z[kk - 1]; }
                        }
                    }
                }
                if (change) {
                    __syncthreads();
                    // re-label in[] in closed form, add sigma over the re-hung subtree (UpdatePotentials, NS.cs:1185-1209)
                    const int base = b < a ? b + 1 : b - s + 1;                         // new index of u_in: first child of v_in
                    const long long piU = in_side1 ? piF : piS, piV = in_side1 ? piS : piF;
                    const long long sigma = piV - piU - (dir_new_up ? (long long)a_cost : -(long long)a_cost);
                    for (int j = tid; j < cntn; j += kThreads) {
                        const int x = in_s[j];
                        if ((unsigned)(x - a) < (unsigned)s) {
                            int l = 0, r = ns - 1;                                      // smallest k with x inside subtree(stem k)
                            while (l < r) { const int mid = (l + r) >> 1; if ((unsigned)(x - st_in[mid]) < (unsigned)st_z[mid]) r = mid; else l = mid + 1; }
                            int off;
                            if (l == 0) off = x - st_in[0];
                            else {
                                int rr = x - st_in[l];
                                if (x > st_in[l - 1]) rr -= st_z[l - 1];
                                off = st_z[l - 1] + rr;
                            }
                            const int nx = base + off;
                            const long long np = pi_s[j] + sigma;
                            in_s[j] = nx; pi_s[j] = np;
                            NodeRec r2; r2.pi = np; r2.in = nx; r2.pad = 0;
                            P.node[lo + j] = r2;
                        } else if (b < a) {
                            if (x > b && x < a) { in_s[j] = x + s; P.node[lo + j].in = x + s; }
                        } else {
                            if (x >= a + s && x <= b) { in_s[j] = x - s; P.node[lo + j].in = x - s; }
                        }
                    }
                }
                // hop 3: everything this CTA wrote for pivot k is visible before DONE(k)
                __threadfence();
                __syncthreads();
                if (tid == 0) st_vol_u32(P.done + (size_t)cta * 32, (unsigned)k);
            }
            TICK(t_update);
        }
        if (P.stop_after > 0 && iterations >= P.stop_after) { status = ST_STOPPED_EARLY; break; }
    }
#undef TICK

    // =================================================================== epilogue
    const bool clean = status != ST_ERR_BARRIER_TIMEOUT;
    if (clean) {
        // one conventional grid barrier (counter + fences): all slices final, all global writes visible
        for (int j = tid; j < cntn; j += kThreads) if (lo + j < n) P.pi_out[lo + j] = pi_s[j];
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&P.ctl->bar, 1ULL);
            const long long t0 = clock64();
            while (*(volatile unsigned long long*)&P.ctl->bar < (unsigned long long)G) {
                if ((unsigned long long)(clock64() - t0) > P.timeout_cycles) { *(volatile int*)&P.ctl->abort = 1; break; }
            }
            __threadfence();
        }
        __syncthreads();
    }
    if (status == ST_OPTIMAL && clean) {
        // CheckFeasibility (NS.cs:1272-1283) over arcs [m, m+n); GetTotalCost (NS.cs:452-465) over [0, m)
        int bad = 0;
        for (int e = P.m + cta * kThreads + tid; e < S; e += G * kThreads) bad |= __ldcg(P.flow + e) != 0;
        if (bad) atomicOr(&P.ctl->infeasible, 1);
        long long acc = 0;
        for (int e = cta * kThreads + tid; e < P.m; e += G * kThreads) {
            long long f = __ldcg(P.flow + e);
            if (P.orig_lower) { const long long l = __ldg(P.orig_lower + e); if (l != 0) { f += l; P.flow[e] = f; } }   // NS.cs:375-388
            acc += f * (long long)__ldg(P.cost + e);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && acc != 0) atomicAdd(reinterpret_cast<unsigned long long*>(&P.ctl->total_cost), (unsigned long long)acc);
    }
    if (pricer && tid == 0) {
        Ctl* c = P.ctl;
        c->status = status; c->iterations = iterations; c->arcs_checked = arcs_checked; c->final_block_size = B;
        c->degenerate = degenerate; c->cycle_nodes = cycle_nodes; c->moved_nodes = moved_nodes;
        c->max_cycle = max_cycle; c->max_stem = max_stem; c->pricing_rounds = rounds_total;
        c->ns_price = t_price; c->ns_cycle = t_cycle; c->ns_update = t_update; c->ns_total = gtimer() - t_begin;
        c->ns_wait_done = t_wdone; c->ns_wait_cyc = t_wcyc; c->ns_stem = t_stem; c->stem_exchanges = stem_x;
    }
}

}  // namespace mcf

// ------------------------------------------------------------------------------------------------ launchers

extern "C" size_t mcfk_team_smem_bytes(int slice) { return (size_t)6 * mcf::kStemCap * 4 + (size_t)slice * mcf::kNodeSmemBytes + 16; }

// largest slice (nodes per owner CTA) that fits the opt-in shared memory of the device next to the kernel's static part
extern "C" int mcfk_team_max_slice(int device)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, mcf::ns_team_kernel) != cudaSuccess) return -2;
    const long long avail = (long long)prop.sharedMemPerBlockOptin - (long long)fa.sharedSizeBytes - 6LL * mcf::kStemCap * 4 - 1024;
    long long s = avail / mcf::kNodeSmemBytes;
    if (s > 65535) s = 65535;                       // the cycle list stores local indices as u16
    return (int)(s & ~7LL);
}

extern "C" int mcfk_launch_team(const mcf::TeamParams* p, cudaStream_t stream)
{
    const size_t smem = mcfk_team_smem_bytes(p->slice);
    cudaError_t e = cudaFuncSetAttribute(mcf::ns_team_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    void* args[] = {(void*)p};
    e = cudaLaunchCooperativeKernel((const void*)mcf::ns_team_kernel, dim3(p->team), dim3(mcf::kThreads), args, smem, stream);
    return (int)e;
}
