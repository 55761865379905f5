// mcf_team.cu - the "team" pivot engine of libmcfgpu (sm_100a): Block Search network simplex as one persistent
// cooperative kernel in which the basis tree never leaves the chip.
//
// Why: a pivot of NetworkSimplex.Solve() (NS.cs:282-341) is a chain of pointer walks over parent/pred/thread/
// succ_num/last_succ (NS.cs:925-1209).  On B200 a dependent L2 load costs ~150 ns, a DRAM miss ~1 us and a grid-wide
// barrier ~1.3 us (profiles/r01_micro_latency.txt), so the walks are replaced by flat passes over an interval labelling
// (in[u] = DFS index, sz[u] = subtree size; see mcf_device.cuh) and the whole basis - labels, pred arcs, and the flow and
// capacity of every tree arc - is kept in SHARED MEMORY, sliced by node id over the CTAs of the team ("owners").  What
// has to cross between CTAs per pivot is then tiny, and it crosses as 16-byte words that carry their own sequence number
// (the pivot index) in the same 128-bit store - no fence, no barrier (profiles/r01_micro_hop.txt):
//
//   hop 1  ENTER   pricing CTA -> all    entering arc, its endpoints' (pi, in), cost, state, capacity
//   hop 2  CYC     every owner -> all    its best leaving-arc candidate per side of the cycle (+ counts)
//  (hop 2b STEM    every owner -> all    only when the re-hung stem is longer than one node: the stem entries)
//   hop 3  DONE    every owner -> pricer "my pi / in updates of this pivot are globally visible" (after one fence)
//
// CTA 0 ("pricer") runs BlockSearchPivot.FindEnteringArc (NS.cs:1339-1441) over the arc arrays and the global node
// mirror {pi, in}; the next block's arc data is staged in its shared memory while the other hops are in flight.  Owners
// run FindJoinNode + FindLeavingArc as an interval test over their slice, every CTA reduces the candidates redundantly
// to the same decision (strict '<' on the first walk, '<=' on the second, NS.cs:958-998), owners apply ChangeFlow /
// UpdateTreeStructure / UpdatePotentials (NS.cs:1012-1209) to the nodes they own in ONE fused pass.  Arc flows of tree
// arcs live with the node below the arc; flow[] in global memory is written when an arc leaves the tree and at the end.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "mcf_device.cuh"

namespace mcf {

namespace {

constexpr int kTT = 512;                                // threads per CTA of the team kernel: latency-bound code, 128 registers each
constexpr int kTW = kTT / 32;
constexpr int kPf = 8;                                  // arcs per pricer thread staged ahead (first block up to 4096 arcs)

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
};

struct PWin {                       // payload of a pricing candidate
    int src, tgt, cost, state, in_s, in_t;
    long long pi_s, pi_t, upper;
};

struct Book {                       // statistics and timers: touched by thread 0 only, kept out of the register file
    long long arcs_checked, rounds_total, degenerate, cycle_nodes, moved_nodes, max_cycle, max_stem, stem_x;
    unsigned long long t_price, t_cycle, t_update, t_wdone, t_stem, t_mark, t_begin, c_begin, pr_mark;
    unsigned long long pr[16];
    int cons_low, cons_high;
};

struct TeamShared {
    Book bk;
    Key red[2][kTW];
    Cand wc[2][kTW];             // per-warp winners' payloads
    PKey pkey[kTW];
    PWin pwin[kTW];
    int4 ent[kMailWords];           // ENTER record of this pivot
    int nstem, abort, cnt;
    int pre[kTeamMax + 1];
};

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
            if (*(volatile int*)&P.ctl->abort) { out = v; return false; }
            if ((unsigned long long)(clock64() - t0) > P.timeout_cycles) { *(volatile int*)&P.ctl->abort = 1; out = v; return false; }
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(kTT, 1) ns_team_kernel(const TeamParams P)
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

    // dynamic shared memory.  Everybody: stem staging.  Owners: the resident slice.  Pricer: the staged next block.
    long long* const st_fl = reinterpret_cast<long long*>(dyn_smem);            // [kTeamStemCap] flow on stem k's old pred arc (after augmentation)
    int* const st_in = reinterpret_cast<int*>(st_fl + kTeamStemCap);                // [kTeamStemCap] sorted: stem 0 = u_in (deepest) .. u_out
    int* const st_z = st_in + kTeamStemCap;
    int* const st_pd = st_z + kTeamStemCap;
    int* const tmp_in = st_pd + kTeamStemCap;
    unsigned char* const body = reinterpret_cast<unsigned char*>(tmp_in + kTeamStemCap);
    // owners
    long long* const fl_s = reinterpret_cast<long long*>(body);                 // flow on the pred arc of node j
    long long* const up_s = fl_s + P.slice;                                     // capacity of the pred arc
    int* const in_s = reinterpret_cast<int*>(up_s + P.slice);
    int* const sz_s = in_s + P.slice;
    int* const pd_s = sz_s + P.slice;
    // pricer
    long long* const pf_up = reinterpret_cast<long long*>(body);                // [kPf * kTT]
    int* const pf_src = reinterpret_cast<int*>(pf_up + kPf * kTT);
    int* const pf_tgt = pf_src + kPf * kTT;
    int* const pf_cost = pf_tgt + kPf * kTT;
    int* const pf_st = pf_cost + kPf * kTT;

    for (int j = tid; j < cntn; j += kTT) {
        const int u = lo + j;
        const int pd = P.pd0[u];
        in_s[j] = P.node[u].in; sz_s[j] = P.sz0[u]; pd_s[j] = pd;
        fl_s[j] = pd >= 0 ? P.flow[pd >> 1] : 0; up_s[j] = pd >= 0 ? P.upper[pd >> 1] : 0;
    }
    if (tid == 0) sh.abort = 0;
    __syncthreads();

    // pricer state (BlockSearchPivot fields, NS.cs:1294-1302)
    int next_arc = 0, B = P.block_size;
    int pf_next = -1, pf_B = 0;                          // what is staged: block [pf_next, pf_next + pf_B) of the cyclic scan
    int patch_arc0 = -1, patch_st0 = 0, patch_arc1 = -1, patch_st1 = 0;   // state changes decided after the staging loads
    // replicated state
    long long iterations = 0;
    int status = ST_NOT_SOLVED;
    // phase accumulators in SM clock ticks (a %globaltimer read costs microseconds, so it is read twice per solve)
    if (tid == 0) {
        Book z = {};
        sh.bk = z;
        sh.bk.t_begin = gtimer(); sh.bk.c_begin = sh.bk.t_mark = sh.bk.pr_mark = (unsigned long long)clock64();
    }
#define PROBE(i) do { if (tid == 0 && ((i) < 8 ? pricer : cta == 1)) { const unsigned long long t__ = (unsigned long long)clock64(); sh.bk.pr[i] += t__ - sh.bk.pr_mark; sh.bk.pr_mark = t__; } } while (0)
#define TICK(acc) do { if (pricer && tid == 0) { const unsigned long long t__ = (unsigned long long)clock64(); sh.bk.acc += t__ - sh.bk.t_mark; sh.bk.t_mark = t__; } } while (0)

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
            PROBE(0);

            // ---- BlockSearchPivot.FindEnteringArc, NS.cs:1339-1397
            bool found = false;
            int arcs_this = 0;
            long long off0 = 0;
            bool first_group = true;
            const bool staged = pf_next == next_arc && pf_B == B;
            for (;;) {
                const int nb = first_group ? 1 : 8;
                long long hi = off0 + (long long)nb * B; if (hi > S) hi = S;
                PKey best; best.rc = 0; best.blk = INT_MAX; best.off = INT_MAX; best.idx = warp;
                PWin bw; bw.src = bw.tgt = bw.cost = bw.state = bw.in_s = bw.in_t = 0; bw.pi_s = bw.pi_t = bw.upper = 0;
                if (first_group && staged) {
                    // the block was staged in shared memory while the previous pivot's hops were in flight; all gathers of a
                    // thread are issued before the first use
#pragma unroll
                    for (int jb = 0; jb < kPf; jb += 4) {
                        int4 rs[4], rt[4];
                        int st[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int off = tid + (jb + j) * kTT;
                            if (off < hi) {
                                int idx = next_arc + off; if (idx >= S) idx -= S;
                                rs[j] = __ldcg(reinterpret_cast<const int4*>(P.node + pf_src[off]));
                                rt[j] = __ldcg(reinterpret_cast<const int4*>(P.node + pf_tgt[off]));
                                st[j] = idx == patch_arc0 ? patch_st0 : (idx == patch_arc1 ? patch_st1 : pf_st[off]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int off = tid + (jb + j) * kTT;
                            if (off < hi) {
                                const long long ps = mk64(rs[j].x, rs[j].y), pt = mk64(rt[j].x, rt[j].y);
                                const int c = pf_cost[off];
                                const long long rc = (long long)st[j] * ((long long)c + ps - pt);
                                if (rc < best.rc) {
                                    best.blk = 0; best.rc = rc; best.off = off;
                                    bw.src = pf_src[off]; bw.tgt = pf_tgt[off]; bw.cost = c; bw.state = st[j]; bw.in_s = rs[j].z; bw.in_t = rt[j].z;
                                    bw.pi_s = ps; bw.pi_t = pt; bw.upper = pf_up[off];
                                }
                            }
                        }
                    }
                } else {
                    for (long long off = off0 + tid; off < hi; off += kTT) {
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
                                bw.src = s; bw.tgt = t; bw.cost = c; bw.state = st; bw.in_s = rs.z; bw.in_t = rt.z; bw.pi_s = ps; bw.pi_t = pt;
                                bw.upper = LLONG_MIN;                   // fetched by the winner only
                            }
                        }
                    }
                }
                // warp winner -> shared memory (key + payload), then every thread reduces the 32 warp keys itself
                {
                    PKey wk = best; wk.idx = lane;
                    wk = warp_pmin(wk);
                    if (wk.idx == lane) {
                        if (bw.upper == LLONG_MIN && wk.blk != INT_MAX) { int idx = next_arc + wk.off; if (idx >= S) idx -= S; bw.upper = __ldg(P.upper + idx); }
                        PKey o = wk; o.idx = warp; sh.pkey[warp] = o; sh.pwin[warp] = bw;
                    }
                }
                PROBE(1);
                __syncthreads();
                PKey win = sh.pkey[lane & (kTW - 1)];
                win = warp_pmin(win);
                if (tid == 0) sh.bk.rounds_total++;
                if (win.blk != INT_MAX) {
                    long long end = off0 + (long long)(win.blk + 1) * B; if (end > S) end = S;
                    arcs_this = (int)end;
                    const PWin w = sh.pwin[win.idx];
                    if (tid < 5) {
                        int widx = next_arc + win.off; if (widx >= S) widx -= S;
                        int4 o;
                        if (tid == 0) o = make_int4(widx, w.src, w.tgt, seq);
                        else if (tid == 1) o = make_int4(w.cost, w.state, 1, seq);
                        else if (tid == 2) o = make_int4(lo32(w.pi_s), hi32(w.pi_s), w.in_s, seq);
                        else if (tid == 3) o = make_int4(lo32(w.pi_t), hi32(w.pi_t), w.in_t, seq);
                        else o = make_int4(lo32(w.upper), hi32(w.upper), 0, seq);
                        st_vol4(P.enter + (size_t)par * kMailWords + tid, o);
                        sh.ent[tid] = o;
                    }
                    // `_nextArc = e` (NS.cs:1397): the last arc examined, or unchanged after a full sweep that ended inside a block
                    if (end < S || (long long)S % B == 0) { int e = next_arc + (int)end - 1; if (e >= S) e -= S; next_arc = e; }
                    found = true;
                    break;
                }
                __syncthreads();                                        // sh.pkey is rewritten by the next group
                off0 = hi;
                if (off0 >= S) { arcs_this = S; break; }
                first_group = false;
            }
            if (tid == 0) sh.bk.arcs_checked += arcs_this;
            if (!found) {
                if (tid < 5) { const int4 o = make_int4(-1, 0, 0, seq); st_vol4(P.enter + (size_t)par * kMailWords + tid, o); sh.ent[tid] = o; }
            } else if (P.adaptive) {                                    // NS.cs:1399-1438
                const double hit = arcs_this > 0 ? 1.0 / arcs_this : 0;
                int cl = sh.bk.cons_low, ch = sh.bk.cons_high;                  // every thread computes the same B; thread 0 stores the counters
                if (hit < P.low_thr) {
                    ch = 0; cl++;
                    if (cl >= P.consecutive) { const int ns = (int)(B * P.shrink); B = P.dyn_min_block > ns ? P.dyn_min_block : ns; cl = 0; }
                } else if (hit > P.high_thr) {
                    cl = 0; ch++;
                    if (ch >= P.consecutive) { const int ns = (int)(B * P.grow); B = P.max_block_size < ns ? P.max_block_size : ns; ch = 0; }
                } else { cl = 0; ch = 0; }
                __syncthreads();
                if (tid == 0) { sh.bk.cons_low = cl; sh.bk.cons_high = ch; }
            }
            __syncthreads();
            TICK(t_price);
            PROBE(2);
            // ---- stage the next pivot's first block: arc data streams from DRAM while hops 1..3 are in flight
            if (found && B <= kPf * kTT) {
                const int lim = B < S ? B : S;
#pragma unroll
                for (int j = 0; j < kPf; ++j) {
                    const int off = tid + j * kTT;
                    if (off < lim) {
                        int idx = next_arc + off; if (idx >= S) idx -= S;
                        pf_src[off] = __ldg(P.src + idx); pf_tgt[off] = __ldg(P.tgt + idx); pf_cost[off] = __ldg(P.cost + idx);
                        pf_st[off] = __ldcg(P.state + idx); pf_up[off] = __ldg(P.upper + idx);
                    }
                }
                pf_next = next_arc; pf_B = B;
            } else pf_next = -1;
            patch_arc0 = patch_arc1 = -1;
            PROBE(3);
        } else {
            PROBE(8);
            // ============================================================ owners: hop 1, wait ENTER(k)
            if (tid < 5) {
                int4 w;
                if (!poll_word(P.enter + (size_t)par * kMailWords + tid, seq, w, P)) sh.abort = 1;
                sh.ent[tid] = w;
            }
            __syncthreads();
            if (sh.abort) { status = ST_ERR_BARRIER_TIMEOUT; break; }
            PROBE(9);
        }

        const int in_arc = sh.ent[0].x, a_src = sh.ent[0].y, a_tgt = sh.ent[0].z;
        const int a_cost = sh.ent[1].x, a_state = sh.ent[1].y;
        if (in_arc < 0) { status = ST_OPTIMAL; break; }
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
        if (!pricer) {
            Key k1 = key_none(), k2 = key_none();
            Cand b1, b2; b1.d = b2.d = 0; b1.in = b2.in = b1.sz = b2.sz = b1.pd = b2.pd = b1.zero = b2.zero = 0;
            int c = 0;
            if (tid == 0) sh.cnt = 0;
            for (int j = tid; j < cntn; j += kTT) {
                const int in_u = in_s[j], sz_u = sz_s[j];
                const bool hasF = (unsigned)(inF - in_u) < (unsigned)sz_u;
                const bool hasS = (unsigned)(inS - in_u) < (unsigned)sz_u;
                if (hasF != hasS) {
                    c++;
                    const int pd = pd_s[j];
                    const long long fl = fl_s[j], up = up_s[j];
                    const long long res = up == LLONG_MAX ? (LLONG_MAX / 2) : up - fl;      // NS.cs:970-971
                    const bool dir_up = pd & 1;
                    // first walk: residual capacity when pred_dir == DOWN, else the flow; second walk mirrored (NS.cs:968, :986)
                    const bool increase = hasF ? !dir_up : dir_up;
                    Key kk; kk.a = increase ? res : fl; kk.idx = lane;
                    Cand cd; cd.d = kk.a; cd.in = in_u; cd.sz = sz_u; cd.pd = pd; cd.zero = (!increase) || up == 0;
                    if (hasF) { kk.b = -in_u; if (key_less(kk, k1)) { k1 = kk; b1 = cd; } }   // strict '<' walking up: deepest minimum
                    else      { kk.b = in_u;  if (key_less(kk, k2)) { k2 = kk; b2 = cd; } }   // '<=' walking up: shallowest minimum
                }
            }
            const int any = __syncthreads_or(c > 0);
            PROBE(10);
            if (!any) {
                if (tid < 5) st_vol4(P.cyc + ((size_t)par * G + cta) * kMailWords + tid, make_int4(0, 0, 0, seq));
            } else {
                if (__any_sync(0xffffffffu, c > 0)) {
                    int cw = c;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cw += __shfl_xor_sync(0xffffffffu, cw, o);
                    if (lane == 0) atomicAdd(&sh.cnt, cw);
                    k1 = warp_min(k1); k2 = warp_min(k2);
                    if (k1.idx == lane) sh.wc[0][warp] = b1;
                    if (k2.idx == lane) sh.wc[1][warp] = b2;
                } else { k1 = key_none(); k2 = key_none(); }
                if (lane == 0) { k1.idx = k1.idx >= 0 ? warp : -1; k2.idx = k2.idx >= 0 ? warp : -1; sh.red[0][warp] = k1; sh.red[1][warp] = k2; }
                __syncthreads();
                if (warp == 0) {
                    Key f1 = warp_min(sh.red[0][lane & (kTW - 1)]), f2 = warp_min(sh.red[1][lane & (kTW - 1)]);
                    if (lane < 5) {
                        Cand m1, m2; m1.d = m2.d = 0; m1.in = m2.in = m1.sz = m2.sz = m1.pd = m2.pd = m1.zero = m2.zero = 0;
                        if (f1.idx >= 0) m1 = sh.wc[0][f1.idx];
                        if (f2.idx >= 0) m2 = sh.wc[1][f2.idx];
                        int4 w;
                        if (lane == 0) w = make_int4(sh.cnt, 0, (m1.zero ? 1 : 0) | (m2.zero ? 2 : 0), seq);
                        else if (lane == 1) w = make_int4(lo32(m1.d), hi32(m1.d), m1.in, seq);
                        else if (lane == 2) w = make_int4(m1.sz, m1.pd, f1.idx >= 0 ? 1 : 0, seq);
                        else if (lane == 3) w = make_int4(lo32(m2.d), hi32(m2.d), m2.in, seq);
                        else w = make_int4(m2.sz, m2.pd, f2.idx >= 0 ? 1 : 0, seq);
                        st_vol4(P.cyc + ((size_t)par * G + cta) * kMailWords + lane, w);
                    }
                }
            }
        }

        PROBE(11);
        // ================================================================ all: hop 2, gather CYC(k) and decide
        {
            Key k1 = key_none(), k2 = key_none();
            Cand b1, b2; b1.d = b2.d = 0; b1.in = b2.in = b1.sz = b2.sz = b1.pd = b2.pd = b1.zero = b2.zero = 0;
            if (tid == 0) sh.cnt = 0;
            __syncthreads();
            const int nw = (nown + 31) >> 5;                                            // warps that poll
            if (warp < nw) {
                int c = 0;
                if (tid < nown) {
                    const int4* rec = P.cyc + ((size_t)par * G + tid + 1) * kMailWords;
                    int4 w0, w1, w2, w3, w4;
                    bool ok = poll_word(rec + 0, seq, w0, P);
                    ok = ok && poll_word(rec + 1, seq, w1, P) && poll_word(rec + 2, seq, w2, P) && poll_word(rec + 3, seq, w3, P) && poll_word(rec + 4, seq, w4, P);
                    if (!ok) sh.abort = 1;
                    else {
                        c = w0.x;
                        if (w2.z) { b1.d = mk64(w1.x, w1.y); b1.in = w1.z; b1.sz = w2.x; b1.pd = w2.y; b1.zero = w0.z & 1; k1.a = b1.d; k1.b = -b1.in; k1.idx = lane; }
                        if (w4.z) { b2.d = mk64(w3.x, w3.y); b2.in = w3.z; b2.sz = w4.x; b2.pd = w4.y; b2.zero = (w0.z >> 1) & 1; k2.a = b2.d; k2.b = b2.in; k2.idx = lane; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (lane == 0 && c) atomicAdd(&sh.cnt, c);
                k1 = warp_min(k1); k2 = warp_min(k2);
                if (k1.idx == lane) sh.wc[0][warp] = b1;
                if (k2.idx == lane) sh.wc[1][warp] = b2;
                if (lane == 0) { k1.idx = k1.idx >= 0 ? warp : -1; k2.idx = k2.idx >= 0 ? warp : -1; sh.red[0][warp] = k1; sh.red[1][warp] = k2; }
            }
            __syncthreads();
            if (sh.abort) { status = ST_ERR_BARRIER_TIMEOUT; break; }
            Key f1 = key_none(), f2 = key_none();
            for (int w = 0; w < nw; ++w) {
                const Key t1 = sh.red[0][w], t2 = sh.red[1][w];
                if (t1.idx >= 0 && key_less(t1, f1)) f1 = t1;
                if (t2.idx >= 0 && key_less(t2, f2)) f2 = t2;
            }
            const bool has1 = f1.idx >= 0, has2 = f2.idx >= 0;
            Cand w1, w2; w1.d = w2.d = 0; w1.in = w2.in = w1.sz = w2.sz = w1.pd = w2.pd = w1.zero = w2.zero = 0;
            if (has1) w1 = sh.wc[0][f1.idx];
            if (has2) w2 = sh.wc[1][f2.idx];
            const int cnt = sh.cnt;
            TICK(t_cycle);
            PROBE(4); PROBE(12);

            long long delta = upper_in;                                                 // NS.cs:958
            int result = 0;
            if (has1 && w1.d < delta) { delta = w1.d; result = 1; }
            if (has2 && w2.d <= delta) { delta = w2.d; result = 2; }
            const bool change = result != 0;
            if (!change && delta == 0) { status = ST_UNBOUNDED; break; }                // NS.cs:321-325
            if (tid == 0) { if (delta == 0) sh.bk.degenerate++; sh.bk.cycle_nodes += cnt; if (cnt > sh.bk.max_cycle) sh.bk.max_cycle = cnt; }
            const Cand out = result == 1 ? w1 : w2;
            const long long val = (long long)a_state * delta;                          // NS.cs:1017
            const bool in_side1 = result == 1;
            const int u_in = in_side1 ? first : (lower_state ? a_tgt : a_src);          // NS.cs:999-1008
            const int a = out.in, s = out.sz;                                           // old interval of the re-hung subtree
            const int in_uin = in_side1 ? inF : inS;
            const int b = in_side1 ? inS : inF;                                         // in[v_in]
            const bool src_side1 = lower_state;                                         // is `first` the source of the entering arc?
            int ns = 1;

            // ---- stem = cycle nodes on u_in's side from u_in up to u_out.  One node (74 % of pivots): nothing to exchange.
            if (change && a != in_uin) {
                // hop 2b: owners publish their stem entries (with the flow AFTER the augmentation), everybody sorts them
                if (tid == 0) sh.bk.stem_x++;
                __syncthreads();                                                        // sh.red / sh.wc readers are done
                if (!pricer) {
                    if (tid == 0) sh.nstem = 0;
                    __syncthreads();
                    for (int j = tid; j < cntn; j += kTT) {
                        const int in_u = in_s[j], sz_u = sz_s[j];
                        const bool hasF = (unsigned)(inF - in_u) < (unsigned)sz_u;
                        const bool hasS = (unsigned)(inS - in_u) < (unsigned)sz_u;
                        if (hasF != hasS && hasF == in_side1 && in_u >= a) {
                            const int pd = pd_s[j];
                            long long fl = fl_s[j];
                            if (delta > 0) { const long long dv = (pd & 1) ? val : -val; fl = (hasF == src_side1) ? fl - dv : fl + dv; }
                            const int q = atomicAdd(&sh.nstem, 1);
                            int4* e = P.stemseg + ((size_t)par * (n + 1) + lo + q) * 2;
                            st_vol4(e, make_int4(in_u, sz_u, pd, seq));
                            st_vol4(e + 1, make_int4(lo32(fl), hi32(fl), 0, seq));
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
                if (warp == 0) {                                                        // inclusive scan of pre[1..nown]
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
                if (ns > kTeamStemCap || ns < 2) { status = ST_ERR_STEM_TOO_LONG; break; }
                constexpr int kEnt = kTeamStemCap / kTT;                                // entries per thread
                int e_in[kEnt], e_z[kEnt], e_pd[kEnt]; long long e_fl[kEnt];
#pragma unroll
                for (int i = 0; i < kEnt; ++i) {
                    const int q = tid + i * kTT;
                    e_in[i] = e_z[i] = e_pd[i] = 0; e_fl[i] = 0;
                    if (q < ns) {
                        int l = 0, r = nown;                                            // owner l (0-based) with pre[l] <= q < pre[l+1]
                        while (r - l > 1) { const int mid = (l + r) >> 1; if (sh.pre[mid] <= q) l = mid; else r = mid; }
                        const int4* e = P.stemseg + ((size_t)par * (n + 1) + (size_t)l * P.slice + (q - sh.pre[l])) * 2;
                        int4 wa, wb;
                        if (!poll_word(e, seq, wa, P) || !poll_word(e + 1, seq, wb, P)) sh.abort = 1;
                        e_in[i] = wa.x; e_z[i] = wa.y; e_pd[i] = wa.z; e_fl[i] = mk64(wb.x, wb.y);
                        tmp_in[q] = wa.x;
                    }
                }
                __syncthreads();
                if (sh.abort) { status = ST_ERR_BARRIER_TIMEOUT; break; }
#pragma unroll
                for (int i = 0; i < kEnt; ++i) {
                    const int q = tid + i * kTT;
                    if (q < ns) {                                                       // rank by counting (in[] values are distinct)
                        int rank = 0;
                        for (int r = 0; r < ns; ++r) rank += tmp_in[r] > e_in[i];
                        st_in[rank] = e_in[i]; st_z[rank] = e_z[i]; st_pd[rank] = e_pd[i]; st_fl[rank] = e_fl[i];
                    }
                }
                __syncthreads();
                TICK(t_stem);
            }
            if (change && tid == 0) { if (ns > sh.bk.max_stem) sh.bk.max_stem = ns; sh.bk.moved_nodes += s; }

            // ================================================================ updates
            const bool dir_new_up = u_in == a_src;                                      // NS.cs:1143
            if (pricer) {
                // arc states (ChangeFlow, NS.cs:1031-1039): only the pricing scan reads them
                if (change) { patch_arc0 = in_arc; patch_st0 = STATE_TREE; patch_arc1 = out.pd >> 1; patch_st1 = out.zero ? STATE_LOWER : STATE_UPPER; }
                else { patch_arc0 = in_arc; patch_st0 = -a_state; patch_arc1 = -1; }
                if (tid == 0) {
                    P.state[patch_arc0] = patch_st0;
                    if (patch_arc1 >= 0) P.state[patch_arc1] = patch_st1;
                }
            } else {
                if (!change && delta > 0 && tid == 0 && first >= lo && first < lo + cntn)
                    P.flow[in_arc] = (lower_state ? 0 : upper_in) + val;                // NS.cs:1018: stays a non-tree arc, at the other bound
                const int base = b < a ? b + 1 : b - s + 1;                             // new index of u_in: first child of v_in
                const long long piU = in_side1 ? piF : piS, piV = in_side1 ? piS : piF;
                const long long sigma = piV - piU - (dir_new_up ? (long long)a_cost : -(long long)a_cost);   // NS.cs:1187-1188
                // one fused pass: ChangeFlow (NS.cs:1012-1040), UpdateTreeStructure (:1042-1183), UpdatePotentials (:1185-1209)
                for (int j = tid; j < cntn; j += kTT) {
                    const int x = in_s[j], sz_u = sz_s[j];
                    const bool hasF = (unsigned)(inF - x) < (unsigned)sz_u;
                    const bool hasS = (unsigned)(inS - x) < (unsigned)sz_u;
                    if (hasF != hasS) {
                        const int pd = pd_s[j];
                        long long fl = fl_s[j];
                        if (delta > 0) {                                                // NS.cs:1020-1029
                            const long long dv = (pd & 1) ? val : -val;                 // pred_dir * val
                            fl = (hasF == src_side1) ? fl - dv : fl + dv;
                            fl_s[j] = fl;
                        }
                        if (change) {
                            if (hasF != in_side1) sz_s[j] = sz_u + s;                   // v_in .. join (NS.cs:1174-1177)
                            else if (x < a) sz_s[j] = sz_u - s;                         // v_out .. join (NS.cs:1179-1182)
                            else {                                                      // stem node (NS.cs:1095-1146)
                                if (x == a) P.flow[pd >> 1] = out.zero ? 0 : up_s[j];   // u_out: its pred arc leaves the tree at a bound
                                int kk = 0;
                                if (ns > 1) { int l = 0, r = ns - 1; while (l < r) { const int mid = (l + r) >> 1; if (st_in[mid] <= x) r = mid; else l = mid + 1; } kk = l; }
                                if (kk == 0) {
                                    pd_s[j] = in_arc * 2 + (dir_new_up ? 1 : 0); sz_s[j] = s;
                                    fl_s[j] = (lower_state ? 0 : upper_in) + val; up_s[j] = upper_in;
                                } else {
                                    const int npd = st_pd[kk - 1] ^ 1;
                                    pd_s[j] = npd; sz_s[j] = s - st_z[kk - 1];
                                    fl_s[j] = st_fl[kk - 1]; up_s[j] = __ldg(P.upper + (npd >> 1));
                                }
                            }
                        }
                    }
                    if (change) {
                        if ((unsigned)(x - a) < (unsigned)s) {
                            int off;
                            if (ns == 1) off = x - a;
                            else {
                                int l = 0, r = ns - 1;                                  // smallest k with x inside subtree(stem k)
                                while (l < r) { const int mid = (l + r) >> 1; if ((unsigned)(x - st_in[mid]) < (unsigned)st_z[mid]) r = mid; else l = mid + 1; }
                                if (l == 0) off = x - st_in[0];
                                else {
                                    int rr = x - st_in[l];
                                    if (x > st_in[l - 1]) rr -= st_z[l - 1];
                                    off = st_z[l - 1] + rr;
                                }
                            }
                            const int nx = base + off;
                            in_s[j] = nx;
                            atomicAdd(reinterpret_cast<unsigned long long*>(&P.node[lo + j].pi), (unsigned long long)sigma);
                            P.node[lo + j].in = nx;
                        } else if (b < a) {
                            if (x > b && x < a) { in_s[j] = x + s; P.node[lo + j].in = x + s; }
                        } else {
                            if (x >= a + s && x <= b) { in_s[j] = x - s; P.node[lo + j].in = x - s; }
                        }
                    }
                }
                // hop 3: everything this CTA wrote for pivot k is visible before DONE(k)
                PROBE(13);
                __syncthreads();
                if (tid == 0) { __threadfence(); st_vol_u32(P.done + (size_t)cta * 32, (unsigned)k); }
                PROBE(14);
            }
            TICK(t_update);
            PROBE(5);
        }
        if (P.stop_after > 0 && iterations >= P.stop_after) { status = ST_STOPPED_EARLY; break; }
    }
#undef TICK
    if (tid == 0 && (pricer || cta == 1)) for (int i = pricer ? 0 : 8; i < (pricer ? 8 : 16); ++i) P.ctl->clk[i] = sh.bk.pr[i];

    // =================================================================== epilogue
    const bool clean = status != ST_ERR_BARRIER_TIMEOUT;
    if (clean) {
        // flows of the tree arcs go back to flow[]; then one conventional grid barrier (counter + fences)
        for (int j = tid; j < cntn; j += kTT) { const int pd = pd_s[j]; if (pd >= 0) P.flow[pd >> 1] = fl_s[j]; }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(&P.ctl->bar, 1ULL);
            const long long t0 = clock64();
            while (*(volatile unsigned long long*)&P.ctl->bar < (unsigned long long)G) {
                if ((unsigned long long)(clock64() - t0) > P.timeout_cycles) { *(volatile int*)&P.ctl->abort = 1; break; }
            }
            __threadfence();
        }
        __syncthreads();
        for (int u = cta * kTT + tid; u < n; u += G * kTT) P.pi_out[u] = __ldcg(&P.node[u].pi);
    }
    if (status == ST_OPTIMAL && clean) {
        // CheckFeasibility (NS.cs:1272-1283) over arcs [m, m+n); GetTotalCost (NS.cs:452-465) over [0, m)
        int bad = 0;
        for (int e = P.m + cta * kTT + tid; e < S; e += G * kTT) bad |= __ldcg(P.flow + e) != 0;
        if (bad) atomicOr(&P.ctl->infeasible, 1);
        long long acc = 0;
        for (int e = cta * kTT + tid; e < P.m; e += G * kTT) {
            long long f = __ldcg(P.flow + e);
            if (P.orig_lower) { const long long l = __ldg(P.orig_lower + e); if (l != 0) { f += l; P.flow[e] = f; } }   // NS.cs:375-388
            acc += f * (long long)__ldg(P.cost + e);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && acc != 0) atomicAdd(reinterpret_cast<unsigned long long*>(&P.ctl->total_cost), (unsigned long long)acc);
    }
    if (pricer && tid == 0) {
        Ctl* c = P.ctl;
        const Book& bk = sh.bk;
        c->status = status; c->iterations = iterations; c->arcs_checked = bk.arcs_checked; c->final_block_size = B;
        c->degenerate = bk.degenerate; c->cycle_nodes = bk.cycle_nodes; c->moved_nodes = bk.moved_nodes;
        c->max_cycle = bk.max_cycle; c->max_stem = bk.max_stem; c->pricing_rounds = bk.rounds_total;
        c->ns_price = bk.t_price; c->ns_cycle = bk.t_cycle; c->ns_update = bk.t_update; c->ns_total = gtimer() - bk.t_begin;
        c->clk_total = (unsigned long long)clock64() - bk.c_begin;
        c->ns_wait_done = bk.t_wdone; c->ns_wait_cyc = 0; c->ns_stem = bk.t_stem; c->stem_exchanges = bk.stem_x;
    }
}

}  // namespace mcf

// ------------------------------------------------------------------------------------------------ launchers

namespace {
constexpr size_t kStemBytes = (size_t)mcf::kTeamStemCap * (8 + 4 * 4);
constexpr size_t kPricerBytes = (size_t)mcf::kPf * mcf::kTT * (8 + 4 * 4);
}  // namespace

extern "C" size_t mcfk_team_smem_bytes(int slice)
{
    const size_t owner = (size_t)slice * mcf::kNodeSmemBytes;
    return kStemBytes + (owner > kPricerBytes ? owner : kPricerBytes) + 16;
}

// largest slice (nodes per owner CTA) that fits the opt-in shared memory of the device next to the kernel's static part
extern "C" int mcfk_team_max_slice(int device)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, mcf::ns_team_kernel) != cudaSuccess) return -2;
    const long long avail = (long long)prop.sharedMemPerBlockOptin - (long long)fa.sharedSizeBytes - (long long)kStemBytes - 64;
    const long long s = avail / mcf::kNodeSmemBytes;
    return (int)(s & ~7LL);
}

extern "C" int mcfk_launch_team(const mcf::TeamParams* p, cudaStream_t stream)
{
    const size_t smem = mcfk_team_smem_bytes(p->slice);
    cudaError_t e = cudaFuncSetAttribute(mcf::ns_team_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    void* args[] = {(void*)p};
    e = cudaLaunchCooperativeKernel((const void*)mcf::ns_team_kernel, dim3(p->team), dim3(mcf::kTT), args, smem, stream);
    return (int)e;
}
