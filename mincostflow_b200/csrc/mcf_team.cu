// mcf_team.cu - the "team" pivot engine of libmcfgpu (sm_100a): Block Search network simplex as one persistent
// cooperative kernel in which the basis tree never leaves the chip.
//
// Why: a pivot of NetworkSimplex.Solve() (NS.cs:282-341) is a chain of pointer walks over parent/pred/thread/
// succ_num/last_succ (NS.cs:925-1209).  On B200 a dependent L2 load costs ~150 ns, a DRAM miss ~1 us and a grid-wide
// barrier ~1.3 us (profiles/r01_micro_latency.txt), so the walks are replaced by flat passes over an interval labelling
// (in[u] = DFS index, sz[u] = subtree size; see mcf_device.cuh) and the whole basis - labels, pred arcs, and the flow and
// capacity of every tree arc - is kept in SHARED MEMORY, sliced by node id over the "owner" CTAs of the team.  What has
// to cross between CTAs per pivot is then tiny, and it crosses as 16-byte words that carry their own sequence number
// (the pivot index) in the same 128-bit store - no fence, no barrier (profiles/r01_micro_hop.txt):
//
//   hop 1  ENTER   every pricer -> all   best candidate of its share of the block (arc, endpoints' (pi, in), cost, state, cap)
//   hop 2  CYC     every owner -> all    its best leaving-arc candidate per side of the cycle (+ counts)
//  (hop 2b STEM    every owner -> all    only when the re-hung stem is longer than one node: the stem entries)
//   hop 3  DONE    every owner -> pricers "my pi / in updates of this pivot are globally visible" (after one fence)
//
// The first CTAs ("pricers") run BlockSearchPivot.FindEnteringArc (NS.cs:1339-1441): the block of B arcs is split evenly
// over them (one SM alone is gather-throughput bound on a 3072-arc block), each prices its share against the global node
// mirror {pi, in}, and every CTA picks the same winner from the pricers' records.  The share of the next block is staged
// in shared memory while the other hops are in flight.  Owners run FindJoinNode + FindLeavingArc as an interval test over
// their slice, every CTA reduces the candidates redundantly to the same decision (strict '<' on the first walk, '<=' on
// the second, NS.cs:958-998), owners apply ChangeFlow / UpdateTreeStructure / UpdatePotentials (NS.cs:1012-1209) to the
// nodes they own in ONE fused pass.  Flows of tree arcs live with the node below the arc; flow[] in global memory is
// written when an arc leaves the tree and at the end.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "mcf_device.cuh"

namespace mcf {

namespace {

constexpr int kTT = 512;                                // threads per CTA: latency-bound code, up to 128 registers each
constexpr int kTW = kTT / 32;
#ifndef MCF_PF
#define MCF_PF 4
#endif
constexpr int kPf = MCF_PF;                             // arcs per pricer thread staged ahead (kPf x 512 per pricing CTA)
#ifndef MCF_REP_ENT
#define MCF_REP_ENT 4
#endif
#ifndef MCF_REP_CYC
#define MCF_REP_CYC 6
#endif
constexpr int kRepEnt = MCF_REP_ENT;                    // replicas of every ENTER record: a reader polls replica (cta % kRepEnt)
constexpr int kRepCyc = MCF_REP_CYC;                    // replicas of every CYC record (fewer pollers per line, profiles/r01_micro_hop.txt)
constexpr int kRelUnroll = 4;                            // nodes per thread in flight in the relabel pass
constexpr int kCandCap = 32;                            // cycle nodes of one slice handled by the single-warp path

__device__ __forceinline__ int4 ld_vol4(const int4* p)
{
    int4 v;
    // one 128-bit relaxed.gpu access: single-copy atomic by the PTX memory model (ISA 8.3+), which .volatile.v4 is not
    asm volatile("{\n .reg .b128 t;\n ld.relaxed.gpu.global.b128 t, [%4];\n mov.b128 {%0,%1,%2,%3}, t;\n}" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol4(int4* p, int4 v)
{
    asm volatile("{\n .reg .b128 t;\n mov.b128 t, {%1,%2,%3,%4};\n st.relaxed.gpu.global.b128 [%0], t;\n}" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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
// asynchronous global -> shared copies (LDGSTS): immutable arc data streams from DRAM without holding registers
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ int lo32(long long v) { return (int)(unsigned)(unsigned long long)v; }
__device__ __forceinline__ int hi32(long long v) { return (int)(unsigned)((unsigned long long)v >> 32); }
__device__ __forceinline__ long long mk64(int lo, int hi) { return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo); }

// Lane holding the lexicographic minimum of (a, b) among the lanes with `valid` (-1 when none); three REDUX steps
// instead of a shuffle tree.  All 32 lanes must call.
__device__ __forceinline__ int warp_argmin(bool valid, long long a, int b)
{
    const unsigned m = 0xffffffffu;
    if (!__any_sync(m, valid)) return -1;
    const int hi = valid ? hi32(a) : INT_MAX;
    const int mh = __reduce_min_sync(m, hi);
    bool c = valid && hi == mh;
    const unsigned lo = c ? (unsigned)lo32(a) : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(m, lo);
    c = c && lo == ml;
    const int bb = c ? b : INT_MAX;
    const int mb = __reduce_min_sync(m, bb);
    return __ffs(__ballot_sync(m, c && bb == mb)) - 1;
}

struct Cand {                       // leaving-arc candidate of one side of the cycle
    long long d;                    // residual in cycle direction
    int in, sz, pd;                 // labels and pred word of the node below the candidate arc
    int zero;                       // bit 0: flow on the arc is 0 after the augmentation (-> STATE_LOWER); bit 1: side 1
    int dp, j;                      // depth of the node; its index in the owner's slice
};

struct PWin {                       // a pricing candidate
    long long rc;
    int off;                        // scan offset from next_arc (< 0: none)
    int arc, src, tgt, cost, state, in_s, in_t, dp_s, dp_t;
    int blk;                        // block of the scan the candidate lies in (later rounds of a search)
    long long pi_s, pi_t, upper;
};

struct Book {                       // statistics and timers: touched by thread 0 only, kept out of the register file
    long long arcs_checked, rounds_total, degenerate, cycle_nodes, moved_nodes, max_cycle, max_stem, stem_x;
    unsigned long long t_price, t_cycle, t_update, t_wdone, t_stem, t_mark, t_begin, c_begin, pr_mark;
    unsigned long long pr[16];
    int cons_low, cons_high;
};

struct Pending {                    // one pivot's update in closed form: what the pricers replay on staged node records
    int valid, change;
    int a, s, b;                    // re-hung subtree = old interval [a, a+s); b = in[v_in]
    int ns, longstem, dshift, par, seq;
    long long sigma;
};

struct TeamShared {
    Book bk;
    Cand cl[kCandCap];              // cycle-node candidates of this slice (owner scan)
    Cand wc[2][kTW];                // per-warp winners (hop 2 gather, owner slow path)
    PWin pw[kTW];                   // per-warp pricing winners
    PWin win;                       // entering arc of this pivot
    int4 rec[kMaxPricers][7];       // pricing records as received
    int ncand, nstem, abort, cnt;
    int wide_req;                   // narrow mode: a flow of this slice left the int32 range; travels in the next CYC record
    int pre[kTeamMax + 1];
};

// time-out / abort check for spin loops; true = give up
__device__ __forceinline__ bool spin_check(unsigned& spins, long long& t0, const TeamParams& P)
{
#ifdef MCF_SPIN_SLEEP
    __nanosleep(MCF_SPIN_SLEEP);                        // back off: fewer polling requests in flight at L2
#endif
    if ((++spins & 255u) != 0) return false;
    if (t0 == 0) { t0 = clock64(); return false; }
    if (*(volatile int*)&P.ctl->abort) return true;
    if ((unsigned long long)(clock64() - t0) > P.timeout_cycles) { *(volatile int*)&P.ctl->abort = 1; return true; }
    return false;
}

// poll one self-validating word until its sequence number matches; false = abandoned (abort flag or time-out)
__device__ __forceinline__ bool poll_word(const int4* p, int seq, int4& out, const TeamParams& P)
{
    unsigned spins = 0; long long t0 = 0;
    for (;;) {
        const int4 v = ld_vol4(p);
        if (v.w == seq) { out = v; return true; }
        if (spin_check(spins, t0, P)) { out = v; return false; }
    }
}

// poll NW words of one record, all loads in flight together
template <int NW>
__device__ __forceinline__ bool poll_rec(const int4* rec, int seq, int4 (&w)[NW], const TeamParams& P)
{
    unsigned spins = 0; long long t0 = 0;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = ld_vol4(rec + i);
#pragma unroll
        for (int i = 0; i < NW; ++i) ok = ok && w[i].w == seq;
        if (ok) return true;
        if (spin_check(spins, t0, P)) return false;
    }
}

__device__ __forceinline__ void post_pwin(int4* rec, const PWin& w, int round, int seq, int lane)
{
    int4 o;
    if (lane == 0) o = make_int4(w.off >= 0 ? w.arc : -1, w.src, w.tgt, seq);
    else if (lane == 1) o = make_int4(w.cost, w.state, round, seq);
    else if (lane == 2) o = make_int4(lo32(w.pi_s), hi32(w.pi_s), w.in_s, seq);
    else if (lane == 3) o = make_int4(lo32(w.pi_t), hi32(w.pi_t), w.in_t, seq);
    else if (lane == 4) o = make_int4(lo32(w.upper), hi32(w.upper), w.off, seq);
    else if (lane == 5) o = make_int4(lo32(w.rc), hi32(w.rc), w.blk, seq);
    else o = make_int4(w.dp_s, w.dp_t, 0, seq);
    st_vol4(rec + lane, o);
}
__device__ __forceinline__ PWin unpack_pwin(const int4* r)
{
    PWin w;
    w.arc = r[0].x; w.src = r[0].y; w.tgt = r[0].z; w.cost = r[1].x; w.state = r[1].y;
    w.pi_s = mk64(r[2].x, r[2].y); w.in_s = r[2].z; w.pi_t = mk64(r[3].x, r[3].y); w.in_t = r[3].z;
    w.upper = mk64(r[4].x, r[4].y); w.off = r[0].x >= 0 ? r[4].z : -1; w.rc = mk64(r[5].x, r[5].y);
    w.dp_s = r[6].x; w.dp_t = r[6].y; w.blk = r[5].z;
    return w;
}
__device__ __forceinline__ PWin pwin_none()
{
    PWin w; w.rc = 0; w.off = -1; w.arc = -1; w.src = w.tgt = w.cost = w.state = w.in_s = w.in_t = w.dp_s = w.dp_t = w.blk = 0; w.pi_s = w.pi_t = w.upper = 0;
    return w;
}
__device__ __forceinline__ Cand cand_none() { Cand c; c.d = 0; c.in = c.sz = c.zero = c.dp = c.j = 0; c.pd = -1; return c; }

template <typename F> struct FlowTraits;
template <> struct FlowTraits<int> {
    __device__ static __forceinline__ int cap_in(long long up) { return up >= (long long)INT_MAX ? INT_MAX : (int)up; }        // INT_MAX = uncapacitated
    __device__ static __forceinline__ long long residual(int up, int fl) { return up == INT_MAX ? (LLONG_MAX / 2) - fl : (long long)up - fl; }
    __device__ static __forceinline__ bool fits(long long v) { return v >= 0 && v < (long long)INT_MAX; }
    __device__ static __forceinline__ long long cap_out(int up) { return up == INT_MAX ? (LLONG_MAX / 2) : (long long)up; }     // the host admits only INF or < 2^31-1
};
template <> struct FlowTraits<long long> {
    __device__ static __forceinline__ long long cap_in(long long up) { return up; }
    __device__ static __forceinline__ long long residual(long long up, long long fl) { return up == LLONG_MAX ? (LLONG_MAX / 2) : up - fl; }   // NS.cs:970-971
    __device__ static __forceinline__ bool fits(long long) { return true; }
    __device__ static __forceinline__ long long cap_out(long long up) { return up; }
};

}  // namespace

template <typename F>
__global__ void __launch_bounds__(kTT, 1) ns_team_kernel(const TeamParams P)
{
    using FT = FlowTraits<F>;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ TeamShared sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = P.team, NP = P.pricers, cta = blockIdx.x, nown = G - NP;
    const int n = P.n, S = P.S;
    const bool pricer = cta < NP;
    const int own = cta - NP;
    const int lo = pricer ? 0 : own * P.slice;
    const int cntn = pricer ? 0 : max(0, min(n + 1, lo + P.slice) - lo);

    // dynamic shared memory.  Everybody: stem staging.  Owners: the resident slice.  Pricers: the staged share of the next block.
    long long* const st_fl = reinterpret_cast<long long*>(dyn_smem);            // [kTeamStemCap] flow on stem k's old pred arc (after augmentation)
    int* const st_in = reinterpret_cast<int*>(st_fl + kTeamStemCap);            // sorted: stem 0 = u_in (deepest) .. u_out
    int* const st_z = st_in + kTeamStemCap;
    int* const st_pd = st_z + kTeamStemCap;
    int* const st_up = st_pd + kTeamStemCap;                                    // capacity of that arc (INT_MAX: infinite, -1: fetch)
    unsigned char* const body = reinterpret_cast<unsigned char*>(st_up + kTeamStemCap);
    // owners
    // (spill mode: the two arrays below are this owner's part of a global array instead - same code, generic addressing; only the
    // cycle nodes of a pivot ever touch them, and only their owner does)
    F* const fl_s = P.spill ? reinterpret_cast<F*>(P.fl_g) + lo : reinterpret_cast<F*>(body);                 // flow on the pred arc of node j
    F* const up_s = P.spill ? reinterpret_cast<F*>(P.up_g) + lo : reinterpret_cast<F*>(body) + P.slice;       // capacity of the pred arc
    int* const in_s = P.spill ? reinterpret_cast<int*>(body) : reinterpret_cast<int*>(reinterpret_cast<F*>(body) + 2 * (size_t)P.slice);
    int* const sz_s = in_s + P.slice;
    int* const pd_s = sz_s + P.slice;
    int* const dp_s = pd_s + P.slice;                                           // depth in the basis tree
    // spill mode: the depth is read from the node mirror, which this CTA alone writes (NodeRec::dp), instead of a resident copy
    const bool dp_res = !(MCF_SPILL_DP && P.spill);
    auto dp_get = [&](int j) -> int { return dp_res ? dp_s[j] : P.node[lo + j].dp; };
    // pricers: arc data and both ends' node records of this pricer's share of the staged block
    constexpr int kStage = kPf * kTT;
    long long* const pf_up = reinterpret_cast<long long*>(body);
    long long* const pf_pis = pf_up + kStage;
    long long* const pf_pit = pf_pis + kStage;
    int* const pf_src = reinterpret_cast<int*>(pf_pit + kStage);
    int* const pf_tgt = pf_src + kStage;
    int* const pf_cost = pf_tgt + kStage;
    int* const pf_st = pf_cost + kStage;
    int* const pf_ins = pf_st + kStage;
    int* const pf_int = pf_ins + kStage;
    int* const pf_dps = pf_int + kStage;
    int* const pf_dpt = pf_dps + kStage;

    int status = ST_NOT_SOLVED;
    {
        int bad = 0;
        for (int j = tid; j < cntn; j += kTT) {
            const int u = lo + j;
            const int pd = P.pd0[u];
            in_s[j] = P.in_g[u]; if (dp_res) dp_s[j] = P.node[u].dp; sz_s[j] = P.sz0[u]; pd_s[j] = pd;
            const long long fl = pd >= 0 ? P.flow[pd >> 1] : 0, up = pd >= 0 ? P.upper[pd >> 1] : 0;
            bad |= !FT::fits(fl);
            fl_s[j] = (F)fl; up_s[j] = FT::cap_in(up);
        }
        if (!pricer) for (int j = cntn + tid; j < P.slice; j += kTT) { in_s[j] = 0; sz_s[j] = 0; if (dp_res) dp_s[j] = 0; pd_s[j] = -2; fl_s[j] = 0; up_s[j] = 0; }   // padding: on no cycle, never relabelled
        if (tid == 0) { sh.abort = 0; sh.wide_req = 0; Book z = {}; sh.bk = z; }
        if (__syncthreads_or(bad)) { if (tid == 0) { P.ctl->needs_wide = 1; sh.wide_req = 1; } }
    }
    if (tid == 0) { sh.bk.t_begin = gtimer(); sh.bk.c_begin = sh.bk.t_mark = sh.bk.pr_mark = (unsigned long long)clock64(); }

    // pricer state (BlockSearchPivot fields, NS.cs:1294-1302); identical in every pricer
    int next_arc = 0, B = P.block_size;
    int pf_next = -1, pf_B = 0;                          // what is staged: this pricer's share of block [pf_next, pf_next + pf_B)
    long long pf_upto = 0;                               // ... as of "all updates of pivots <= pf_upto applied"
    long long done_seen = 0;                             // DONE(j) observed from every owner for all j <= done_seen
    int spec_cursor = -1, spec_B = 0;                    // arc data of block [spec_cursor, +spec_B) is in flight into the staging area
    int stv[kPf];                                        // ... with its arc states here (see stage_static_begin)
#pragma unroll
    for (int j = 0; j < kPf; ++j) stv[j] = 0;
    // arc-state changes of the last two pivots: applied on top of whatever a scan reads (staged or global), newest first, so a
    // scan never depends on how fast this CTA's own state[] stores become visible to its other warps
    int patch_arc0 = -1, patch_st0 = 0, patch_arc1 = -1, patch_st1 = 0, patch2_arc0 = -1, patch2_st0 = 0, patch2_arc1 = -1, patch2_st1 = 0;
    auto fix_state = [&](int idx, int st) -> int {
        return idx == patch_arc0 ? patch_st0 : idx == patch_arc1 ? patch_st1 : idx == patch2_arc0 ? patch2_st0 : idx == patch2_arc1 ? patch2_st1 : st;
    };
    long long iterations = 0;
    Pending Uprev;                                       // the previous pivot's update (every thread computes it; see `replay`)
    Uprev.valid = Uprev.change = Uprev.a = Uprev.s = Uprev.b = Uprev.longstem = Uprev.dshift = Uprev.par = Uprev.seq = 0; Uprev.ns = 1; Uprev.sigma = 0;
#define PROBE(i) do { if (tid == 0 && ((i) < 8 ? cta == 0 : cta == NP)) { const unsigned long long t__ = (unsigned long long)clock64(); sh.bk.pr[i] += t__ - sh.bk.pr_mark; sh.bk.pr_mark = t__; } } while (0)
#define TICK(acc) do { if (cta == 0 && tid == 0) { const unsigned long long t__ = (unsigned long long)clock64(); sh.bk.acc += t__ - sh.bk.t_mark; sh.bk.t_mark = t__; } } while (0)

    // pricers: wait until every owner's updates of pivots <= j are visible (hop 3); CTA-uniform, false = abandoned
    auto wait_done = [&](long long j) -> bool {
        if (done_seen >= j) return true;
        if (tid < nown) {
            const unsigned want = (unsigned)j;
            const unsigned* p = P.done + (size_t)(NP + tid) * 32;
            unsigned spins = 0; long long t0 = 0;
            while ((int)(ld_vol_u32(p) - want) < 0) if (spin_check(spins, t0, P)) { sh.abort = 1; break; }
        }
        __syncthreads();
        done_seen = j;
        return sh.abort == 0;
    };
    // this pricer's share [s_lo, s_hi) of the block of `bsz` arcs that starts at the cursor
    auto share = [&](int bsz, int& s_lo, int& s_hi) -> bool {
        const int blk0 = bsz < S ? bsz : S;
        const int seg = (blk0 + NP - 1) / NP;
        s_lo = cta * seg; s_hi = min(blk0, s_lo + seg);
        return seg <= kStage;
    };
    // stage the arc data of the share (streams from DRAM; independent of the basis except `state`, which is patched later)
    auto stage_static = [&](int cursor, int s_lo, int s_hi) {
#pragma unroll
        for (int j = 0; j < kPf; ++j) {
            const int q = tid + j * kTT, off = s_lo + q;
            if (off < s_hi) {
                int idx = cursor + off; if (idx >= S) idx -= S;
                pf_src[q] = __ldg(P.src + idx); pf_tgt[q] = __ldg(P.tgt + idx); pf_cost[q] = __ldg(P.cost + idx);
                pf_st[q] = __ldcg(P.state + idx); pf_up[q] = __ldg(P.upper + idx);
            }
        }
    };
    // the same, asynchronously: src / tgt / cost / capacity are immutable and go global -> shared with cp.async; `state` is
    // mutable, so it is read around L1 into `stv` and stored by stage_static_finish()
    auto stage_static_begin = [&](int cursor, int s_lo, int s_hi, int (&stv)[kPf]) {
#pragma unroll
        for (int j = 0; j < kPf; ++j) {
            const int q = tid + j * kTT, off = s_lo + q;
            stv[j] = 0;
            if (off < s_hi) {
                int idx = cursor + off; if (idx >= S) idx -= S;
                cp_async4(pf_src + q, P.src + idx); cp_async4(pf_tgt + q, P.tgt + idx); cp_async4(pf_cost + q, P.cost + idx);
                cp_async8(pf_up + q, P.upper + idx);
                stv[j] = __ldcg(P.state + idx);
            }
        }
    };
    auto stage_static_finish = [&](int s_lo, int s_hi, const int (&stv)[kPf]) {
        cp_async_wait_all();
#pragma unroll
        for (int j = 0; j < kPf; ++j) { const int q = tid + j * kTT; if (s_lo + q < s_hi) pf_st[q] = stv[j]; }
    };
    // gather both ends' node records {pi, in, depth} of the staged share from the mirror
    auto stage_gather = [&](int s_lo, int s_hi) {
#pragma unroll
        for (int j = 0; j < kPf; ++j) {
            const int q = tid + j * kTT, off = s_lo + q;
            if (off < s_hi) {
                const int4 rs = __ldcg(reinterpret_cast<const int4*>(P.node + pf_src[q]));
                const int4 rt = __ldcg(reinterpret_cast<const int4*>(P.node + pf_tgt[q]));
                const int is = __ldcg(P.in_g + pf_src[q]), it = __ldcg(P.in_g + pf_tgt[q]);
                pf_pis[q] = mk64(rs.x, rs.y); pf_ins[q] = is; pf_dps[q] = rs.w;
                pf_pit[q] = mk64(rt.x, rt.y); pf_int[q] = it; pf_dpt[q] = rt.w;
            }
        }
    };
    // closed-form re-labelling of one node by the update described in `U` (UpdateTreeStructure seen through in[] / depth):
    // nodes of the re-hung subtree [a, a+s) get their new place under v_in, nodes between the old and new place shift by s
    auto relabel = [&](const Pending& U, int x, int dp, int& nx, int& ndp) -> bool {
        const int4* const stem_g = P.stemseg + (size_t)U.par * (n + 1) * 2;
        auto stem_io = [&](int kx, int& o_in, int& o_z) {
            if (!U.longstem) { o_in = st_in[kx]; o_z = st_z[kx]; }
            else { int4 w; if (!poll_word(stem_g + (size_t)(U.ns - 1 - kx) * 2, U.seq, w, P)) sh.abort = 1; o_in = w.x; o_z = w.y; }
        };
        nx = x; ndp = dp;
        if ((unsigned)(x - U.a) < (unsigned)U.s) {
            int off, l = 0;
            if (U.ns == 1) off = x - U.a;
            else {
                int r = U.ns - 1, l_in, l_z;                                    // smallest l with x inside subtree(stem l)
                while (l < r) { const int mid = (l + r) >> 1; stem_io(mid, l_in, l_z); if ((unsigned)(x - l_in) < (unsigned)l_z) r = mid; else l = mid + 1; }
                stem_io(l, l_in, l_z);
                if (l == 0) off = x - l_in;
                else {
                    int p_in, p_z; stem_io(l - 1, p_in, p_z);
                    int rr = x - l_in;
                    if (x > p_in) rr -= p_z;
                    off = p_z + rr;
                }
            }
            nx = (U.b < U.a ? U.b + 1 : U.b - U.s + 1) + off;                   // u_in becomes the first child of v_in
            ndp = dp + U.dshift + 2 * l;
            return true;
        }
        const int sh_lo = U.b < U.a ? U.b + 1 : U.a + U.s, sh_len = U.b < U.a ? U.a - U.b - 1 : U.b - U.a - U.s + 1;
        if ((unsigned)(x - sh_lo) < (unsigned)sh_len) nx = x + (U.b < U.a ? U.s : -U.s);
        return false;
    };

    if (pricer) {                                        // stage the very first block; the initial basis is what the host uploaded
        int s_lo, s_hi;
        if (share(B, s_lo, s_hi)) { stage_static(0, s_lo, s_hi); stage_gather(s_lo, s_hi); pf_next = 0; pf_B = B; pf_upto = 0; }
    }

    for (;;) {
        const long long k = iterations + 1;
        const int seq = (int)(unsigned)k;
        const int par = (int)(k & 1);
        bool have_win = false;                            // sh.win holds the entering arc
        int search_end = 0;                               // pricers: scan offset just past the winning block

        // ================================================================ pricers: price their share of the first block, post
        if (pricer) {
            TICK(t_wdone);
            if (tid == 0 && cta == 0) sh.bk.pr_mark = (unsigned long long)clock64();
            // ---- round 0 of BlockSearchPivot.FindEnteringArc (NS.cs:1339-1397): the first block, split over the pricers.
            // The share was staged (arc data + both ends' node records) BEFORE the previous pivot's update was applied; that
            // one update is replayed here from its closed form (sh.pend), so pricing does not wait for hop 3.
            int s_lo, s_hi;
            const bool fits = share(B, s_lo, s_hi);
            PWin best = pwin_none();
            if (fits) {
                if (!(pf_next == next_arc && pf_B == B)) [[unlikely]] {               // nothing usable staged (first use of a new block size): stage now
                    if (!wait_done(k - 1)) { status = ST_ERR_BARRIER_TIMEOUT; break; }
                    stage_static(next_arc, s_lo, s_hi); stage_gather(s_lo, s_hi);
                    pf_next = next_arc; pf_B = B; pf_upto = k - 1;
                }
                const bool replay = pf_upto < k - 1;                     // exactly update k-1 is missing from the staged records
                const Pending U = Uprev;
                int bq = -1;
#pragma unroll
                for (int j = 0; j < kPf; ++j) {
                    const int q = tid + j * kTT, off = s_lo + q;
                    if (off < s_hi) {
                        int idx = next_arc + off; if (idx >= S) idx -= S;
                        const int st = fix_state(idx, pf_st[q]);
                        long long ps = pf_pis[q], pt = pf_pit[q];
                        if (replay && U.change) {                        // UpdatePotentials of the pending pivot (NS.cs:1185-1209)
                            if ((unsigned)(pf_ins[q] - U.a) < (unsigned)U.s) ps += U.sigma;
                            if ((unsigned)(pf_int[q] - U.a) < (unsigned)U.s) pt += U.sigma;
                        }
                        const long long rc = (long long)st * ((long long)pf_cost[q] + ps - pt);
                        if (rc < best.rc) { best.rc = rc; best.off = off; best.arc = idx; best.state = st; best.pi_s = ps; best.pi_t = pt; bq = q; }
                    }
                }
                const int wl = warp_argmin(best.off >= 0, best.rc, best.off);
                if (wl < 0) { if (lane == 0) sh.pw[warp].off = -1; }
                else if (lane == wl) {
                    best.src = pf_src[bq]; best.tgt = pf_tgt[bq]; best.cost = pf_cost[bq]; best.upper = pf_up[bq];
                    best.in_s = pf_ins[bq]; best.in_t = pf_int[bq]; best.dp_s = pf_dps[bq]; best.dp_t = pf_dpt[bq];
                    if (replay && U.change) {                            // the winner's labels as they are after the pending update
                        int nx, nd;
                        relabel(U, best.in_s, best.dp_s, nx, nd); best.in_s = nx; best.dp_s = nd;
                        relabel(U, best.in_t, best.dp_t, nx, nd); best.in_t = nx; best.dp_t = nd;
                    }
                    sh.pw[warp] = best;
                }
            } else [[unlikely]] {
                // share larger than the staging area (B > pricers x 2048): price straight from global memory
                if (!wait_done(k - 1)) { status = ST_ERR_BARRIER_TIMEOUT; break; }
                for (int off = s_lo + tid; off < s_hi; off += kTT) {
                    int idx = next_arc + off; if (idx >= S) idx -= S;
                    const int s = __ldg(P.src + idx), t = __ldg(P.tgt + idx), c = __ldg(P.cost + idx);
                    const int st = fix_state(idx, __ldcg(P.state + idx));
                    const long long up = __ldg(P.upper + idx);
                    const int4 rs = __ldcg(reinterpret_cast<const int4*>(P.node + s));
                    const int4 rt = __ldcg(reinterpret_cast<const int4*>(P.node + t));
                    const long long ps = mk64(rs.x, rs.y), pt = mk64(rt.x, rt.y);
                    const long long rc = (long long)st * ((long long)c + ps - pt);
                    if (rc < best.rc) {
                        best.rc = rc; best.off = off; best.arc = idx; best.src = s; best.tgt = t; best.cost = c; best.state = st;
                        best.in_s = __ldcg(P.in_g + s); best.in_t = __ldcg(P.in_g + t); best.dp_s = rs.w; best.dp_t = rt.w; best.pi_s = ps; best.pi_t = pt; best.upper = up;
                    }
                }
                const int wl = warp_argmin(best.off >= 0, best.rc, best.off);
                if (wl < 0) { if (lane == 0) sh.pw[warp].off = -1; }
                else if (lane == wl) sh.pw[warp] = best;
            }
            __syncthreads();
            if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
            if (warp == 0) {                                             // CTA arg-min of (rc, off) over the warp winners, post the record
                const PWin* q = &sh.pw[lane & (kTW - 1)];
                const int ww = warp_argmin(lane < kTW && q->off >= 0, q->rc, q->off);
                PWin mine = pwin_none();
                if (ww >= 0) mine = sh.pw[ww];
                if (lane < 7 * kRepEnt) post_pwin(P.ent0 + (((size_t)par * kRepEnt + lane / 7) * NP + cta) * kMailWords, mine, 0, seq, lane % 7);
            }
            // 98.5 % of the searches end in this first block: the next pivot's block then starts at its last arc (NS.cs:1397).
            // Start streaming that block's arc data now, under the collect hop; it is re-done in the rare other case.
            spec_cursor = -1;
            if (fits) {
                const int blk0 = B < S ? B : S;
                int e = next_arc;
                if (blk0 < S || (long long)S % B == 0) { e = next_arc + blk0 - 1; if (e >= S) e -= S; }
                stage_static_begin(e, s_lo, s_hi, stv);                  // (every read of the staging area is behind the barrier above)
                spec_cursor = e; spec_B = B;
            }
            PROBE(1);
        }

        // ================================================================ all: hop 1, collect the pricers' round-0 records
        if (tid < NP * 7) {
            const int p = tid / 7, w = tid - p * 7;
            int4 v;
            if (!poll_word(P.ent0 + (((size_t)par * kRepEnt + cta % kRepEnt) * NP + p) * kMailWords + w, seq, v, P)) sh.abort = 1;
            sh.rec[p][w] = v;
        }
        int done_ok = 1;
        if (pricer && k > 1 && done_seen < k - 1 && tid >= 128 && tid < 128 + nown)
            // pricers: one look at DONE(k-1) under the same hop - if every owner is through, the next block's node records can
            // be gathered without another round trip (otherwise wait_done() below polls)
            done_ok = (int)(ld_vol_u32(P.done + (size_t)(NP + tid - 128) * 32) - (unsigned)(k - 1)) >= 0;
        done_ok = __syncthreads_and(done_ok);
        if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
        if (pricer && k > 1 && done_seen < k - 1 && done_ok) done_seen = k - 1;
        int win_rec = -1;                                 // round-0 record that holds the entering arc (every thread decodes it itself)
        {
            // arg-min of (reduced cost, scan offset) over the pricers' records, redundantly in every warp
            const bool v = lane < NP && sh.rec[lane < NP ? lane : 0][0].x >= 0;
            const int pl = lane < NP ? lane : 0;
            win_rec = warp_argmin(v, mk64(sh.rec[pl][5].x, sh.rec[pl][5].y), sh.rec[pl][4].z);
            if (win_rec >= 0) { have_win = true; search_end = B < S ? B : S; }
        }
        if (!have_win) [[unlikely]] {
            if (pricer) {
                // ---- later rounds, straight from global memory (every update visible first): in round r each pricer prices M
                // consecutive blocks (M = 1, 2, 4, ... while NP*M <= 16), NP*M blocks per exchange; the lowest block with a negative reduced
                // cost wins, inside it the smallest reduced cost, then the first in scan order
                if (!wait_done(k - 1)) { status = ST_ERR_BARRIER_TIMEOUT; break; }
                const long long nblk = ((long long)S + B - 1) / B;
                long long next_blk = 1;
                int found_blk = -1;
                for (int r = 1; found_blk < 0; ++r) {
                    if (next_blk >= nblk) break;
                    const int mcap = NP >= 16 ? 1 : 16 / NP;                         // about 16 blocks per exchange at most
                    const int M = min(mcap, r < 4 ? 1 << (r - 1) : 8);
                    const long long b_lo = next_blk + (long long)cta * M, b_hi = min(nblk, b_lo + M);
                    PWin best = pwin_none();
                    if (b_lo < nblk) {
                        const long long o_lo = b_lo * B; long long o_hi = b_hi * B; if (o_hi > S) o_hi = S;
                        for (long long off = o_lo + tid; off < o_hi; off += kTT) {
                            const int blk = M == 1 ? (int)b_lo : (int)(off / B);
                            if (best.off >= 0 && blk > best.blk) break;             // a thread's offsets ascend: later blocks cannot win
                            int idx = next_arc + (int)off; if (idx >= S) idx -= S;
                            const int s = __ldg(P.src + idx), t = __ldg(P.tgt + idx), c = __ldg(P.cost + idx);
                            const int st = fix_state(idx, __ldcg(P.state + idx));
                            const int4 rs = __ldcg(reinterpret_cast<const int4*>(P.node + s));
                            const int4 rt = __ldcg(reinterpret_cast<const int4*>(P.node + t));
                            const long long ps = mk64(rs.x, rs.y), pt = mk64(rt.x, rt.y);
                            const long long rc = (long long)st * ((long long)c + ps - pt);
                            if (rc < best.rc) {
                                best.rc = rc; best.off = (int)off; best.blk = blk; best.arc = idx; best.src = s; best.tgt = t; best.cost = c; best.state = st;
                                best.in_s = __ldcg(P.in_g + s); best.in_t = __ldcg(P.in_g + t); best.dp_s = rs.w; best.dp_t = rt.w; best.pi_s = ps; best.pi_t = pt;
                            }
                        }
                    }
                    __syncthreads();                                    // sh.pw / sh.rec of the previous round are consumed
                    {   // lowest block first, then (rc, off)
                        const int mb = __reduce_min_sync(0xffffffffu, best.off >= 0 ? best.blk : INT_MAX);
                        const int wl = warp_argmin(best.off >= 0 && best.blk == mb, best.rc, best.off);
                        if (wl < 0) { if (lane == 0) sh.pw[warp].off = -1; }
                        else if (lane == wl) { best.upper = __ldg(P.upper + best.arc); sh.pw[warp] = best; }
                    }
                    __syncthreads();
                    if (warp == 0) {
                        const PWin* q = &sh.pw[lane & (kTW - 1)];
                        const bool qv = lane < kTW && q->off >= 0;
                        const int mb = __reduce_min_sync(0xffffffffu, qv ? q->blk : INT_MAX);
                        const int ww = warp_argmin(qv && q->blk == mb, q->rc, q->off);
                        PWin mine = pwin_none();
                        if (ww >= 0) mine = sh.pw[ww];
                        // words 0, 2..6 first, one fence, then word 1 (which carries the round) as the flag: a round record
                        // reuses the slot of round r-2 of the same pivot, so the sequence number alone cannot validate it
                        int4* const dst = P.prc + ((size_t)(r & 1) * NP + cta) * kMailWords;
                        if (lane < 7 && lane != 1) post_pwin(dst, mine, r, seq, lane);
                        __syncwarp();
                        if (lane == 1) { __threadfence(); post_pwin(dst, mine, r, seq, 1); }
                    }
                    if (tid < NP * 7) {                                 // round records carry (seq, round)
                        const int p = tid / 7, w = tid - p * 7;
                        const int4* src = P.prc + ((size_t)(r & 1) * NP + p) * kMailWords;
                        unsigned spins = 0; long long t0 = 0;
                        for (;;) {
                            const int4 v1 = ld_vol4(src + 1);
                            if (v1.w == seq && v1.z == r) { sh.rec[p][w] = w == 1 ? v1 : ld_vol4(src + w); break; }
                            if (spin_check(spins, t0, P)) { sh.abort = 1; break; }
                        }
                    }
                    __syncthreads();
                    if (sh.abort) break;
                    if (tid == 0) sh.bk.rounds_total++;
                    // pricer p holds blocks below those of pricer p+1: the first record with a candidate is the lowest block
                    for (int p = 0; p < NP; ++p) if (sh.rec[p][0].x >= 0) { found_blk = sh.rec[p][5].z; if (tid == 0) sh.win = unpack_pwin(sh.rec[p]); break; }
                    next_blk += (long long)NP * M;
                }
                if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
                if (found_blk >= 0) {
                    long long e = ((long long)found_blk + 1) * B; if (e > S) e = S;
                    search_end = (int)e; have_win = true;
                } else search_end = S;
                __syncthreads();
                if (cta == 0 && warp == 0) {
                    PWin w = have_win ? sh.win : pwin_none();
                    if (lane < 7) post_pwin(P.late + (size_t)par * kMailWords, w, 0, seq, lane);
                }
            } else {
                if (tid < 7) {
                    int4 v;
                    if (!poll_word(P.late + (size_t)par * kMailWords + tid, seq, v, P)) sh.abort = 1;
                    sh.rec[0][tid] = v;
                }
                __syncthreads();
                if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
                have_win = sh.rec[0][0].x >= 0;
                if (have_win && tid == 0) sh.win = unpack_pwin(sh.rec[0]);
            }
            __syncthreads();
        }
        if (pricer) {
            // NS.cs:1397-1438: cursor, counters, adaptive block size - every pricer keeps the same copy
            if (tid == 0) { sh.bk.arcs_checked += search_end; sh.bk.rounds_total++; }
            if (have_win) {
                const int Bold = B;
                if (P.adaptive) {
                    const double hit = search_end > 0 ? 1.0 / search_end : 0;
                    int cl = sh.bk.cons_low, ch = sh.bk.cons_high;
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
                // `_nextArc = e` (NS.cs:1397): the last arc examined, or unchanged after a full sweep that ended inside a block
                if (search_end < S || (long long)S % Bold == 0) { int e = next_arc + search_end - 1; if (e >= S) e -= S; next_arc = e; }
            }
            TICK(t_price);
            PROBE(2);
        }
        if (!have_win) { status = ST_OPTIMAL; break; }
        iterations = k;
        if (iterations > P.max_iterations) { status = ST_INFEASIBLE; break; }          // NS.cs:311-317

        if (pricer) {
            // ---- stage this pricer's share of the NEXT pivot's first block.  Arc data streams from DRAM right away; the node
            // records are gathered once every update up to pivot k-1 is visible (hop 3 of the previous pivot, off the critical
            // path) and BEFORE any owner applies update k: owners wait for GATHERED(k+1) below.  Update k is replayed at pricing.
            int s_lo, s_hi;
            const bool fits = share(B, s_lo, s_hi);
            if (fits && !(spec_cursor == next_arc && spec_B == B)) [[unlikely]] {     // not what was predicted: drain and start over
                cp_async_wait_all();
                stage_static_begin(next_arc, s_lo, s_hi, stv);
            }
            spec_cursor = -1;
            PROBE(3);
            if (!wait_done(k - 1)) { status = ST_ERR_BARRIER_TIMEOUT; break; }
            PROBE(6);
            if (fits) { stage_static_finish(s_lo, s_hi, stv); PROBE(0); stage_gather(s_lo, s_hi); pf_next = next_arc; pf_B = B; pf_upto = k - 1; } else pf_next = -1;
            __syncthreads();
            if (tid == 0) st_vol_u32(P.done + (size_t)(G + cta) * 32, (unsigned)(k + 1));      // GATHERED(k+1)
            PROBE(7);
        } else PROBE(9);

        const PWin ent = win_rec >= 0 ? unpack_pwin(sh.rec[win_rec]) : sh.win;
        const int in_arc = ent.arc, a_src = ent.src, a_tgt = ent.tgt, a_cost = ent.cost, a_state = ent.state;
        const long long upper_in = ent.upper;
        const bool lower_state = a_state == STATE_LOWER;
        const int first = lower_state ? a_src : a_tgt;                                  // NS.cs:948-957
        const int inF = lower_state ? ent.in_s : ent.in_t, inS = lower_state ? ent.in_t : ent.in_s;
        const long long piF = lower_state ? ent.pi_s : ent.pi_t, piS = lower_state ? ent.pi_t : ent.pi_s;
        const int dpF = lower_state ? ent.dp_s : ent.dp_t, dpS = lower_state ? ent.dp_t : ent.dp_s;

        // ================================================================ owners: cycle discovery over the slice, post CYC(k)
        // A node is on the pivot cycle iff exactly one end of the entering arc lies in its subtree (FindJoinNode + both walks
        // of FindLeavingArc, NS.cs:925-1010, as one interval test per node).
        auto make_cand = [&](int j, int in_u, int sz_u, bool hasF) -> Cand {
            const int pd = pd_s[j];
            const F fl = fl_s[j], up = up_s[j];
            const bool dir_up = pd & 1;
            // first walk: residual capacity when pred_dir == DOWN, else the flow; second walk mirrored (NS.cs:968, :986)
            const bool increase = hasF ? !dir_up : dir_up;
            Cand cd; cd.d = increase ? FT::residual(up, fl) : (long long)fl; cd.in = in_u; cd.sz = sz_u; cd.pd = pd; cd.dp = dp_get(j); cd.j = j;
            cd.zero = (((!increase) || up == 0) ? 1 : 0) | (hasF ? 2 : 0);
            return cd;
        };
        int nc = 0;
        if (!pricer) {
            if (tid == 0) sh.ncand = 0;
            __syncthreads();
            // four nodes per 128-bit shared-memory load; the slice is padded to a multiple of 8 with nodes that match nothing
            const int nquad = cntn > 0 ? (cntn + 3) >> 2 : 0;
            for (int q4 = tid; q4 < nquad; q4 += kTT) {
                const int4 vi = reinterpret_cast<const int4*>(in_s)[q4], vz = reinterpret_cast<const int4*>(sz_s)[q4];
                const int xi[4] = {vi.x, vi.y, vi.z, vi.w}, xz[4] = {vz.x, vz.y, vz.z, vz.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool hasF = (unsigned)(inF - xi[e]) < (unsigned)xz[e];
                    const bool hasS = (unsigned)(inS - xi[e]) < (unsigned)xz[e];
                    if (hasF != hasS) {
                        const int slot = atomicAdd(&sh.ncand, 1);
                        if (slot < kCandCap) sh.cl[slot] = make_cand(q4 * 4 + e, xi[e], xz[e], hasF);
                    }
                }
            }
            __syncthreads();
            nc = sh.ncand;
            PROBE(10);
            // the record goes out in kRepCyc copies (lane l writes word l % 5 of copy l / 5)
            int4* const rec = P.cyc + (((size_t)par * kRepCyc + lane / 5) * G + cta) * kMailWords;
            if (nc == 0) {
                if (tid < 5 * kRepCyc) st_vol4(rec + lane % 5, make_int4(0, sh.wide_req ? 16 : 0, 0, seq));
            } else {
                Cand m1 = cand_none(), m2 = cand_none();
                if (nc <= kCandCap) {
                    if (warp == 0) {
                        // strict '<' walking up from `first`: deepest minimum; '<=' walking up from `second`: shallowest minimum
                        Cand c = cand_none();
                        if (lane < nc) c = sh.cl[lane];
                        const int w1 = warp_argmin(lane < nc && (c.zero & 2), c.d, -c.in);
                        const int w2 = warp_argmin(lane < nc && !(c.zero & 2), c.d, c.in);
                        if (w1 >= 0) m1 = sh.cl[w1];
                        if (w2 >= 0) m2 = sh.cl[w2];
                    }
                } else [[unlikely]] {
                    // rare: many cycle nodes in one slice - recompute the candidates and reduce over the whole CTA
                    Cand b1 = cand_none(), b2 = cand_none();
                    for (int j = tid; j < cntn; j += kTT) {
                        const int in_u = in_s[j], sz_u = sz_s[j];
                        const bool hasF = (unsigned)(inF - in_u) < (unsigned)sz_u;
                        const bool hasS = (unsigned)(inS - in_u) < (unsigned)sz_u;
                        if (hasF != hasS) {
                            const Cand cd = make_cand(j, in_u, sz_u, hasF);
                            if (hasF) { if (b1.pd < 0 || cd.d < b1.d || (cd.d == b1.d && cd.in > b1.in)) b1 = cd; }
                            else      { if (b2.pd < 0 || cd.d < b2.d || (cd.d == b2.d && cd.in < b2.in)) b2 = cd; }
                        }
                    }
                    const int l1 = warp_argmin(b1.pd >= 0, b1.d, -b1.in), l2 = warp_argmin(b2.pd >= 0, b2.d, b2.in);
                    if (lane == 0) { sh.wc[0][warp].pd = -1; sh.wc[1][warp].pd = -1; }
                    __syncwarp();
                    if (l1 >= 0 && lane == l1) sh.wc[0][warp] = b1;
                    if (l2 >= 0 && lane == l2) sh.wc[1][warp] = b2;
                    __syncthreads();
                    if (warp == 0) {
                        const Cand c1 = sh.wc[0][lane & (kTW - 1)], c2 = sh.wc[1][lane & (kTW - 1)];
                        const int w1 = warp_argmin(lane < kTW && c1.pd >= 0, c1.d, -c1.in);
                        const int w2 = warp_argmin(lane < kTW && c2.pd >= 0, c2.d, c2.in);
                        if (w1 >= 0) m1 = sh.wc[0][w1];
                        if (w2 >= 0) m2 = sh.wc[1][w2];
                    }
                }
                if (warp == 0 && lane < 5 * kRepCyc) {
                    const int wd = lane % 5;
                    int4 w;
                    if (wd == 0) w = make_int4(nc, (m1.zero & 1) | ((m2.zero & 1) << 1) | (m1.pd >= 0 ? 4 : 0) | (m2.pd >= 0 ? 8 : 0) | (sh.wide_req ? 16 : 0), 0, seq);
                    else if (wd == 1) w = make_int4(lo32(m1.d), hi32(m1.d), m1.in, seq);
                    else if (wd == 2) w = make_int4(m1.sz, m1.pd, m1.dp, seq);
                    else if (wd == 3) w = make_int4(lo32(m2.d), hi32(m2.d), m2.in, seq);
                    else w = make_int4(m2.sz, m2.pd, m2.dp, seq);
                    st_vol4(rec + wd, w);
                }
            }
            PROBE(11);
        }

        // ================================================================ all: hop 2, gather CYC(k) and decide
        {
            if (tid == 0) sh.cnt = 0;
            __syncthreads();
            const int nw = (nown + 31) >> 5;                                            // warps that poll
            if (!pricer && warp == kTW - 1 && lane < NP) {
                // owners: update k may touch the mirror only after every pricer has gathered the next block's node records
                // (GATHERED(k+1)); polled here, next to the CYC records, so that it costs nothing when it is already there
                const unsigned want = (unsigned)(k + 1);
                const unsigned* p = P.done + (size_t)(G + lane) * 32;
                unsigned spins = 0; long long t0 = 0;
                while ((int)(ld_vol_u32(p) - want) < 0) if (spin_check(spins, t0, P)) { sh.abort = 1; break; }
            }
            int wreq = 0;
            if (warp < nw) {
                Cand b1 = cand_none(), b2 = cand_none();
                int c = 0;
                if (tid < nown) {
                    int4 w[5];
                    if (!poll_rec<5>(P.cyc + (((size_t)par * kRepCyc + cta % kRepCyc) * G + NP + tid) * kMailWords, seq, w, P)) sh.abort = 1;
                    else {
                        c = w[0].x; wreq = w[0].y & 16;
                        if (w[0].y & 4) { b1.d = mk64(w[1].x, w[1].y); b1.in = w[1].z; b1.sz = w[2].x; b1.pd = w[2].y; b1.dp = w[2].z; b1.zero = w[0].y & 1; }
                        if (w[0].y & 8) { b2.d = mk64(w[3].x, w[3].y); b2.in = w[3].z; b2.sz = w[4].x; b2.pd = w[4].y; b2.dp = w[4].z; b2.zero = (w[0].y >> 1) & 1; }
                    }
                }
                c = __reduce_add_sync(0xffffffffu, c);
                if (lane == 0 && c) atomicAdd(&sh.cnt, c);
                const int l1 = warp_argmin(b1.pd >= 0, b1.d, -b1.in), l2 = warp_argmin(b2.pd >= 0, b2.d, b2.in);
                if (lane == 0) { sh.wc[0][warp].pd = -1; sh.wc[1][warp].pd = -1; }
                __syncwarp();
                if (l1 >= 0 && lane == l1) sh.wc[0][warp] = b1;
                if (l2 >= 0 && lane == l2) sh.wc[1][warp] = b2;
            }
            // narrow mode: some owner saw a flow leave the int32 range during update k-1.  The flag arrives with the CYC records, so
            // every CTA of the team leaves the loop on this same pivot and the host re-runs wide at once (the narrow
            // solve used to carry on with truncated flows until it ended by itself).
            const int any_wide = __syncthreads_or(wreq);
            if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
            if (any_wide) [[unlikely]] { status = ST_ERR_NEEDS_WIDE; break; }
            Cand w1 = cand_none(), w2 = cand_none();
            for (int w = 0; w < nw; ++w) {
                const Cand t1 = sh.wc[0][w], t2 = sh.wc[1][w];
                if (t1.pd >= 0 && (w1.pd < 0 || t1.d < w1.d || (t1.d == w1.d && t1.in > w1.in))) w1 = t1;
                if (t2.pd >= 0 && (w2.pd < 0 || t2.d < w2.d || (t2.d == w2.d && t2.in < w2.in))) w2 = t2;
            }
            const bool has1 = w1.pd >= 0, has2 = w2.pd >= 0;
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
            const int b = in_side1 ? inS : inF;                                         // in[v_in]
            const int dp_uin = in_side1 ? dpF : dpS, dp_vin = in_side1 ? dpS : dpF;
            const bool src_side1 = lower_state;                                         // is `first` the source of the entering arc?
            // stem = cycle nodes on u_in's side from u_in (index 0, deepest) up to u_out (index ns-1); depths give the index
            const int ns = change ? dp_uin - out.dp + 1 : 1;
            const bool longstem = ns > kTeamStemCap;
            int4* const stem_g = P.stemseg + (size_t)par * (n + 1) * 2;                 // entry of stem index k at slot t = ns-1-k

            // new flow on the pred arc of a cycle node (ChangeFlow, NS.cs:1020-1029)
            auto new_flow = [&](long long fl, int pd, bool hasF) -> long long {
                if (delta <= 0) return fl;
                const long long dv = (pd & 1) ? val : -val;                             // pred_dir * val
                return (hasF == src_side1) ? fl - dv : fl + dv;
            };

            // ---- hop 2b (26 % of pivots): the stem is longer than one node; its owners publish the entries, index = depth
            if (ns > 1) {
                if (tid == 0) sh.bk.stem_x++;
                if (!pricer) {
                    auto publish = [&](int j, int in_u, int sz_u, int pd, int dp, bool hasF) {
                        const long long fl = new_flow((long long)fl_s[j], pd, hasF);
                        const long long upl = FT::cap_out(up_s[j]);                      // capacity travels too when it fits 31 bits (-1: fetch)
                        int4* e = stem_g + (size_t)(dp - out.dp) * 2;
                        st_vol4(e, make_int4(in_u, sz_u, pd, seq));
                        st_vol4(e + 1, make_int4(lo32(fl), hi32(fl), upl == LLONG_MAX / 2 ? INT_MAX : (upl < (long long)INT_MAX ? (int)upl : -1), seq));
                    };
                    if (nc <= kCandCap) {
                        if (tid < nc) {
                            const Cand c = sh.cl[tid];
                            const bool hasF = (c.zero & 2) != 0;
                            if (hasF == in_side1 && c.in >= a) publish(c.j, c.in, c.sz, c.pd, c.dp, hasF);
                        }
                    } else {
                        for (int j = tid; j < cntn; j += kTT) {
                            const int in_u = in_s[j], sz_u = sz_s[j];
                            const bool hasF = (unsigned)(inF - in_u) < (unsigned)sz_u;
                            const bool hasS = (unsigned)(inS - in_u) < (unsigned)sz_u;
                            if (hasF != hasS && hasF == in_side1 && in_u >= a) publish(j, in_u, sz_u, pd_s[j], dp_get(j), hasF);
                        }
                    }
                }
                if (!longstem) {
                    for (int q = tid; q < ns; q += kTT) {
                        int4 w[2];
                        if (!poll_rec<2>(stem_g + (size_t)q * 2, seq, w, P)) sh.abort = 1;
                        const int kx = ns - 1 - q;
                        st_in[kx] = w[0].x; st_z[kx] = w[0].y; st_pd[kx] = w[0].z; st_fl[kx] = mk64(w[1].x, w[1].y); st_up[kx] = w[1].z;
                    }
                    __syncthreads();
                    if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
                }
                TICK(t_stem);
            }
            if (change && tid == 0) { if (ns > sh.bk.max_stem) sh.bk.max_stem = ns; sh.bk.moved_nodes += s; }

            // ================================================================ updates
            const bool dir_new_up = u_in == a_src;                                      // NS.cs:1143
            const long long piU = in_side1 ? piF : piS, piV = in_side1 ? piS : piF;
            Pending U;                                                                  // this pivot's update in closed form
            U.valid = 1; U.change = change ? 1 : 0; U.a = a; U.s = s; U.b = b; U.ns = ns; U.longstem = longstem ? 1 : 0;
            U.dshift = dp_vin + 1 - dp_uin; U.par = par; U.seq = seq;
            U.sigma = piV - piU - (dir_new_up ? (long long)a_cost : -(long long)a_cost);              // NS.cs:1187-1188
            if (pricer) {
                // arc states (ChangeFlow, NS.cs:1031-1039): only the pricing scans read them; every pricer keeps its own view
                patch2_arc0 = patch_arc0; patch2_st0 = patch_st0; patch2_arc1 = patch_arc1; patch2_st1 = patch_st1;
                if (change) { patch_arc0 = in_arc; patch_st0 = STATE_TREE; patch_arc1 = out.pd >> 1; patch_st1 = (out.zero & 1) ? STATE_LOWER : STATE_UPPER; }
                else { patch_arc0 = in_arc; patch_st0 = -a_state; patch_arc1 = -1; }
                if (tid == 0) {
                    P.state[patch_arc0] = patch_st0;
                    if (patch_arc1 >= 0) P.state[patch_arc1] = patch_st1;
                }
                Uprev = U;                                                              // replayed by the next pricing (see above)
            } else {
                if (!change && delta > 0 && tid == 0 && first >= lo && first < lo + cntn)
                    P.flow[in_arc] = (lower_state ? 0 : upper_in) + val;                // NS.cs:1018: stays a non-tree arc, at the other bound
                int bad = 0;
                // ---- cycle nodes: ChangeFlow (NS.cs:1012-1040) and the pred / succ_num part of UpdateTreeStructure (:1042-1183)
                auto update_cycle_node = [&](int j, int x, int sz_u, int pd, int dp, bool hasF) {
                    if (delta > 0) {
                        const long long fl = new_flow((long long)fl_s[j], pd, hasF);
                        bad |= !FT::fits(fl);
                        fl_s[j] = (F)fl;
                    }
                    if (!change) return;
                    if (hasF != in_side1) { sz_s[j] = sz_u + s; return; }               // v_in .. join (NS.cs:1174-1177)
                    if (x < a) { sz_s[j] = sz_u - s; return; }                          // v_out .. join (NS.cs:1179-1182)
                    // stem node kx (NS.cs:1095-1146): takes over the pred arc of the stem node below it, reversed
                    if (x == a) P.flow[pd >> 1] = (out.zero & 1) ? 0 : FT::cap_out(up_s[j]);   // u_out: its pred arc leaves the tree at a bound
                    const int kx = dp_uin - dp;
                    if (kx == 0) {
                        const long long nf = (lower_state ? 0 : upper_in) + val;
                        bad |= !FT::fits(nf);
                        pd_s[j] = in_arc * 2 + (dir_new_up ? 1 : 0); sz_s[j] = s;
                        fl_s[j] = (F)nf; up_s[j] = FT::cap_in(upper_in);
                    } else {
                        int p_z, p_pd, p_up; long long p_fl;
                        if (!longstem) { p_z = st_z[kx - 1]; p_pd = st_pd[kx - 1]; p_up = st_up[kx - 1]; p_fl = st_fl[kx - 1]; }
                        else {
                            int4 w[2];
                            if (!poll_rec<2>(stem_g + (size_t)(ns - kx) * 2, seq, w, P)) sh.abort = 1;
                            p_z = w[0].y; p_pd = w[0].z; p_fl = mk64(w[1].x, w[1].y); p_up = w[1].z;
                        }
                        const int npd = p_pd ^ 1;
                        bad |= !FT::fits(p_fl);
                        pd_s[j] = npd; sz_s[j] = s - p_z;
                        fl_s[j] = (F)p_fl;
                        up_s[j] = FT::cap_in(p_up == INT_MAX ? LLONG_MAX / 2 : (p_up >= 0 ? (long long)p_up : __ldg(P.upper + (npd >> 1))));
                    }
                };
                if (nc <= kCandCap) {
                    if (tid < nc) { const Cand c = sh.cl[tid]; update_cycle_node(c.j, c.in, c.sz, c.pd, c.dp, (c.zero & 2) != 0); }
                } else {
                    for (int j = tid; j < cntn; j += kTT) {
                        const int x = in_s[j], sz_u = sz_s[j];
                        const bool hasF = (unsigned)(inF - x) < (unsigned)sz_u;
                        const bool hasS = (unsigned)(inS - x) < (unsigned)sz_u;
                        if (hasF != hasS) update_cycle_node(j, x, sz_u, pd_s[j], dp_get(j), hasF);
                    }
                    __syncthreads();                                                    // the relabel pass below rewrites in_s
                }
                // ---- every node: re-label in[] in closed form; re-hung subtree: new depth and pi += sigma (NS.cs:1185-1209)
                PROBE(15);
                if (change) {
                    const int sh_lo = b < a ? b + 1 : a + s, sh_len = b < a ? a - b - 1 : b - a - s + 1, sh_by = b < a ? s : -s;
                    // consecutive lanes <-> consecutive nodes (the mirror stores coalesce), four independent nodes per thread in flight
                    for (int j0 = tid; j0 < cntn; j0 += kRelUnroll * kTT) {
                        int xv[kRelUnroll];
#pragma unroll
                        for (int e = 0; e < kRelUnroll; ++e) { const int j = j0 + e * kTT; xv[e] = j < cntn ? in_s[j] : 0; }
#pragma unroll
                        for (int e = 0; e < kRelUnroll; ++e) {
                            const int x = xv[e], j = j0 + e * kTT;
                            if ((unsigned)(x - sh_lo) < (unsigned)sh_len) {                // between the old and the new place: shift
                                in_s[j] = x + sh_by; P.in_g[lo + j] = x + sh_by;
                            } else if ((unsigned)(x - a) < (unsigned)s) {                  // re-hung subtree
                                int nx, nd;
                                relabel(U, x, dp_get(j), nx, nd);
                                in_s[j] = nx; if (dp_res) dp_s[j] = nd;
                                atomicAdd(reinterpret_cast<unsigned long long*>(&P.node[lo + j].pi), (unsigned long long)U.sigma);
                                P.in_g[lo + j] = nx; P.node[lo + j].dp = nd;
                            }
                        }
                    }
                }
                if (bad) { P.ctl->needs_wide = 1; sh.wide_req = 1; }
                // hop 3: everything this CTA wrote for pivot k is visible before DONE(k)
                PROBE(13);
                __syncthreads();
                if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
                if (tid == kTT - 32) { __threadfence(); st_vol_u32(P.done + (size_t)cta * 32, (unsigned)k); }   // not a thread that polls ENTER next
                PROBE(14);
            }
            TICK(t_update);
            PROBE(5);
        }
        if (P.stop_after > 0 && iterations >= P.stop_after) { status = ST_STOPPED_EARLY; break; }
    }
#undef TICK
#undef PROBE
    if (tid == 0 && (cta == 0 || cta == NP)) for (int i = cta == 0 ? 0 : 8; i < (cta == 0 ? 8 : 16); ++i) P.ctl->clk[i] = sh.bk.pr[i];

    // =================================================================== epilogue
    const bool clean = status != ST_ERR_BARRIER_TIMEOUT;
    if (clean) {
        // flows of the tree arcs go back to flow[]; then one conventional grid barrier (counter + fences)
        for (int j = tid; j < cntn; j += kTT) { const int pd = pd_s[j]; if (pd >= 0) P.flow[pd >> 1] = (long long)fl_s[j]; }
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
    if (cta == 0 && tid == 0) {
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

extern "C" cudaError_t mcfk_device_props(int device, cudaDeviceProp* out);      // mcf_kernels.cu: cached cudaGetDeviceProperties

namespace {
constexpr size_t kStemBytes = (size_t)mcf::kTeamStemCap * (8 + 4 * 4);
constexpr size_t kPricerBytes = (size_t)mcf::kPf * mcf::kTT * (3 * 8 + 8 * 4);
inline const void* team_fn(int wide) { return wide ? (const void*)mcf::ns_team_kernel<long long> : (const void*)mcf::ns_team_kernel<int>; }
}  // namespace

static inline int node_bytes(int wide, int spill) { return spill ? mcf::kNodeSmemSpill : wide ? mcf::kNodeSmemWide : mcf::kNodeSmemNarrow; }

extern "C" size_t mcfk_team_smem_bytes(int slice, int wide, int spill)
{
    const size_t owner = (size_t)slice * node_bytes(wide, spill);
    return kStemBytes + (owner > kPricerBytes ? owner : kPricerBytes) + 16;
}

// largest slice (nodes per owner CTA) that fits the opt-in shared memory of the device next to the kernel's static part
extern "C" int mcfk_team_max_slice(int device, int wide, int spill)
{
    cudaDeviceProp prop;
    if (mcfk_device_props(device, &prop) != cudaSuccess) return -1;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, team_fn(wide)) != cudaSuccess) return -2;
    const long long avail = (long long)prop.sharedMemPerBlockOptin - (long long)fa.sharedSizeBytes - (long long)kStemBytes - 64;
    const long long s = avail / node_bytes(wide, spill);
    return (int)(s & ~7LL);
}

// the dynamic shared-memory ceiling of the kernel is always raised to the device's opt-in maximum: several host threads may
// prepare launches with different slice sizes at the same time (mcf_solve_batch_concurrent)
static cudaError_t raise_smem_limit(int device, int wide)
{
    cudaDeviceProp prop;
    cudaError_t e = mcfk_device_props(device, &prop);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, team_fn(wide));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(team_fn(wide), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(prop.sharedMemPerBlockOptin - fa.sharedSizeBytes));
}

extern "C" int mcfk_team_max_ctas(int device, int slice, int wide, int spill)
{
    cudaDeviceProp prop;
    if (mcfk_device_props(device, &prop) != cudaSuccess) return -1;
    const size_t smem = mcfk_team_smem_bytes(slice, wide, spill);
    if (raise_smem_limit(device, wide) != cudaSuccess) return -2;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, team_fn(wide), mcf::kTT, smem) != cudaSuccess) return -3;
    return per_sm * prop.multiProcessorCount;
}

extern "C" void mcfk_team_replicas(int* ent, int* cyc) { *ent = mcf::kRepEnt; *cyc = mcf::kRepCyc; }

extern "C" int mcfk_launch_team(const mcf::TeamParams* p, cudaStream_t stream)
{
    const size_t smem = mcfk_team_smem_bytes(p->slice, p->wide, p->spill);
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e == cudaSuccess) e = raise_smem_limit(device, p->wide);
    if (e != cudaSuccess) return (int)e;
    void* args[] = {(void*)p};
    e = cudaLaunchCooperativeKernel(team_fn(p->wide), dim3(p->team), dim3(mcf::kTT), args, smem, stream);
    return (int)e;
}
