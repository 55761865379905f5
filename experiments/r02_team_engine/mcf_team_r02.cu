// mcf_team.cu - the "team" pivot engine of libmcfgpu (sm_100a): Block Search network simplex as one persistent
// cooperative kernel in which the basis tree never leaves the chip.
//
// Why: a pivot of NetworkSimplex.Solve() (NS.cs:282-341) is a chain of pointer walks over parent/pred/thread/
// succ_num/last_succ (NS.cs:925-1209).  On B200 a dependent L2 load costs ~150 ns and one message between two SMs
// through L2 ~0.4-0.55 us (profiles/r02_micro_cluster.txt), so the walks are replaced by flat passes over an interval
// labelling (in[u] = DFS index, sz[u] = subtree size; see mcf_device.cuh) and the whole basis - labels, pred arcs, depth
// and the flow and capacity of every tree arc - is kept in SHARED MEMORY, sliced by node id over the "owner" CTAs of the
// team.  CTA 0 is the pricer.  What crosses between CTAs per pivot is two messages on the critical path and one off it,
// all made of 16-byte words that carry their own sequence number in the same 128-bit relaxed.gpu store (no fence, no
// barrier):
//
//   ENTER(k)   pricer -> all     the entering arc of pivot k with both ends' (pi, in) + the request "stage arcs [c, c+B)"
//   CYC(k)     every owner -> all its best leaving-arc candidate per side of the cycle, counts, depth of the arc's ends
//  (STEM(k)    owners -> all     only when the re-hung stem is longer than one node: the stem entries, indexed by depth)
//   STAGE      owners -> pricer  off the critical path: {pi, in} of both ends of every arc of the requested range, each
//                                written by the CTA that owns the node, from its own shared memory / its own part of pi[]
//
// The pricer runs BlockSearchPivot.FindEnteringArc (NS.cs:1339-1441).  The block of the NEXT pivot is known exactly when
// ENTER(k) is posted (its cursor is the last arc examined, NS.cs:1397); its arc data streams from DRAM with cp.async and
// its node records are served by the owners while CYC(k) is in flight - as of the basis BEFORE update k.  Update k is then
// replayed on the staged records in closed form at pricing time (pi += sigma inside the re-hung interval, labels through
// the same relabel formula the owners use), so pricing never waits for any owner's writes to become visible and no node
// array lives in global memory at all except pi[], of which every entry has exactly one reader/writer CTA.
// Owners run FindJoinNode + FindLeavingArc as an interval test over their slice, every CTA reduces the candidates
// redundantly to the same decision (strict '<' on the first walk, '<=' on the second, NS.cs:958-998), owners apply
// ChangeFlow / UpdateTreeStructure / UpdatePotentials (NS.cs:1012-1209) to the nodes they own.  Flows of tree arcs live
// with the node below the arc; flow[] in global memory is written when an arc leaves the tree and at the end.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "mcf_device.cuh"

namespace mcf {

namespace {

constexpr int kTT = 512;                                // threads per CTA: latency-bound code, up to 128 registers each
constexpr int kTW = kTT / 32;
constexpr int kStagePos = kStageMax + 16;                // staging positions: the block plus alignment gaps
constexpr int kPos = (kStagePos + kTT - 1) / kTT;       // positions per pricer thread
constexpr int kRepEnt = 4;                              // replicas of the ENTER record: a reader polls replica (cta % kRepEnt)
constexpr int kRepCyc = 4;                              // replicas of every CYC record
constexpr int kRelUnroll = 4;                           // nodes per thread in flight in the relabel pass
constexpr int kMaxPricers = 6;                          // pricing CTAs of a team (4 words each + the request word are polled by one warp)
constexpr int kCandCap = 32;                            // cycle nodes of one slice handled by the single-warp path

// mailbox words: one 128-bit relaxed.gpu access each (single-copy atomic, PTX ISA 8.3+), polled until the sequence number matches
__device__ __forceinline__ int4 ld_mail(const int4* p)
{
    int4 v;
    asm volatile("{\n .reg .b128 t;\n ld.relaxed.gpu.global.b128 t, [%4];\n mov.b128 {%0,%1,%2,%3}, t;\n}" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_mail(int4* p, int4 v)
{
    asm volatile("{\n .reg .b128 t;\n mov.b128 t, {%1,%2,%3,%4};\n st.relaxed.gpu.global.b128 [%0], t;\n}" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
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
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
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
    long long rc;                   // reduced cost state * (cost + pi_s - pi_t), negative when valid
    int off;                        // scan offset from next_arc (< 0: none)
    int arc, src, tgt, state, in_s, in_t;
    int blk;                        // block of the scan the candidate lies in (later rounds of a search)
    long long rcb, upper;           // cost + pi_s - pi_t; capacity
};

struct Book {                       // statistics and timers: touched by thread 0 only, kept out of the register file
    long long arcs_checked, rounds_total, degenerate, cycle_nodes, moved_nodes, max_cycle, max_stem, stem_x;
    unsigned long long t_price, t_cycle, t_update, t_wdone, t_stem, t_mark, t_begin, c_begin, pr_mark;
    unsigned long long pr[16];
    int cons_low, cons_high;
};

struct Pending {                    // one pivot's update in closed form: what the pricer replays on staged node records
    int valid, change;
    int a, s, b;                    // re-hung subtree = old interval [a, a+s); b = in[v_in]
    int ns, longstem, dshift, par, seq;
    long long sigma;
};

struct Ent {                        // the entering arc of a pivot as every CTA knows it
    int arc, src, tgt, state, in_s, in_t;
    long long upper, rcb;           // capacity; cost + pi_s - pi_t (all that UpdatePotentials needs of the potentials, NS.cs:1187-1188)
};

struct Dec {                        // the decision of a pivot (FindLeavingArc, NS.cs:943-1010), identical in every CTA
    bool change, in_side1, longstem, dir_new_up;
    int a, s, dp_uin, ns, inF, inS, first;
    long long delta, val;
    Cand out;
    // new flow on the pred arc of a cycle node (ChangeFlow, NS.cs:1020-1029)
    __device__ __forceinline__ long long new_flow(long long fl, int pd, bool hasF, bool src_side1) const
    {
        if (delta <= 0) return fl;
        const long long dv = (pd & 1) ? val : -val;                                 // pred_dir * val
        return (hasF == src_side1) ? fl - dv : fl + dv;
    }
};

struct TeamShared {
    Book bk;
    Cand cl[kCandCap];              // cycle-node candidates of this slice (owner scan)
    Cand wc[2][kTW];                // per-warp winners (CYC gather, owner slow path)
    longlong2 pk[kTW];              // per-warp pricing winners: {reduced cost, position}
    Ent win;                        // entering arc of this pivot (pricer)
    int patch[4];                   // pricer: the two arc-state changes of this pivot
    int4 ent[5];                    // the winning ENTER record (words 0-3) and the staging request (word 4) as received
    int4 crec[kTW][4];              // pricer: the candidate record every warp prepares
    int ncand, abort, cnt, mode, dpF, dpS, ovf;
};

// time-out / abort check for spin loops; true = give up
__device__ __forceinline__ bool spin_check(unsigned& spins, long long& t0, const TeamParams& P, int site = 0)
{
#ifdef MCF_SPIN_SLEEP
    __nanosleep(MCF_SPIN_SLEEP);                        // back off: fewer polling requests in flight at L2
#endif
    if ((++spins & 255u) != 0) return false;
    if (t0 == 0) { t0 = clock64(); return false; }
    if (*(volatile int*)&P.ctl->abort) return true;
    if ((unsigned long long)(clock64() - t0) > P.timeout_cycles) {
        if (atomicCAS(&P.ctl->abort, 0, 1) == 0) P.ctl->pad0 = site * 1000 + (int)blockIdx.x;      // who gave up first, and where (for the error message)
        return true;
    }
    return false;
}

// poll one self-validating word until its sequence number matches; false = abandoned (abort flag or time-out)
__device__ __forceinline__ bool poll_word(const int4* p, int seq, int4& out, const TeamParams& P, int site = 1)
{
    unsigned spins = 0; long long t0 = 0;
    for (;;) {
        const int4 v = ld_mail(p);
        if (v.w == seq) { out = v; return true; }
        if (spin_check(spins, t0, P, site)) { out = v; return false; }
    }
}

// poll NW words of one record, all loads in flight together
template <int NW>
__device__ __forceinline__ bool poll_rec(const int4* rec, int seq, int4 (&w)[NW], const TeamParams& P, int site = 2)
{
    unsigned spins = 0; long long t0 = 0;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = ld_mail(rec + i);
#pragma unroll
        for (int i = 0; i < NW; ++i) ok = ok && w[i].w == seq;
        if (ok) return true;
        if (spin_check(spins, t0, P, site)) return false;
    }
}

__device__ __forceinline__ PWin pwin_none()
{
    PWin w; w.rc = 0; w.off = -1; w.arc = -1; w.src = w.tgt = w.state = w.in_s = w.in_t = w.blk = 0; w.rcb = w.upper = 0;
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
    const int Gp = (G + 7) & ~7;                         // CYC word arrays are padded to whole 128-byte lines
    const int n = P.n, S = P.S;
    const bool pricer = cta < NP;
    const int own = cta - NP;
    const int lo = pricer ? 0 : own * P.slice;
    const int cntn = pricer ? 0 : max(0, min(n + 1, lo + P.slice) - lo);

    // dynamic shared memory.  Everybody: stem staging.  Owners: the resident slice.  Pricer: the staged block.
    long long* const st_fl = reinterpret_cast<long long*>(dyn_smem);            // [kTeamStemCap] flow on stem k's old pred arc (after augmentation)
    int* const st_in = reinterpret_cast<int*>(st_fl + kTeamStemCap);            // sorted: stem 0 = u_in (deepest) .. u_out
    int* const st_z = st_in + kTeamStemCap;
    int* const st_pd = st_z + kTeamStemCap;
    int* const st_up = st_pd + kTeamStemCap;                                    // capacity of that arc (INT_MAX: infinite, -1: fetch)
    unsigned char* const body = reinterpret_cast<unsigned char*>(st_up + kTeamStemCap);
    // owners
    F* const fl_s = reinterpret_cast<F*>(body);                                 // flow on the pred arc of node j
    F* const up_s = fl_s + P.slice;                                             // capacity of the pred arc
    int* const in_s = reinterpret_cast<int*>(up_s + P.slice);
    int* const sz_s = in_s + P.slice;
    int* const pd_s = sz_s + P.slice;
    int* const dp_s = pd_s + P.slice;                                           // depth in the basis tree
    // pricer: arc data and both ends' node records of the staged block [pf_next, pf_next + pf_B)
    long long* const pf_up = reinterpret_cast<long long*>(dyn_smem);           // capacity (the pricer stages no stems: the whole area is its)
    long long* const pf_rcb = pf_up + kStagePos;                                // [2] cost + pi_s - pi_t as of the basis the records were served from
    int2* const pf_lab = reinterpret_cast<int2*>(pf_rcb + 2 * kStagePos);       // [2] {in[src], in[tgt]} as of the same basis
    int* const pf_src = reinterpret_cast<int*>(pf_lab + 2 * kStagePos);
    int* const pf_tgt = pf_src + kStagePos;
    int* const pf_st = pf_tgt + kStagePos;
    int* const pf_cost = pf_st + kStagePos;

    int status = ST_NOT_SOLVED;
    {
        int bad = 0;
        for (int j = tid; j < cntn; j += kTT) {
            const int u = lo + j;
            const int pd = P.pd0[u];
            in_s[j] = P.in0[u]; dp_s[j] = P.dp0[u]; sz_s[j] = P.sz0[u]; pd_s[j] = pd;
            const long long fl = pd >= 0 ? P.flow[pd >> 1] : 0, up = pd >= 0 ? P.upper[pd >> 1] : 0;
            bad |= !FT::fits(fl);
            fl_s[j] = (F)fl; up_s[j] = FT::cap_in(up);
        }
        if (!pricer) for (int j = cntn + tid; j < P.slice; j += kTT) { in_s[j] = 0; sz_s[j] = 0; dp_s[j] = 0; pd_s[j] = -2; fl_s[j] = 0; up_s[j] = 0; }   // padding: on no cycle, never relabelled
        if (tid == 0) { sh.abort = 0; sh.ncand = 0; sh.ovf = 0; Book z = {}; sh.bk = z; }
        if (__syncthreads_or(bad)) { if (tid == 0) sh.ovf = 1; }                // reported with CYC(1): the host re-runs wide
        __syncthreads();
    }
    if (tid == kTT - 32) { sh.bk.t_begin = gtimer(); sh.bk.c_begin = sh.bk.t_mark = sh.bk.pr_mark = (unsigned long long)clock64(); }

    long long iterations = 0;
    // statistics and phase timers are kept by ONE thread of the last warp of CTA 0 (pricer: slots 0-7) and CTA 1 (first owner: 8-15):
    // it polls nothing and posts nothing, so that reading the clock never sits in front of a message
    const bool probe_thr = tid == kTT - 32 && (cta == 0 || cta == NP);
    int cons_low = 0, cons_high = 0;                     // pricer: adaptive block size counters (NS.cs:1399-1438)
#define PROBE(i) do { if (probe_thr && ((i) < 8) == (cta == 0)) { const unsigned long long t__ = (unsigned long long)clock64(); sh.bk.pr[i] += t__ - sh.bk.pr_mark; sh.bk.pr_mark = t__; } } while (0)
#define TICK(acc) do { if (probe_thr && cta == 0) { const unsigned long long t__ = (unsigned long long)clock64(); sh.bk.acc += t__ - sh.bk.t_mark; sh.bk.t_mark = t__; } } while (0)

    // closed-form re-labelling of one node by the update described in `U` (UpdateTreeStructure seen through in[] / depth):
    // nodes of the re-hung subtree [a, a+s) get their new place under v_in, nodes between the old and new place shift by s
    auto relabel = [&](const Pending& U, int x, int dp, int& nx, int& ndp) -> bool {
        const int4* const stem_g = P.stemseg + (size_t)U.par * (n + 1) * 2;
        auto stem_io = [&](int kx, int& o_in, int& o_z) {
            if (!U.longstem) { o_in = st_in[kx]; o_z = st_z[kx]; }
            else { int4 w; if (!poll_word(stem_g + (size_t)(U.ns - 1 - kx) * 2, U.seq, w, P, 3)) sh.abort = 1; o_in = w.x; o_z = w.y; }
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

    // ---------------------------------------------------------------------------------------------- all CTAs: gather CYC(k), decide
    // Every CTA reads every owner's record and reduces them to the same decision (leaving arc, delta, the re-hung interval),
    // then stages the stem when it is longer than one node.  Returns 0 or the status that ends the solve.
    auto gather_decide = [&]<bool kPricer>(int seq, int par, const Ent& E, int nc, Dec& D, Pending& U) -> int {
        const bool lower_state = E.state == STATE_LOWER;
        const int first = lower_state ? E.src : E.tgt, second = lower_state ? E.tgt : E.src;      // NS.cs:948-957
        const int inF = lower_state ? E.in_s : E.in_t, inS = lower_state ? E.in_t : E.in_s;
        if (tid == 0) { sh.cnt = 0; sh.ncand = 0; }
        __syncthreads();
        const int nw = (nown + 31) >> 5;                                            // warps that poll
        if (warp < nw) {
            Cand b1 = cand_none(), b2 = cand_none();
            int c = 0;
            if (tid < nown) {
                // the records are stored word-major (word w of every owner side by side): consecutive lanes poll consecutive
                // 16-byte words of the same lines.  Word 0 says whether the owner has candidates at all; only then are words 1-4 read.
                const int4* const wbase = P.cyc + ((size_t)par * kRepCyc + cta % kRepCyc) * 5 * Gp + NP + tid;
                int4 w0;
                if (!poll_word(wbase, seq, w0, P, 4)) sh.abort = 1;
                else {
                    const int f = w0.x;
                    c = f & 0xffff;
                    if (f & (3 << 18)) {
                        int4 w[4];
                        unsigned spins = 0; long long t0 = 0;
                        for (;;) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) w[i] = ld_mail(wbase + (size_t)(i + 1) * Gp);
                            if (w[0].w == seq && w[1].w == seq && w[2].w == seq && w[3].w == seq) break;
                            if (spin_check(spins, t0, P, 5)) { sh.abort = 1; break; }
                        }
                        if (f & (1 << 18)) { b1.d = mk64(w[0].x, w[0].y); b1.in = w[0].z; b1.sz = w[1].x; b1.pd = w[1].y; b1.dp = w[1].z; b1.zero = (f >> 16) & 1; }
                        if (f & (1 << 19)) { b2.d = mk64(w[2].x, w[2].y); b2.in = w[2].z; b2.sz = w[3].x; b2.pd = w[3].y; b2.dp = w[3].z; b2.zero = (f >> 17) & 1; }
                    }
                    if (f & (1 << 20)) sh.dpF = w0.y;
                    if (f & (1 << 21)) sh.dpS = w0.z;
                    if (f & (1 << 22)) sh.ovf = 2;
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
        __syncthreads();
        if (sh.abort) [[unlikely]] return ST_ERR_BARRIER_TIMEOUT;
        if (sh.ovf == 2) [[unlikely]] return ST_ERR_NEEDS_WIDE;                     // every CTA reads every record: all leave on the same pivot
        Cand w1 = cand_none(), w2 = cand_none();
        {   // second stage, redundantly in every warp: lane w looks at the winners of polling warp w
            const int wl = lane < nw ? lane : 0;
            const long long d1 = sh.wc[0][wl].d, d2 = sh.wc[1][wl].d;
            const int i1 = sh.wc[0][wl].in, i2 = sh.wc[1][wl].in;
            const int l1 = warp_argmin(lane < nw && sh.wc[0][wl].pd >= 0, d1, -i1), l2 = warp_argmin(lane < nw && sh.wc[1][wl].pd >= 0, d2, i2);
            if (l1 >= 0) w1 = sh.wc[0][l1];
            if (l2 >= 0) w2 = sh.wc[1][l2];
        }
        const bool has1 = w1.pd >= 0, has2 = w2.pd >= 0;
        const int cnt = sh.cnt;
        const int dpF = sh.dpF, dpS = sh.dpS;
        TICK(t_cycle);
        PROBE(4); PROBE(12);

        long long delta = E.upper;                                                  // NS.cs:958
        int result = 0;
        if (has1 && w1.d < delta) { delta = w1.d; result = 1; }
        if (has2 && w2.d <= delta) { delta = w2.d; result = 2; }
        const bool change = result != 0;
        if (!change && delta == 0) return ST_UNBOUNDED;                             // NS.cs:321-325
        if (probe_thr) { if (delta == 0) sh.bk.degenerate++; sh.bk.cycle_nodes += cnt; if (cnt > sh.bk.max_cycle) sh.bk.max_cycle = cnt; }
        const Cand out = result == 1 ? w1 : w2;
        const bool in_side1 = result == 1;
        const int u_in = in_side1 ? first : second;                                 // NS.cs:999-1008
        const int a = out.in, s = out.sz;                                           // old interval of the re-hung subtree
        const int b = in_side1 ? inS : inF;                                         // in[v_in]
        const int dp_uin = in_side1 ? dpF : dpS, dp_vin = in_side1 ? dpS : dpF;
        // stem = cycle nodes on u_in's side from u_in (index 0, deepest) up to u_out (index ns-1); depths give the index
        const int ns = change ? dp_uin - out.dp + 1 : 1;
        const bool longstem = ns > kTeamStemCap;
        D.change = change; D.delta = delta; D.val = (long long)E.state * delta;     // NS.cs:1017
        D.in_side1 = in_side1; D.a = a; D.s = s; D.dp_uin = dp_uin; D.ns = ns; D.longstem = longstem; D.out = out;
        D.dir_new_up = u_in == E.src;                                               // NS.cs:1143
        D.inF = inF; D.inS = inS; D.first = first;

        // ---- STEM(k) (26 % of pivots): the stem is longer than one node; its owners publish the entries, index = depth
        if (ns > 1) {
            int4* const stem_g = P.stemseg + (size_t)par * (n + 1) * 2;             // entry of stem index k at slot t = ns-1-k
            if (probe_thr) sh.bk.stem_x++;
            if constexpr (!kPricer) {
                auto publish = [&](int j, int in_u, int sz_u, int pd, int dp, bool hasF) {
                    const long long fl = D.new_flow((long long)fl_s[j], pd, hasF, lower_state);
                    const long long upl = FT::cap_out(up_s[j]);                      // capacity travels too when it fits 31 bits (-1: fetch)
                    int4* e = stem_g + (size_t)(dp - out.dp) * 2;
                    st_mail(e, make_int4(in_u, sz_u, pd, seq));
                    st_mail(e + 1, make_int4(lo32(fl), hi32(fl), upl == LLONG_MAX / 2 ? INT_MAX : (upl < (long long)INT_MAX ? (int)upl : -1), seq));
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
                        if (hasF != hasS && hasF == in_side1 && in_u >= a) publish(j, in_u, sz_u, pd_s[j], dp_s[j], hasF);
                    }
                }
            }
            if (!kPricer && !longstem) {
                for (int q = tid; q < ns; q += kTT) {
                    int4 w[2];
                    if (!poll_rec<2>(stem_g + (size_t)q * 2, seq, w, P, 6)) sh.abort = 1;
                    const int kx = ns - 1 - q;
                    st_in[kx] = w[0].x; st_z[kx] = w[0].y; st_pd[kx] = w[0].z; st_fl[kx] = mk64(w[1].x, w[1].y); st_up[kx] = w[1].z;
                }
                __syncthreads();
                if (sh.abort) [[unlikely]] return ST_ERR_BARRIER_TIMEOUT;
            }
            TICK(t_stem);
        }
        if (change && probe_thr) { if (ns > sh.bk.max_stem) sh.bk.max_stem = ns; sh.bk.moved_nodes += s; }
        U.valid = 1; U.change = change ? 1 : 0; U.a = a; U.s = s; U.b = b; U.ns = ns; U.longstem = longstem ? 1 : 0;
        U.dshift = dp_vin + 1 - dp_uin; U.par = par; U.seq = seq;
        U.sigma = D.dir_new_up ? -E.rcb : E.rcb;     // pi[v_in] - pi[u_in] -/+ cost (NS.cs:1187-1188), u_in being the source or the target
        return 0;
    };

    if (pricer) {
        // ========================================================================================== the pricing CTA
        // BlockSearchPivot fields (NS.cs:1294-1302)
        int next_arc = 0, B = P.block_size;
        int ticket = 0, last_tk = 0;                     // last staging request issued; the one before it
        int tag = 0;                                     // rounds posted so far (the tag of the last one)
        // Staging area.  A block [cursor, cursor + cnt) of the arc arrays is laid out in POSITIONS so that every array is copied as
        // aligned 128-bit words: piece 1 = arcs up to the end of the arrays at positions d0 .., piece 2 (after the wrap) at p2 ..;
        // positions that hold no arc of the block are neutral (state 0).
        //   cold part (one copy: src, tgt, capacity, state) - staged with cp.async for the block of the NEXT pivot, exactly known
        //   hot part (two copies by pivot parity: cost + pi_s - pi_t and both labels) - the node records the owners serve.  The block
        //   of pivot k+2 is requested with ENTER(k) at its predicted place (the search of pivot k+1 ends in its first block 9 times out
        //   of 10), served by the owners off their critical path from the basis before update k, collected here one pivot later, and
        //   priced with updates k and k+1 replayed in closed form.  So in the steady state pricing waits for nobody.
        int sg_cursor = -1, sg_cnt = 0, sg_d0 = 0, sg_n1 = 0, sg_p2 = 0, sg_pt = 0;     // cold part: which block, its layout
        int sg_plo = 0, sg_phi = 0;                      // ... and the positions this pricer stages and prices (whole 16-byte chunks)
        // hot copies 0 / 1: which block, as of which basis (low 32 bits of the pivot index), whether its records are in shared memory
        int hb_cur0 = -1, hb_cur1 = -1, hb_cnt0 = 0, hb_cnt1 = 0, hb_basis0 = 0, hb_basis1 = 0;
        bool hb_ok0 = false, hb_ok1 = false;
        auto hb_set = [&](int h, int cur, int cnt, int basis, bool ok) {
            if (h) { hb_cur1 = cur; hb_cnt1 = cnt; hb_basis1 = basis; hb_ok1 = ok; } else { hb_cur0 = cur; hb_cnt0 = cnt; hb_basis0 = basis; hb_ok0 = ok; }
        };
        Pending U1, U2;                                  // the updates of pivots k-2 and k-1, replayed on the staged node records
        U1.valid = U1.change = U1.a = U1.s = U1.b = U1.longstem = U1.dshift = U1.par = U1.seq = 0; U1.ns = 1; U1.sigma = 0;
        U2 = U1;

        auto layout = [&](int cursor, int cnt) {
            sg_cursor = cursor; sg_cnt = cnt; sg_d0 = cursor & 3; sg_n1 = min(cnt, S - cursor);
            sg_p2 = (sg_d0 + sg_n1 + 3) & ~3;
            sg_pt = cnt > sg_n1 ? sg_p2 + ((cnt - sg_n1 + 3) & ~3) : sg_p2;
            const int nch = sg_pt >> 2;
            sg_plo = 4 * (cta * nch / NP); sg_phi = 4 * ((cta + 1) * nch / NP);
        };
        auto pos_valid = [&](int p) -> bool { return p < sg_p2 ? (unsigned)(p - sg_d0) < (unsigned)sg_n1 : p - sg_p2 < sg_cnt - sg_n1; };
        auto pos_arc = [&](int p) -> int { return p < sg_p2 ? sg_cursor - sg_d0 + p : p - sg_p2; };
        // cold part: src / tgt / capacity are immutable and go global -> shared with 16-byte cp.async; `state` is mutable (this
        // CTA is its only writer) and is read around L1 by stage_finish().  Threads [t0, kTT) take part in stage_begin.
        auto stage_begin = [&](int t0) {
            for (int c = (sg_plo >> 2) + tid - t0; c < (sg_phi >> 2); c += kTT - t0) {
                const int p = 4 * c;
                const int g = p < sg_p2 ? sg_cursor - sg_d0 + p : p - sg_p2;
                cp_async16(pf_src + p, P.src + g); cp_async16(pf_tgt + p, P.tgt + g);
                cp_async16(pf_up + p, P.upper + g); cp_async16(pf_up + p + 2, P.upper + g + 2);
                cp_async16(pf_st + p, P.state + g);                      // (.cg: from L2, where this CTA's own state stores are)
                cp_async16(pf_cost + p, P.cost + g);
            }
        };
        auto stage_finish = [&]() {
            cp_async_wait_all();
            __syncthreads();                                            // arc data and states are visible to every thread
            // positions of the aligned chunks that lie outside the block (at most three at each end of a piece) price as state 0
            // (the caller's next barrier orders these stores before the pricing loop)
            if (tid >= 32 && tid < 48) {
                const int e = (tid - 32) >> 2, i = tid & 3;              // e: 0 before piece 1, 1 behind it, 2 behind piece 2 (3: unused)
                const int p = e == 0 ? i : e == 1 ? sg_d0 + sg_n1 + i : sg_p2 + (sg_cnt - sg_n1) + i;
                const int lim = e == 1 ? sg_p2 : sg_pt;
                if (e < 3 && p < lim && !pos_valid(p)) pf_st[p] = 0;
            }
        };
        // post a staging request "owners: write {pi, in} of both ends of arcs [cursor, cursor + cnt) into stage buffer `buf`" (word 4 of the ENTER line)
        auto post_request = [&](int par, int seq, int cursor, int cnt, int tk, int buf) {
            if (cta == 0 && warp == 0 && lane < kRepEnt) st_mail(P.ent + ((size_t)par * kRepEnt + lane) * NP * kMailWords + 4, make_int4(cursor, cnt, tk * 4 + buf, seq));
        };
        // collect the served node records of block (cursor, cnt), request `tk` in stage buffer `buf`, into hot copy h: reduced-cost base and
        // labels per position.  Spins until complete; false = abandoned.
        auto collect = [&](int h, int cursor, int cnt, int tk, int buf, bool probes) -> bool {
            const int d0 = cursor & 3, n1 = min(cnt, S - cursor), p2 = (d0 + n1 + 3) & ~3;
            const int pt = cnt > n1 ? p2 + ((cnt - n1 + 3) & ~3) : p2;
            const int nch = pt >> 2;
            const int plo = 4 * (cta * nch / NP), phi = 4 * ((cta + 1) * nch / NP);
            long long* const rcb = pf_rcb + h * kStagePos;
            int2* const lab = pf_lab + h * kStagePos;
            const int4* const sbuf = P.stage + (size_t)buf * 2 * kStagePos;
            const int tkw = tk * 4 + buf;
            // the owners serve every position of the aligned chunks (the few that lie outside the block are real arcs too, or the
            // zero padding behind the arrays; pricing masks them by state 0)
            unsigned missing = 0;
#pragma unroll
            for (int j = 0; j < kPos; ++j) if (plo + tid + j * kTT < phi) missing |= 1u << j;
            unsigned spins = 0; long long t0 = 0;
            for (;;) {
#pragma unroll
                for (int jb = 0; jb < kPos; jb += 3) {                   // three record pairs in flight
                    int4 vs[3], vt[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) if (jb + j < kPos && (missing >> (jb + j) & 1u)) {
                        const int q = plo + tid + (jb + j) * kTT;
                        vs[j] = ld_mail(sbuf + 2 * q); vt[j] = ld_mail(sbuf + 2 * q + 1);
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) if (jb + j < kPos && (missing >> (jb + j) & 1u)) {
                        const int q = plo + tid + (jb + j) * kTT;
                        if (vs[j].w == tkw && vt[j].w == tkw) {
                            rcb[q] = mk64(vs[j].x, vs[j].y) - mk64(vt[j].x, vt[j].y);
                            lab[q] = make_int2(vs[j].z, vt[j].z);
                            missing &= ~(1u << (jb + j));
                        }
                    }
                    if (jb == 0 && probes) PROBE(7);
                }
                if (probes) PROBE(6);
                if (!__syncthreads_or(missing != 0)) return true;
                if (spin_check(spins, t0, P, 7)) sh.abort = 1;
                if (__syncthreads_or(sh.abort)) return false;
            }
        };
        // one end of an arc through the pending updates: label x as of the basis of the records, `nrep` updates to replay.
        // Returns the sum of the sigmas of the updates that moved the node (UpdatePotentials, NS.cs:1185-1209); x becomes the label now.
        auto replay_end = [&](int& x, int nrep) -> long long {
            long long add = 0;
            if (nrep == 2 && U1.change) {
                if ((unsigned)(x - U1.a) < (unsigned)U1.s) add += U1.sigma;
                int nx, nd; relabel(U1, x, 0, nx, nd); x = nx;
            }
            if (nrep >= 1 && U2.change) {
                if ((unsigned)(x - U2.a) < (unsigned)U2.s) add += U2.sigma;
                int nx, nd; relabel(U2, x, 0, nx, nd); x = nx;
            }
            return add;
        };

        // warp 0: read every pricer's record of round r of this pivot; 1 = the round has a winner (written to sh.win), 0 = none
        // (every word of a record carries the TAG of the round it was posted for - rounds are numbered through the whole solve, the
        // same in every CTA - so that words of two rounds of one pivot can never be taken for one record)
        auto read_records = [&](int par, int tg) -> int {
            const int4* const line = P.ent + ((size_t)par * kRepEnt + cta % kRepEnt) * NP * kMailWords;
            int4 v = make_int4(0, 0, 0, 0);
            unsigned spins = 0; long long t0 = 0;
            for (;;) {
                if (lane < 4 * NP) v = ld_mail(line + (lane >> 2) * kMailWords + (lane & 3));
                const bool ok = lane >= 4 * NP || v.w == tg;
                if (__all_sync(0xffffffffu, ok)) break;
                // a pricer that is already a round further has seen only "none" in this one
                const bool ahead = lane < 4 * NP && v.w - tg == 1;
                if (__any_sync(0xffffffffu, ahead)) return 0;
                if (spin_check(spins, t0, P, 8)) { sh.abort = 1; return 0; }
            }
            const int arc = __shfl_sync(0xffffffffu, v.x, (lane & ~3));
            const int st = __shfl_sync(0xffffffffu, v.x, (lane & ~3) | 1);
            const int r_lo = __shfl_sync(0xffffffffu, v.x, (lane & ~3) | 2), r_hi = __shfl_sync(0xffffffffu, v.y, (lane & ~3) | 2);
            const int pp = __shfl_sync(0xffffffffu, v.z, (lane & ~3) | 2);
            const int best = warp_argmin(lane < 4 * NP && (lane & 3) == 0 && arc >= 0, (long long)st * mk64(r_lo, r_hi), pp);
            if (best < 0) return 0;
            const int4 w0 = make_int4(__shfl_sync(0xffffffffu, v.x, best), __shfl_sync(0xffffffffu, v.y, best), __shfl_sync(0xffffffffu, v.z, best), 0);
            const int4 w1 = make_int4(__shfl_sync(0xffffffffu, v.x, best + 1), __shfl_sync(0xffffffffu, v.y, best + 1), __shfl_sync(0xffffffffu, v.z, best + 1), 0);
            const int4 w2 = make_int4(__shfl_sync(0xffffffffu, v.x, best + 2), __shfl_sync(0xffffffffu, v.y, best + 2), 0, 0);
            const int4 w3 = make_int4(__shfl_sync(0xffffffffu, v.x, best + 3), __shfl_sync(0xffffffffu, v.y, best + 3), 0, 0);
            if (lane == 0) { Ent e; e.arc = w0.x; e.src = w0.y; e.tgt = w0.z; e.state = w1.x; e.in_s = w1.y; e.in_t = w1.z; e.rcb = mk64(w2.x, w2.y); e.upper = mk64(w3.x, w3.y); sh.win = e; }
            return 1;
        };

        for (;;) {
            const long long k = iterations + 1;
            const int seq = (int)(unsigned)k;
            const int par = (int)(k & 1);
            const int h = par;
            TICK(t_wdone);
            if (probe_thr) sh.bk.pr_mark = (unsigned long long)clock64();
            // ================================================================ FindEnteringArc (NS.cs:1339-1397), post ENTER(k)
            // Round r prices block r of the scan, offsets [r * B, (r + 1) * B) from the cursor; every pricer its share of the positions.
            // Round 0 is staged (see above); when it is not (mispredicted place, first pivots) and in later rounds (one pivot in ten)
            // the block is requested, staged and collected here.  After each round the pricers post their candidates; everybody
            // (pricers and owners) reads all of them and picks the same winner: smallest reduced cost, then first in scan order.
            int search_end = 0, nrep = 0, win_tag = 0;
            bool have_win = false, win_read = false;                  // win_read: sh.win holds the winner already
            bool explicit_req = false;                                // this search posted an explicit staging request
            for (int r = 0;; ++r) {
                const long long o_lo = (long long)r * B;
                const int cnt = (int)min((long long)B, (long long)S - o_lo);
                const bool last_round = o_lo + cnt >= S;
                int cur = next_arc + (int)o_lo; if (cur >= S) cur -= S;
                nrep = (int)(unsigned)(k - 1) - (h ? hb_basis1 : hb_basis0);
                if (!(r == 0 && (h ? hb_ok1 : hb_ok0) && (h ? hb_cur1 : hb_cur0) == cur && (h ? hb_cnt1 : hb_cnt0) == cnt && (unsigned)nrep <= 2u && sg_cursor == cur && sg_cnt == cnt)) {
                    ++ticket;
                    explicit_req = true;
                    __syncthreads();                                        // the staging area is no longer read
                    post_request(par, seq, cur, cnt, ticket, 2);
                    layout(cur, cnt);
                    stage_begin(0);
                    if (!collect(h, cur, cnt, ticket, 2, false)) { status = ST_ERR_BARRIER_TIMEOUT; break; }
                    stage_finish();
                    __syncthreads();
                    hb_set(h, cur, cnt, (int)(unsigned)(k - 1), true); nrep = 0;
                    if (probe_thr) sh.bk.rounds_total++;
                }
                // ---- the pricing loop: two positions per 128-bit shared-memory load
                const long long* const rcb = pf_rcb + h * kStagePos;
                const int2* const lab = pf_lab + h * kStagePos;
                // interval tests of the updates to replay ([a, a+s): re-hung subtree, [lo, lo+len): labels that shift by `by`); s == 0: no-op
                const bool r1 = nrep == 2 && U1.change, r2 = nrep >= 1 && U2.change;
                const unsigned a1 = (unsigned)U1.a, s1 = r1 ? (unsigned)U1.s : 0u, a2 = (unsigned)U2.a, s2 = r2 ? (unsigned)U2.s : 0u;
                const unsigned lo1 = (unsigned)(U1.b < U1.a ? U1.b + 1 : U1.a + U1.s), len1 = r1 ? (unsigned)(U1.b < U1.a ? U1.a - U1.b - 1 : U1.b - U1.a - U1.s + 1) : 0u;
                const int by1 = U1.b < U1.a ? U1.s : -U1.s;
                long long bk = 0;
                int bp = -1;
                for (int i = (sg_plo >> 1) + tid; i < (sg_phi >> 1); i += kTT) {
                    const longlong2 vv = reinterpret_cast<const longlong2*>(rcb)[i];
                    const int4 ll = reinterpret_cast<const int4*>(lab)[i];
                    const int2 ss = reinterpret_cast<const int2*>(pf_st)[i];
                    const int2 cc = reinterpret_cast<const int2*>(pf_cost)[i];
                    long long v0 = vv.x + cc.x, v1 = vv.y + cc.y;            // cost + pi_s - pi_t
                    int xa = ll.x, xb = ll.y, xc = ll.z, xd = ll.w;
                    // rare: an end was re-hung by the older update - its new label needs the stem (relabel)
                    if (((unsigned)xa - a1 < s1) | ((unsigned)xb - a1 < s1) | ((unsigned)xc - a1 < s1) | ((unsigned)xd - a1 < s1)) [[unlikely]] {
                        v0 += replay_end(xa, nrep) - replay_end(xb, nrep);
                        v1 += replay_end(xc, nrep) - replay_end(xd, nrep);
                    } else {
                        if ((unsigned)xa - lo1 < len1) xa += by1;
                        if ((unsigned)xb - lo1 < len1) xb += by1;
                        if ((unsigned)xc - lo1 < len1) xc += by1;
                        if ((unsigned)xd - lo1 < len1) xd += by1;
                        if ((unsigned)xa - a2 < s2) v0 += U2.sigma;
                        if ((unsigned)xb - a2 < s2) v0 -= U2.sigma;
                        if ((unsigned)xc - a2 < s2) v1 += U2.sigma;
                        if ((unsigned)xd - a2 < s2) v1 -= U2.sigma;
                    }
                    v0 *= (long long)ss.x;                                   // state is -1, 0 or +1 (SpanningTree.cs:53-71)
                    v1 *= (long long)ss.y;
                    if (v0 < bk) { bk = v0; bp = 2 * i; }
                    if (v1 < bk) { bk = v1; bp = 2 * i + 1; }
                }
                const int wl = warp_argmin(bp >= 0, bk, bp);                 // positions ascend with the scan offset: lowest position = first in scan order
                if (lane == 0) sh.pk[warp] = make_longlong2(0, -1);
                __syncwarp();
                if (wl >= 0 && lane == wl) {
                    // every warp prepares the full record of its own candidate (the pending updates replayed on it), all warps side by
                    // side: what is left to do after the barrier is one arg-min and the post
                    const int w_p = bp;
                    int w_ins, w_int;
                    { const int2 lb = pf_lab[h * kStagePos + w_p]; w_ins = lb.x; w_int = lb.y; }
                    long long w_rcb = pf_rcb[h * kStagePos + w_p] + pf_cost[w_p];
                    w_rcb += replay_end(w_ins, nrep); w_rcb -= replay_end(w_int, nrep);
                    const long long w_up = pf_up[w_p];
                    sh.pk[warp] = make_longlong2(bk, bp);
                    sh.crec[warp][0] = make_int4(pos_arc(w_p), pf_src[w_p], pf_tgt[w_p], tag + 1);
                    sh.crec[warp][1] = make_int4(pf_st[w_p], w_ins, w_int, tag + 1);
                    sh.crec[warp][2] = make_int4(lo32(w_rcb), hi32(w_rcb), w_p, tag + 1);
                    sh.crec[warp][3] = make_int4(lo32(w_up), hi32(w_up), last_round ? 1 : 0, tag + 1);
                }
                __syncthreads();
                if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
                // ---- warp 0: this pricer's candidate = arg-min of (rc, position) over the warp winners, posted at once
                const longlong2 q = sh.pk[lane & (kTW - 1)];
                const int ww = warp_argmin(lane < kTW && q.y >= 0, q.x, (int)q.y);      // (every warp computes it: `mine` is needed by all)
                const bool mine = ww >= 0;
                ++tag;                                                       // the tag of this round
                if (warp == 0 && lane < 4 * kRepEnt) {
                    const int wd = lane & 3;
                    int4 o = mine ? sh.crec[ww][wd] : (wd == 3 ? make_int4(0, 0, last_round ? 1 : 0, tag) : make_int4(-1, 0, 0, tag));
                    st_mail(P.ent + (((size_t)par * kRepEnt + (lane >> 2)) * NP + cta) * kMailWords + wd, o);
                }
                PROBE(0);
                win_tag = tag;
                // (pricer 0 after an explicit request reads the records at once: the request word may only be overwritten - by the next
                // request - when every pricer has collected what the owners served for this one, i.e. has posted its record)
                if (mine && !(explicit_req && cta == 0)) {                                 // this round has a winner (which one is read later)
                    have_win = true; search_end = (int)(o_lo + cnt);
                    if (NP == 1) {
                        if (tid == 0) { const int4 c0 = sh.crec[ww][0], c1 = sh.crec[ww][1], c2r = sh.crec[ww][2], c3r = sh.crec[ww][3];
                            Ent e; e.arc = c0.x; e.src = c0.y; e.tgt = c0.z; e.state = c1.x; e.in_s = c1.y; e.in_t = c1.z; e.rcb = mk64(c2r.x, c2r.y); e.upper = mk64(c3r.x, c3r.y); sh.win = e; }
                        win_read = true;
                    }
                    break;
                }
                // no candidate here: did another pricer find one?
                if (warp == 0) {
                    const int outcome = read_records(par, tag);
                    if (lane == 0) sh.mode = outcome;
                }
                __syncthreads();
                if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
                if (sh.mode) { have_win = true; win_read = true; search_end = (int)(o_lo + cnt); break; }
                if (last_round) { search_end = S; break; }
            }
            if (status == ST_ERR_BARRIER_TIMEOUT) break;
            // NS.cs:1397-1438: cursor, counters, adaptive block size
            if (probe_thr) { sh.bk.arcs_checked += search_end; sh.bk.rounds_total++; }
            if (have_win) {
                const int Bold = B;
                if (P.adaptive) {
                    const double hit = search_end > 0 ? 1.0 / search_end : 0;
                    if (hit < P.low_thr) {
                        cons_high = 0; cons_low++;
                        if (cons_low >= P.consecutive) { const int ns = (int)(B * P.shrink); B = P.dyn_min_block > ns ? P.dyn_min_block : ns; cons_low = 0; }
                    } else if (hit > P.high_thr) {
                        cons_low = 0; cons_high++;
                        if (cons_high >= P.consecutive) { const int ns = (int)(B * P.grow); B = P.max_block_size < ns ? P.max_block_size : ns; cons_high = 0; }
                    } else { cons_low = 0; cons_high = 0; }
                }
                // `_nextArc = e` (NS.cs:1397): the last arc examined, or unchanged after a full sweep that ended inside a block
                if (search_end < S || S % Bold == 0) { int e = next_arc + search_end - 1; if (e >= S) e -= S; next_arc = e; }
            }
            const int nb0 = B < S ? B : S;
            // the block of pivot k+2 at its predicted place: where the cursor ends up if the next search stops in its first block
            int c2 = next_arc;
            if (nb0 < S || S % B == 0) { c2 = next_arc + nb0 - 1; if (c2 >= S) c2 -= S; }
            if (have_win) { ++ticket; post_request(par, seq, c2, nb0, ticket, h); }
            if (!have_win) { status = ST_OPTIMAL; break; }
            iterations = k;
            if (iterations > P.max_iterations) { status = ST_INFEASIBLE; break; }          // NS.cs:311-317
            // what hot copy h will hold: the block requested just now (served from the basis before update k)
            const int req_tk = ticket;
            hb_set(h, c2, nb0, (int)(unsigned)(k - 1), false);
            layout(next_arc, nb0);
            TICK(t_price);
            PROBE(1);
            // ---- while the owners scan: collect the node records of the next pivot's block (requested one pivot ago, served long ago)
            // when they are for the right place; otherwise the next pricing requests them itself
            {
                const int h1 = h ^ 1;
                const int o_cur = h1 ? hb_cur1 : hb_cur0, o_cnt = h1 ? hb_cnt1 : hb_cnt0;
                bool ok = false;
                if (o_cur == next_arc && o_cnt == nb0 && k >= 2) {
                    if (!collect(h1, next_arc, nb0, last_tk, h1, true)) { status = ST_ERR_BARRIER_TIMEOUT; break; }
                    ok = true;
                }
                if (h1) hb_ok1 = ok; else hb_ok0 = ok;
            }
            last_tk = req_tk;
            PROBE(2);
            // ---- cold part of the next pivot's block (exactly known) streams in behind the records (loads complete in issue order
            // on an SM: the records, L2 hits, must not queue behind DRAM misses); the block after it is pulled into L2
            stage_begin(0);
            int c3 = c2;                                                    // the block after the one requested just now
            if (nb0 < S) { c3 = c2 + nb0 - 1; if (c3 >= S) c3 -= S; }
            for (int q = (sg_plo + tid * 32); q < sg_phi; q += kTT * 32) {
                int idx = c3 + q; if (idx >= S) idx -= S;
                prefetch_l2(P.src + idx); prefetch_l2(P.tgt + idx); prefetch_l2(P.cost + idx); prefetch_l2(P.state + idx);
                prefetch_l2(P.upper + idx); prefetch_l2(P.upper + min(idx + 16, S - 1));
            }
            if (!win_read && warp == 0) read_records(par, win_tag);         // (the other pricers' candidates have long arrived)
            __syncthreads();
            if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
            const Ent E = sh.win;
            PROBE(3);
            Dec D; Pending U;
            const int rcd = gather_decide.template operator()<true>(seq, par, E, 0, D, U);
            if (rcd != 0) { status = rcd; break; }
            stage_finish();                                                         // (arrived while CYC(k) was in flight)
            // arc states (ChangeFlow, NS.cs:1031-1039): only the pricing scans read them - state[] in global memory and, when the arc
            // lies in the block staged for the next pivot, its copy in shared memory, staged before this decision
            if (tid < 2) {
                const int arc = tid == 0 ? E.arc : (D.change ? D.out.pd >> 1 : -1);
                const int st = tid == 0 ? (D.change ? STATE_TREE : -E.state) : ((D.out.zero & 1) ? STATE_LOWER : STATE_UPPER);
                if (arc >= 0) {
                    P.state[arc] = st;
                    int off = arc - sg_cursor; if (off < 0) off += S;
                    if (off < sg_cnt) pf_st[off < sg_n1 ? sg_d0 + off : sg_p2 + off - sg_n1] = st;
                }
            }
            U1 = U2; U2 = U; U2.longstem = 1;                                       // (this CTA stages no stems: relabel reads them in place)
            __syncthreads();
            TICK(t_update);
            PROBE(5);
            if (P.stop_after > 0 && iterations >= P.stop_after) { status = ST_STOPPED_EARLY; break; }
        }
        if (tid == 0) sh.mode = B;                                                   // final block size, for the epilogue
    } else {
        // ========================================================================================== the owner CTAs
        int ticket = 0;                                  // last staging request served
        int etag = 0;                                    // rounds of the pricers consumed so far (see read_records)
        // serve a staging request: for every end of arcs [cursor, cursor + cnt) that this CTA owns, write {pi, in, ticket} - the
        // node's record as of the basis this CTA holds right now - into the pricer's staging slots
        auto serve = [&](int cursor, int cnt, int tk) {
            int4* const sbuf = P.stage + (size_t)(tk & 3) * 2 * kStagePos;
            // the range is one or (when it wraps at S) two linear pieces of the arc arrays; each is read as aligned 128-bit words
            // (the slots are the pricer's staging positions: piece 1 at d0 .., piece 2 at the next multiple of four)
            int seg_a = cursor, seg_n = min(cnt, S - cursor), pbase = cursor & 3, done_n = 0;
            for (int piece = 0; piece < 2 && seg_n > 0; ++piece) {
                const int a0 = seg_a & ~3;
                const int nch = ((seg_a + seg_n + 3) >> 2) - (a0 >> 2);
                for (int ch0 = tid; ch0 < nch; ch0 += 2 * kTT) {            // two chunks of four arcs per thread in flight
                    int4 s4[2], t4[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int ch = ch0 + u * kTT;
                        if (ch < nch) { s4[u] = __ldg(reinterpret_cast<const int4*>(P.src + a0) + ch); t4[u] = __ldg(reinterpret_cast<const int4*>(P.tgt + a0) + ch); }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int ch = ch0 + u * kTT;
                        if (ch < nch) {
                            const int sv[4] = {s4[u].x, s4[u].y, s4[u].z, s4[u].w}, tv[4] = {t4[u].x, t4[u].y, t4[u].z, t4[u].w};
                            const int ob = a0 + 4 * ch - seg_a;                 // offset of the chunk's first arc inside the piece
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                // (every element of the aligned chunk, also the few outside the block: the pricer expects a record per position)
                                const unsigned js = (unsigned)(sv[e] - lo), jt = (unsigned)(tv[e] - lo);
                                if (js < (unsigned)cntn) { const long long p = __ldcg(P.pi + sv[e]); st_mail(sbuf + 2 * (pbase + ob + e), make_int4(lo32(p), hi32(p), in_s[js], tk)); }
                                if (jt < (unsigned)cntn) { const long long p = __ldcg(P.pi + tv[e]); st_mail(sbuf + 2 * (pbase + ob + e) + 1, make_int4(lo32(p), hi32(p), in_s[jt], tk)); }
                            }
                        }
                    }
                }
                pbase = ((cursor & 3) + seg_n + 3) & ~3;                        // piece 2 starts on a fresh 16-byte boundary of the staging arrays
                done_n += seg_n; seg_a = 0; seg_n = cnt - done_n;
            }
        };
        for (;;) {
            const long long k = iterations + 1;
            const int seq = (int)(unsigned)k;
            const int par = (int)(k & 1);
            const int4* const line = P.ent + ((size_t)par * kRepEnt + cta % kRepEnt) * NP * kMailWords;   // the pricers' records, this CTA's replica
            // ================================================================ wait for ENTER(k), serving staging requests meanwhile
            for (;;) {
                if (warp == 0) {
                    unsigned spins = 0; long long t0 = 0;
                    int4 v = make_int4(0, 0, 0, 0);
                    int mode = 0;
                    for (;;) {
                        // lanes [0, 4 NP): word (lane & 3) of pricer (lane >> 2)'s record; lane 4 NP: the staging request (pricer 0's word 4)
                        if (lane < 4 * NP) v = ld_mail(line + (lane >> 2) * kMailWords + (lane & 3));
                        else if (lane == 4 * NP) v = ld_mail(line + 4);
                        const int rd0 = __shfl_sync(0xffffffffu, v.z, 3);                    // last-round flag (the same in every record of a round)
                        const bool ok = lane >= 4 * NP || v.w == etag + 1;
                        // the pricers are further than this CTA thought (rounds whose blocks hold none of its nodes pass without it):
                        // everything before the newest tag on display found nothing
                        const int lead = __reduce_max_sync(0xffffffffu, lane < 4 * NP ? v.w - (etag + 1) : 0);
                        if (lead > 0) { etag += lead; continue; }
                        if (__all_sync(0xffffffffu, ok)) {
                            // every record is of this pivot and of the same round: the one with the smallest reduced cost, then the first
                            // in scan order, is the entering arc; none at all = the round found nothing (the last round: optimal)
                            const int arc = __shfl_sync(0xffffffffu, v.x, (lane & ~3));
                            const int st = __shfl_sync(0xffffffffu, v.x, (lane & ~3) | 1);
                            const int r_lo = __shfl_sync(0xffffffffu, v.x, (lane & ~3) | 2), r_hi = __shfl_sync(0xffffffffu, v.y, (lane & ~3) | 2);
                            const int pp = __shfl_sync(0xffffffffu, v.z, (lane & ~3) | 2);
                            const int best = warp_argmin(lane < 4 * NP && (lane & 3) == 0 && arc >= 0, (long long)st * mk64(r_lo, r_hi), pp);
                            if (best >= 0 || (rd0 & 1)) {
                                const int b0 = best >= 0 ? best : 0;
                                const int4 w = make_int4(__shfl_sync(0xffffffffu, v.x, b0 + (lane & 3)), __shfl_sync(0xffffffffu, v.y, b0 + (lane & 3)), __shfl_sync(0xffffffffu, v.z, b0 + (lane & 3)), seq);
                                if (lane < 4) sh.ent[lane] = best >= 0 ? w : make_int4(-1, 0, 0, seq);
                                ++etag;
                                mode = 1; break;
                            }
                            ++etag;                                        // nothing in this round, and it was not the last one
                            continue;
                        }
                        const int4 rq = make_int4(__shfl_sync(0xffffffffu, v.x, 4 * NP), __shfl_sync(0xffffffffu, v.y, 4 * NP), __shfl_sync(0xffffffffu, v.z, 4 * NP), __shfl_sync(0xffffffffu, v.w, 4 * NP));
                        if (rq.w == seq && rq.z != ticket) { if (lane == 0) sh.ent[4] = rq; mode = 2; break; }
                        if (spin_check(spins, t0, P, 9)) { mode = 3; break; }
                    }
                    if (mode == 1 && lane == 0) sh.ent[4] = make_int4(0, 0, 0, 0);             // (the request that comes with ENTER is fetched after the scan)
                    if (lane == 0) sh.mode = mode;
                }
                __syncthreads();
                const int mode = sh.mode;
                if (mode == 2) {
                    const int4 rq = sh.ent[4];
                    serve(rq.x, rq.y, rq.z); ticket = rq.z;
                    __syncthreads();
                    continue;
                }
                break;
            }
            if (sh.mode == 3) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
            PROBE(9);
            Ent E;
            int4 nreq;                                                   // the staging request that came with ENTER(k)
            {
                const int4 r0 = sh.ent[0], r1 = sh.ent[1], r2 = sh.ent[2], r3 = sh.ent[3];
                nreq = sh.ent[4];
                E.arc = r0.x; E.src = r0.y; E.tgt = r0.z; E.state = r1.x; E.in_s = r1.y; E.in_t = r1.z;
                E.rcb = mk64(r2.x, r2.y); E.upper = mk64(r3.x, r3.y);
            }
            if (E.arc < 0) { status = ST_OPTIMAL; break; }
            iterations = k;
            if (iterations > P.max_iterations) { status = ST_INFEASIBLE; break; }          // NS.cs:311-317
            const bool lower_state = E.state == STATE_LOWER;
            const int first = lower_state ? E.src : E.tgt, second = lower_state ? E.tgt : E.src;      // NS.cs:948-957
            const int inF = lower_state ? E.in_s : E.in_t, inS = lower_state ? E.in_t : E.in_s;

            // ================================================================ cycle discovery over the slice, post CYC(k)
            // A node is on the pivot cycle iff exactly one end of the entering arc lies in its subtree (FindJoinNode + both walks
            // of FindLeavingArc, NS.cs:925-1010, as one interval test per node).
            auto make_cand = [&](int j, int in_u, int sz_u, bool hasF) -> Cand {
                const int pd = pd_s[j];
                const F fl = fl_s[j], up = up_s[j];
                const bool dir_up = pd & 1;
                // first walk: residual capacity when pred_dir == DOWN, else the flow; second walk mirrored (NS.cs:968, :986)
                const bool increase = hasF ? !dir_up : dir_up;
                Cand cd; cd.d = increase ? FT::residual(up, fl) : (long long)fl; cd.in = in_u; cd.sz = sz_u; cd.pd = pd; cd.dp = dp_s[j]; cd.j = j;
                cd.zero = (((!increase) || up == 0) ? 1 : 0) | (hasF ? 2 : 0);
                return cd;
            };
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
            const int nc = sh.ncand;
            PROBE(10);
            {
                // the record goes out in kRepCyc copies (lane l writes word l % 5 of copy l / 5), stored word-major
                int4* const rec = P.cyc + ((size_t)par * kRepCyc + lane / 5) * 5 * Gp + cta;
                // word 0 also carries the depth of the entering arc's ends (whoever owns them) and the int32-overflow flag of narrow mode
                int w0x = nc > 0xffff ? 0xffff : nc, dF = 0, dS = 0;
                if ((unsigned)(first - lo) < (unsigned)cntn) { w0x |= 1 << 20; dF = dp_s[first - lo]; }
                if ((unsigned)(second - lo) < (unsigned)cntn) { w0x |= 1 << 21; dS = dp_s[second - lo]; }
                if (sh.ovf) w0x |= 1 << 22;
                if (nc == 0) {
                    if (warp == 0 && lane < 5 * kRepCyc && lane % 5 == 0) st_mail(rec, make_int4(w0x, dF, dS, seq));       // no candidates: word 0 is all anyone reads
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
                        __syncthreads();                                    // sh.wc is reused by the CYC gather
                    }
                    if (warp == 0 && lane < 5 * kRepCyc) {
                        const int wd = lane % 5;
                        int4 w;
                        if (wd == 0) w = make_int4(w0x | ((m1.zero & 1) << 16) | ((m2.zero & 1) << 17) | (m1.pd >= 0 ? 1 << 18 : 0) | (m2.pd >= 0 ? 1 << 19 : 0), dF, dS, seq);
                        else if (wd == 1) w = make_int4(lo32(m1.d), hi32(m1.d), m1.in, seq);
                        else if (wd == 2) w = make_int4(m1.sz, m1.pd, m1.dp, seq);
                        else if (wd == 3) w = make_int4(lo32(m2.d), hi32(m2.d), m2.in, seq);
                        else w = make_int4(m2.sz, m2.pd, m2.dp, seq);
                        st_mail(rec + (size_t)wd * Gp, w);
                    }
                }
            }
            PROBE(11);
            // ---- off the critical path: serve the staging request that came with ENTER(k) - the block of pivot k+2 at its predicted
            // place.  The basis this CTA holds is the one before update k; the pricer replays updates k and k+1 on the records.
            // (stage buffer 2 = an explicit request of this pivot's search, which may still sit in the word; 0 / 1 = the one meant here.
            // Pricer 0 may post it before the other pricers' candidates are out: then it was served inside the wait loop above.)
            {
                if (tid == 0) {
                    int4 v = make_int4(0, 0, 0, 0);
                    unsigned spins = 0; long long t0 = 0;
                    for (;;) {
                        v = ld_mail(line + 4);
                        if (v.w == seq && (v.z & 3) != 2) break;
                        if (spin_check(spins, t0, P, 10)) { sh.abort = 1; break; }
                    }
                    sh.ent[4] = v;
                }
                __syncthreads();
                nreq = sh.ent[4];
            }
            if (!sh.abort && nreq.z != ticket) { serve(nreq.x, nreq.y, nreq.z); ticket = nreq.z; }
            PROBE(8);

            Dec D; Pending U;
            const int rcd = gather_decide.template operator()<false>(seq, par, E, nc, D, U);
            if (rcd != 0) { status = rcd; break; }

            // ================================================================ updates
            const bool change = D.change, in_side1 = D.in_side1, longstem = D.longstem;
            const long long delta = D.delta, val = D.val, upper_in = E.upper;
            const int a = D.a, s = D.s, ns = D.ns, dp_uin = D.dp_uin, in_arc = E.arc;
            int4* const stem_g = P.stemseg + (size_t)par * (n + 1) * 2;
            if (!change && delta > 0 && tid == 0 && first >= lo && first < lo + cntn)
                P.flow[in_arc] = (lower_state ? 0 : upper_in) + val;                // NS.cs:1018: stays a non-tree arc, at the other bound
            int bad = 0;
            // ---- cycle nodes: ChangeFlow (NS.cs:1012-1040) and the pred / succ_num part of UpdateTreeStructure (:1042-1183)
            auto update_cycle_node = [&](int j, int x, int sz_u, int pd, int dp, bool hasF) {
                if (delta > 0) {
                    const long long fl = D.new_flow((long long)fl_s[j], pd, hasF, lower_state);
                    bad |= !FT::fits(fl);
                    fl_s[j] = (F)fl;
                }
                if (!change) return;
                if (hasF != in_side1) { sz_s[j] = sz_u + s; return; }               // v_in .. join (NS.cs:1174-1177)
                if (x < a) { sz_s[j] = sz_u - s; return; }                          // v_out .. join (NS.cs:1179-1182)
                // stem node kx (NS.cs:1095-1146): takes over the pred arc of the stem node below it, reversed
                if (x == a) P.flow[pd >> 1] = (D.out.zero & 1) ? 0 : FT::cap_out(up_s[j]);   // u_out: its pred arc leaves the tree at a bound
                const int kx = dp_uin - dp;
                if (kx == 0) {
                    const long long nf = (lower_state ? 0 : upper_in) + val;
                    bad |= !FT::fits(nf);
                    pd_s[j] = in_arc * 2 + (D.dir_new_up ? 1 : 0); sz_s[j] = s;
                    fl_s[j] = (F)nf; up_s[j] = FT::cap_in(upper_in);
                } else {
                    int p_z, p_pd, p_up; long long p_fl;
                    if (!longstem) { p_z = st_z[kx - 1]; p_pd = st_pd[kx - 1]; p_up = st_up[kx - 1]; p_fl = st_fl[kx - 1]; }
                    else {
                        int4 w[2];
                        if (!poll_rec<2>(stem_g + (size_t)(ns - kx) * 2, seq, w, P, 11)) sh.abort = 1;
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
                    if (hasF != hasS) update_cycle_node(j, x, sz_u, pd_s[j], dp_s[j], hasF);
                }
                __syncthreads();                                                    // the relabel pass below rewrites in_s
            }
            // ---- every node: re-label in[] in closed form; re-hung subtree: new depth and pi += sigma (NS.cs:1185-1209)
            PROBE(15);
            if (change) {
                const int b = U.b;
                const int sh_lo = b < a ? b + 1 : a + s, sh_len = b < a ? a - b - 1 : b - a - s + 1, sh_by = b < a ? s : -s;
                // four nodes per 128-bit shared-memory access; padding entries carry label 0, which no update ever moves
                const int nquad = cntn > 0 ? (cntn + 3) >> 2 : 0;
                for (int q4 = tid; q4 < nquad; q4 += kTT) {
                    int4 v = reinterpret_cast<const int4*>(in_s)[q4];
                    int xi[4] = {v.x, v.y, v.z, v.w};
                    bool touched = false, moved = false;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if ((unsigned)(xi[e] - sh_lo) < (unsigned)sh_len) { xi[e] += sh_by; touched = true; }      // between the old and the new place: shift
                        else if ((unsigned)(xi[e] - a) < (unsigned)s) moved = true;                             // re-hung subtree
                    }
                    if (moved) {
#pragma unroll 1
                        for (int e = 0; e < 4; ++e) {
                            const int j = q4 * 4 + e;
                            const int x = e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w;
                            if ((unsigned)(x - a) < (unsigned)s && !((unsigned)(x - sh_lo) < (unsigned)sh_len)) {
                                int nx, nd;
                                relabel(U, x, dp_s[j], nx, nd);
                                xi[e] = nx; dp_s[j] = nd;
                                if (U.sigma != 0) __stcg(P.pi + lo + j, __ldcg(P.pi + lo + j) + U.sigma);   // this CTA is the entry's only reader and writer
                            }
                        }
                        touched = true;
                    }
                    if (touched) reinterpret_cast<int4*>(in_s)[q4] = make_int4(xi[0], xi[1], xi[2], xi[3]);
                }
            }
            if (bad) sh.ovf = 1;                                                    // goes out with CYC(k+1)
            PROBE(13);
            __syncthreads();
            if (sh.abort) [[unlikely]] { status = ST_ERR_BARRIER_TIMEOUT; break; }
            PROBE(14);
            if (P.stop_after > 0 && iterations >= P.stop_after) { status = ST_STOPPED_EARLY; break; }
        }
    }
#undef TICK
#undef PROBE
    if (probe_thr) for (int i = cta == 0 ? 0 : 8; i < (cta == 0 ? 8 : 16); ++i) P.ctl->clk[i] = sh.bk.pr[i];

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
        c->status = status; c->iterations = iterations; c->arcs_checked = bk.arcs_checked; c->final_block_size = sh.mode;
        c->degenerate = bk.degenerate; c->cycle_nodes = bk.cycle_nodes; c->moved_nodes = bk.moved_nodes;
        c->max_cycle = bk.max_cycle; c->max_stem = bk.max_stem; c->pricing_rounds = bk.rounds_total;
        c->ns_price = bk.t_price; c->ns_cycle = bk.t_cycle; c->ns_update = bk.t_update; c->ns_total = gtimer() - bk.t_begin;
        c->clk_total = (unsigned long long)clock64() - bk.c_begin;
        c->ns_wait_done = bk.t_wdone; c->ns_wait_cyc = 0; c->ns_stem = bk.t_stem; c->stem_exchanges = bk.stem_x;
        if (status == ST_ERR_NEEDS_WIDE) c->needs_wide = 1;
    }
}

}  // namespace mcf

// ------------------------------------------------------------------------------------------------ launchers

namespace {
constexpr size_t kStemBytes = (size_t)mcf::kTeamStemCap * (8 + 4 * 4);
constexpr size_t kPricerBytes = (size_t)(mcf::kStageMax + 16) * (8 + 2 * 16 + 4 * 4);   // up, 2 x (pi_s - pi_t, lab), src, tgt, st, cost
inline const void* team_fn(int wide) { return wide ? (const void*)mcf::ns_team_kernel<long long> : (const void*)mcf::ns_team_kernel<int>; }
}  // namespace

extern "C" size_t mcfk_team_smem_bytes(int slice, int wide)
{
    const size_t owner = (size_t)slice * (wide ? mcf::kNodeSmemWide : mcf::kNodeSmemNarrow);
    return (kStemBytes + owner > kPricerBytes ? kStemBytes + owner : kPricerBytes) + 16;
}

// largest slice (nodes per owner CTA) that fits the opt-in shared memory of the device next to the kernel's static part
extern "C" int mcfk_team_max_slice(int device, int wide)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, team_fn(wide)) != cudaSuccess) return -2;
    const long long avail = (long long)prop.sharedMemPerBlockOptin - (long long)fa.sharedSizeBytes - (long long)kStemBytes - 64;
    if (avail + (long long)kStemBytes < (long long)kPricerBytes) return -3;     // the pricing CTA's staging area must fit too
    const long long s = avail / (wide ? mcf::kNodeSmemWide : mcf::kNodeSmemNarrow);
    return (int)(s & ~7LL);
}

// the dynamic shared-memory ceiling of the kernel is always raised to the device's opt-in maximum: several host threads may
// prepare launches with different slice sizes at the same time (mcf_solve_batch_concurrent)
static cudaError_t raise_smem_limit(int device, int wide)
{
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, team_fn(wide));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(team_fn(wide), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(prop.sharedMemPerBlockOptin - fa.sharedSizeBytes));
}

extern "C" int mcfk_team_max_ctas(int device, int slice, int wide)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    const size_t smem = mcfk_team_smem_bytes(slice, wide);
    if (raise_smem_limit(device, wide) != cudaSuccess) return -2;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, team_fn(wide), mcf::kTT, smem) != cudaSuccess) return -3;
    return per_sm * prop.multiProcessorCount;
}

extern "C" void mcfk_team_replicas(int* ent, int* cyc) { *ent = mcf::kRepEnt; *cyc = mcf::kRepCyc; }

extern "C" int mcfk_launch_team(const mcf::TeamParams* p, cudaStream_t stream)
{
    const size_t smem = mcfk_team_smem_bytes(p->slice, p->wide);
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e == cudaSuccess) e = raise_smem_limit(device, p->wide);
    if (e != cudaSuccess) return (int)e;
    void* args[] = {(void*)p};
    e = cudaLaunchCooperativeKernel(team_fn(p->wide), dim3(p->team), dim3(mcf::kTT), args, smem, stream);
    return (int)e;
}
