// mcf_device.cuh - device-side data layout shared by the kernels and the host layer of libmcfgpu.
//
// The reference keeps the spanning tree as parent/pred/thread/rev_thread/succ_num/last_succ linked lists
// (DataStructures/SpanningTree.cs:41-45) and walks them one node at a time.  On a GPU a dependent L2 load
// costs ~130 ns, so a 100-step walk is already slower than a CPU.  This engine therefore replaces the
// *unobservable* part of that structure (thread, rev_thread, last_succ, succ_num) by a nested-interval
// labelling: every node u carries in[u] = its index in a depth-first order of the current basis tree and
// sz[u] = the size of its subtree, so "v is in the subtree of u" is the O(1) test
// in[u] <= in[v] < in[u] + sz[u].  With that test every per-pivot step of NetworkSimplex.cs:925-1209 becomes a
// flat data-parallel pass over the node arrays (cycle discovery, leaving-arc arg-min, re-labelling,
// potential update) instead of a pointer chase.  Only parent / pred / pred_dir / flow / state / pi are
// semantically observable, and they are maintained exactly as the reference maintains them.
#pragma once
#include <stdint.h>

namespace mcf {

constexpr int kThreads = 1024;          // threads per CTA of the persistent pivot kernel
constexpr int kWarps = kThreads / 32;
constexpr int kListSmem = 3584;         // cycle entries staged in shared memory (32 B each = 112 KB); longer cycles are read from global memory
constexpr int kStemCap = 2048;          // longest stem (u_in .. u_out) ranked in shared memory; longer ones go through global scratch

// SpanningTree.cs:53-71
constexpr int STATE_UPPER = -1, STATE_TREE = 0, STATE_LOWER = 1;

// SolverStatus.cs:7-34 (+ engine-internal codes >= 100 that the host turns into error returns)
enum : int {
    ST_NOT_SOLVED = 0, ST_OPTIMAL = 1, ST_INFEASIBLE = 2, ST_UNBOUNDED = 3, ST_UNBALANCED = 4,
    ST_ERR_CYCLE_TOO_LONG = 100, ST_ERR_BARRIER_TIMEOUT = 101, ST_ERR_STEM_TOO_LONG = 102, ST_STOPPED_EARLY = 103,
    ST_ERR_NEEDS_WIDE = 104             // team engine, narrow mode: a tree-arc flow left the int32 range (host re-runs wide)
};

// pricing kinds (PivotRule.cs:7-40; 3 = CachedBlockSearchPivot, NS.cs:1445-1599; 4 = BlockSearchPivotOptimized,
// Internal/BlockSearchPivotOptimized.cs:39-157; 5 / 6 = the Candidate List / Altering List rules PivotRule.cs:33-40 declares and
// NS.cs:884 throws on, defined as LEMON's network_simplex.h:413-518 / :521-635)
enum : int { PK_FIRST = 0, PK_BEST = 1, PK_BLOCK = 2, PK_BLOCK_CACHED = 3, PK_BLOCK_OPT = 4, PK_CAND_LIST = 5, PK_ALT_LIST = 6 };
#ifndef MCF_SORT_CAP
#define MCF_SORT_CAP 4096
#endif
constexpr int kSortCap = MCF_SORT_CAP;          // Altering List: entries sorted per pass in shared memory (16 B each, in the cycle-list staging area)

struct __align__(16) PriceRec {         // one pricing candidate (per CTA, per round)
    long long c;                        // reduced cost (negative when valid, 0 = none)
    int arc, src, tgt, cost;
    int state, off;                     // off = scan offset from next_arc (Block/First)
};

struct __align__(16) CycEnt {           // one tree node of the pivot cycle, captured before any update
    int u, in, sz, pd;                  // pd = pred_arc * 2 + (pred_dir == DIR_UP)
    long long flow, d;                  // flow on pred arc, residual in cycle direction
};

struct Ctl {                            // control block in global memory (zeroed before launch)
    unsigned long long bar;             // monotonically increasing grid-barrier counter
    int abort;                          // set by a barrier time-out
    int list_count[2];                  // cycle-list fill, double buffered by pivot parity
    int status;
    int final_block_size;
    int infeasible;                     // CheckFeasibility, NS.cs:1272-1283
    int needs_wide;                     // team engine, narrow mode: a tree-arc flow left the int32 range
    int pad0;
    long long iterations;
    long long arcs_checked;             // SolverMetrics.TotalArcsChecked
    long long total_cost;               // GetTotalCost, NS.cs:452-465
    long long degenerate;               // pivots with delta == 0
    long long cycle_nodes;              // sum of cycle lengths (tree nodes)
    long long moved_nodes;              // sum of re-hung subtree sizes
    long long max_cycle, max_stem;
    long long pricing_rounds;           // grid-wide pricing rounds (each = one barrier)
    unsigned long long ns_price, ns_cycle, ns_update, ns_total;   // %globaltimer deltas seen by CTA 0
    unsigned long long ns_wait_done, ns_wait_cyc, ns_stem;        // team engine: hop waits of the pricing CTA
    long long stem_exchanges;                                     // team engine: pivots that needed the stem exchange
    unsigned long long clk_total;                                 // team engine: clock64 ticks over the loop (phase accumulators are ticks)
    unsigned long long clk[16];                                   // team engine: sub-phase tick accumulators (0-7 pricing CTA, 8-15 owner CTA 1)
    long long arcs_priced_opt;                                    // flat engine, PK_BLOCK_OPT: arcs priced (the reference keeps no counter there)
};

struct Params {
    int n, m, S, A;                     // nodes, arcs, search arcs (m+n), allocated arcs
    // arcs
    const int* src; const int* tgt; const int* cost;     // [S]
    int* state;                                          // [A]
    long long* flow;                                     // [A]
    const long long* upper;                              // [A]
    const long long* orig_lower;                         // [m] or nullptr (restored into flow at the end)
    long long* rc_cache;                                 // [S] or nullptr (PK_BLOCK_CACHED)
    // nodes, [n+1] (root = n)
    int* in; int* sz; int* parent; int* pd;
    long long* pi;
    // work areas
    PriceRec* part;                     // [2][gridDim.x]
    CycEnt* list; int list_cap;
    int* stem_scratch;                  // [6][n+1] flat engine: stems longer than kStemCap are ranked and read here
    // Candidate List / Altering List rules: the list (double buffered), reduced costs / scratch of the list pass (CTA 0 only)
    int* cand;                          // [2][cand_cap]
    long long* cand_cost;               // [2][cand_cap]
    int* cand_scratch;                  // [2][cand_cap]
    int cand_cap;
    int list_length, minor_limit;       // Candidate List (network_simplex.h:441-458)
    int head_length;                    // Altering List (:563-580); its block size is block_size
    Ctl* ctl;
    // pricing configuration (BlockSearchPivot ctor, NS.cs:1304-1337; adaptive rule :1399-1438)
    int kind;
    int block_size, dyn_min_block, max_block_size, adaptive, consecutive;
    double low_thr, high_thr, shrink, grow;
    int lookahead0;                     // groups priced in the first round of a search
    int simd_width;                     // PK_BLOCK_OPT: Vector<long>.Count of the host the reference would run on (0 = scalar path)
    long long max_iterations;           // NS.cs:280
    long long stop_after;               // >0: stop after this many pivots (bounded samples / tests)
    unsigned long long barrier_timeout_cycles;
};

// ------------------------------------------------------------------------------------------------ team engine
// (mcf_team.cu) node slices resident in shared memory, CTAs exchange 16-byte self-validating words through L2.

struct __align__(16) NodeRec {          // global mirror of a node, read by the pricing gathers (one 128-bit load)
    long long pi;                       // potential (NS.cs:48)
    int in;                             // initial depth-first index (the live one is the dense array TeamParams::in_g: a third of
                                        // all labels shift on every pivot, and 4-byte-stride stores cost a quarter of the sectors)
    int dp;                             // current depth of the node (root = 0)
};

constexpr int kMailWords = 8;           // 16-byte words per mailbox record (one 128-byte line)
constexpr int kTeamMax = 160;           // upper bound on CTAs in a team (>= SM count)
constexpr int kMaxPricers = 32;         // pricing CTAs of a team
#ifndef MCF_STEM_CAP
#define MCF_STEM_CAP 1024
#endif
#ifndef MCF_SPILL_DP
#define MCF_SPILL_DP 0
#endif
constexpr int kTeamStemCap = MCF_STEM_CAP;   // longest stem the team engine stages in shared memory (longer ones are read in place)

// per resident node: in, sz, pd, depth (int) + flow and capacity of its pred arc (int32 in narrow mode, int64 in wide mode)
constexpr int kNodeSmemNarrow = 24, kNodeSmemWide = 32, kNodeSmemSpill = MCF_SPILL_DP ? 12 : 16;   // spill: in, sz, pd (+ depth)

struct TeamParams {
    int n, m, S, A;
    const int* src; const int* tgt; const int* cost;     // [S]
    int* state;                                          // [A]
    long long* flow;                                     // [A]
    const long long* upper;                              // [A]
    const long long* orig_lower;                         // [m] or nullptr
    NodeRec* node;                                       // [n+1] (root = n): pi and depth
    int* in_g;                                           // [n+1] current depth-first index of every node
    const int* sz0; const int* pd0;                      // [n+1] initial basis
    long long* pi_out;                                   // [n]
    int4* ent0;                                          // [2][pricers][kMailWords]  pricer -> all: its part of the first block
    int4* late;                                          // [2][kMailWords]           pricer 0 -> all: result of a multi-block search
    int4* prc;                                           // [2][pricers][kMailWords]  pricer -> pricers: later rounds of a search
    int4* cyc;                                           // [2][team][kMailWords]     owner -> all
    int4* stemseg;                                       // [2][n+1][2]               stem entries, owner o at its slice offset
    unsigned int* done;                                  // [team + pricers][32]      owner -> pricers (DONE), pricer -> owners (GATHERED)
    Ctl* ctl;
    int team;                                            // CTAs: [0, pricers) price, [pricers, team) own node slices
    int pricers;
    int slice;                                           // nodes per owner
    int wide;                                            // 1: tree-arc flows / capacities resident as int64, 0: int32
    int spill;                                           // 1: flows / capacities of the tree arcs live in global memory (fl_g / up_g, touched for the
                                                         // ~30 cycle nodes of a pivot only) and a slice keeps 16 B per node: instances past the
                                                         // resident capacity (n > 1.05 M, or 2^20 nodes with 64-bit flows) still run on this engine
    void* fl_g; void* up_g;                              // [owners x slice] of int (narrow) or long long (wide); entry of node u at index u
    int block_size, dyn_min_block, max_block_size, adaptive, consecutive;
    double low_thr, high_thr, shrink, grow;
    long long max_iterations, stop_after;
    unsigned long long timeout_cycles;
};

// SolutionValidator on the device (mcf_kernels.cu)
struct ValidateParams {
    int n, m, supply_type;
    const int* src; const int* tgt; const int* cost;
    const long long* flow; const long long* lower; const long long* upper; const long long* supply; const long long* pi;
    long long* net; long long* adj; long long* out;
};

}  // namespace mcf
