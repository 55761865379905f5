/*
 * ns_oracle.h - interface of the CPU restatement of the reference solver.
 * TEST INFRASTRUCTURE ONLY (see ns_oracle.c).  Enum values are the reference's:
 *   PivotRule.cs:7-40, SolverStatus.cs:7-34, SupplyType.cs:7-17, OptimizationTypes.cs:8-20.
 */
#ifndef NS_ORACLE_H
#define NS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { NS_PIVOT_FIRST_ELIGIBLE = 0, NS_PIVOT_BEST_ELIGIBLE = 1, NS_PIVOT_BLOCK_SEARCH = 2, NS_PIVOT_CANDIDATE_LIST = 3, NS_PIVOT_ALTERING_LIST = 4 };
enum { NS_STATUS_NOT_SOLVED = 0, NS_STATUS_OPTIMAL = 1, NS_STATUS_INFEASIBLE = 2, NS_STATUS_UNBOUNDED = 3, NS_STATUS_UNBALANCED = 4 };
enum { NS_SUPPLY_GEQ = 0, NS_SUPPLY_LEQ = 1 };
enum {
    NS_FLAG_ADAPTIVE_BLOCK_SIZE = 1, NS_FLAG_SMALL_BLOCKS_FOR_DENSE = 2, NS_FLAG_REDUCED_COST_CACHING = 4,
    NS_FLAG_CANDIDATE_LIST_PIVOT = 8, NS_FLAG_HOT_COLD_SPLITTING = 16, NS_FLAG_EARLY_TERMINATION = 32
};
enum { NS_TYPE_GENERAL = 0, NS_TYPE_CIRCULATION = 1, NS_TYPE_ASSIGNMENT = 2, NS_TYPE_TRANSPORTATION = 3, NS_TYPE_TRANSSHIPMENT = 4, NS_TYPE_TIME_EXPANDED = 5 };  /* ProblemCharacteristics.cs:151-182 */

typedef struct {                 /* OptimizationConfig, OptimizationTypes.cs:25-38 */
    int32_t flags, max_block_size, min_block_size, dense_network_threshold, consecutive_hits_before_adapt, _pad;
    double candidate_list_ratio, block_size_growth_factor, block_size_shrink_factor;
    double low_hit_rate_threshold, high_hit_rate_threshold, min_block_size_ratio;
} ns_oracle_config;

typedef struct {                 /* ProblemCharacteristics.cs */
    int32_t node_count, arc_count, max_degree, source_count, sink_count, transshipment_count;
    int32_t detected_type, is_dense, is_sparse, is_layered, has_uniform_costs, has_uniform_capacities;
    double density, average_degree, degree_variance, degree_cv, cost_variance, average_cost, cost_cv;
    double average_capacity, finite_capacity_ratio;
    int64_t cost_range, capacity_range, total_supply, max_absolute_supply;
} ns_oracle_characteristics;

/* Solver state at a pivot boundary (test infrastructure: lets the CPU baseline time a window in the MIDDLE of a long solve
 * by resuming from a checkpoint this same code wrote).  All arrays are caller-owned: [n+1] node arrays, [m+2n] arc arrays. */
typedef struct {
    int32_t *parent, *pred, *thread, *rev_thread, *succ_num, *last_succ;
    int8_t *pred_dir, *state;
    int64_t *flow, *pi;
    int64_t iterations;
    int32_t next_arc, block_size, consecutive_low, consecutive_high;
} ns_oracle_state;

typedef struct {
    int32_t supply_type;         /* NS_SUPPLY_*  (NS.cs:38 default Geq) */
    int32_t pivot_rule;          /* NS_PIVOT_*   (NS.cs:77 default BlockSearch) */
    int32_t optimized_pivot;     /* EnableOptimizedPivot, NS.cs:532 */
    int32_t auto_config;         /* _useAutoConfiguration, NS.cs:90 (default 1) */
    int32_t simd_width;          /* Vector<long>.Count seen by BlockSearchPivotOptimized.cs:74 (4 = AVX2, 0 = none) */
    int32_t collect_phase_times; /* per-pivot stopwatches like NS.cs:285-339 (slows the loop) */
    int64_t max_pivots;          /* >0: stop after this many pivots (bounded CPU-baseline sample) */
    int64_t trace_capacity;      /* entries available in the trace buffers */
    int32_t *trace_in_arc;       /* entering arc per pivot */
    int32_t *trace_u_out;        /* leaving node per pivot, -1 when the entering arc only flips bound */
    ns_oracle_config config;     /* used when auto_config == 0 (SetOptimizationConfig, NS.cs:557-561) */
    const ns_oracle_state *resume; /* not NULL: continue from this state (its iterations count on; max_pivots is absolute) */
    ns_oracle_state *save;       /* not NULL: filled with the state when the loop ends (at max_pivots, or at optimality - before the
                                    lower bounds are added back to the flows) */
    int32_t emulate_stackalloc;  /* 1: zero-fill an int[n] scratch on every stem re-hang like the reference's
                                    `stackalloc int[_nodeCount]` (NS.cs:1085; no SkipLocalsInit) - timing fidelity only.
                                    Default 0 = the scratch is hoisted (what a C/C++ port would do; conservative baseline). */
    int32_t _pad2;
    const ns_oracle_state *warm; /* not NULL: WARM START from the basis a previous Optimal solve of the same network with
                                    other arc costs saved: tree, arc states and flows are taken over, the potentials are recomputed
                                    along the tree for the current costs, pivot-rule state and counters start fresh */
} ns_oracle_options;

typedef struct {
    int32_t status, pivot_kind, initial_block_size, final_block_size, stopped_early, _pad;
    int64_t iterations, total_arcs_checked, degenerate_pivots;
    int64_t join_steps, max_join_steps, stem_nodes, subtree_nodes, max_subtree_nodes;
    int64_t total_cost, art_cost, sum_supply;
    double total_seconds, loop_seconds, pricing_seconds, tree_seconds, potential_seconds;
    ns_oracle_config config_used;
    ns_oracle_characteristics characteristics;
} ns_oracle_result;

void ns_oracle_default_config(ns_oracle_config *c);
void ns_oracle_analyze(int n, int m, const int32_t *src, const int32_t *tgt, const int64_t *lower,
                       const int64_t *upper, const int64_t *cost, const int64_t *supply,
                       ns_oracle_characteristics *ch);
void ns_oracle_select_config(const ns_oracle_characteristics *ch, ns_oracle_config *cfg);
int ns_oracle_solve(int n, int m, const int32_t *src, const int32_t *tgt, const int64_t *lower,
                    const int64_t *upper, const int64_t *cost, const int64_t *supply,
                    const ns_oracle_options *opt, ns_oracle_result *res, int64_t *flow_out, int64_t *pi_out);
int ns_oracle_validate(int n, int m, const int32_t *src, const int32_t *tgt, const int64_t *lower,
                       const int64_t *upper, const int64_t *cost, const int64_t *supply, int supply_type,
                       const int64_t *flow, const int64_t *pi, int64_t reported_cost, int64_t *dual_cost_out);

#ifdef __cplusplus
}
#endif
#endif
