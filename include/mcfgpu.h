/*
 * mcfgpu.h - C ABI of libmcfgpu.so, the B200 (sm_100a) network-simplex engine.
 *
 * The reference (pdegenhardt/MinCostFlow, C#) has no native boundary at all (no DllImport anywhere); this
 * header DEFINES the boundary at the narrowest seam the reference has: the public solver surface
 *   IMinCostFlowSolver            src/MinCostFlow.Core/IMinCostFlowSolver.cs:8-34
 *   NetworkSimplex (the extras)   src/MinCostFlow.Core/Lemon/Algorithms/NetworkSimplex.cs:119-210, :416-587
 * and follows the pinned-raw-pointer precedent of its private OptimizedPivotWrapper (NetworkSimplex.cs:1699-1722).
 * Every entry point below names the reference member it replaces.  INTEGRATION.md shows the P/Invoke
 * declarations a maintainer adds on the C# side (`CudaNetworkSimplex : IMinCostFlowSolver`).
 *
 * Conventions
 *   - plain pointers and sizes only; all structs are blittable (sequential layout, no pointers inside).
 *   - every function returns 0 (MCF_OK) or a negative mcf_error; no exception crosses the ABI.
 *     The C# / C++ / Python wrappers map codes back to the reference's exceptions:
 *       MCF_ERR_INVALID_ARGUMENT -> ArgumentException          (NetworkSimplex.cs:155-158, :171-174, :185-188)
 *       MCF_ERR_NOT_OPTIMAL      -> InvalidOperationException("Solution not optimal")  (NetworkSimplex.cs:418-421)
 *   - inputs are caller-owned and copied during the call; the handle owns all device and host staging memory;
 *     results are copied device->host once per solve and served from host memory afterwards
 *     (SolutionValidator calls GetFlow/GetPotential per element, SolutionValidator.cs:59-78).
 *   - a handle is not thread-safe (neither is the reference solver); different handles may be used from
 *     different threads.  There is NO CPU fallback: without an sm_100 device mcf_create fails with
 *     MCF_ERR_NO_DEVICE.
 *   - arithmetic: arc ids/endpoints int32; costs must fit int32 on the device (|cost| <= 2^31-1, else
 *     MCF_ERR_RANGE); flows, capacities, supplies, potentials and reduced costs are int64 as in the
 *     reference (`long`, NetworkSimplex.cs:42-48).  "Infinite" capacity is INT64_MAX/2 (NetworkSimplex.cs:127).
 */
#ifndef MCFGPU_H
#define MCFGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCF_API_VERSION 1
#define MCF_INF (INT64_MAX / 2)               /* NetworkSimplex.cs:127 */

typedef enum mcf_error {
    MCF_OK = 0,
    MCF_ERR_INVALID_ARGUMENT = -1,            /* bad handle / null pointer / arc or node id out of range */
    MCF_ERR_NO_DEVICE = -2,                   /* no CUDA device of compute capability 10.x */
    MCF_ERR_CUDA = -3,                        /* a CUDA runtime call failed; see mcf_last_error */
    MCF_ERR_OUT_OF_MEMORY = -4,
    MCF_ERR_NOT_OPTIMAL = -5,                 /* result getter called while Status != Optimal */
    MCF_ERR_RANGE = -6,                       /* |cost| does not fit int32, or too many arcs (m + 2n >= 2^30) */
    MCF_ERR_ENGINE_LIMIT = -7,                /* pivot cycle / stem longer than the in-kernel staging buffers */
    MCF_ERR_TIMEOUT = -8,                     /* a grid barrier timed out (kernel abandoned) */
    MCF_ERR_NOT_SOLVED = -9,                  /* metrics / results requested before mcf_solve */
    MCF_ERR_FORMAT = -10,                     /* malformed DIMACS text (FormatException in DimacsReader.cs); see mcf_io_last_error */
    MCF_ERR_IO = -11                          /* file cannot be read / written */
} mcf_error;

/* Enum values are the reference's: SolverStatus.cs:7-34, PivotRule.cs:7-40, SupplyType.cs:7-17,
 * OptimizationFlags in OptimizationTypes.cs:8-20. */
typedef enum mcf_status { MCF_NOT_SOLVED = 0, MCF_OPTIMAL = 1, MCF_INFEASIBLE = 2, MCF_UNBOUNDED = 3, MCF_UNBALANCED = 4 } mcf_status;
typedef enum mcf_pivot_rule {       /* PivotRule.cs:7-40; 3 and 4 are declared there and thrown on at NetworkSimplex.cs:884 - here they run, defined as
                                       LEMON's CandidateListPivotRule / AlteringListPivotRule (lemon-1.3.1/lemon/network_simplex.h:413-635) */
    MCF_FIRST_ELIGIBLE = 0, MCF_BEST_ELIGIBLE = 1, MCF_BLOCK_SEARCH = 2, MCF_CANDIDATE_LIST = 3, MCF_ALTERING_LIST = 4
} mcf_pivot_rule;
typedef enum mcf_supply_type { MCF_GEQ = 0, MCF_LEQ = 1 } mcf_supply_type;
enum {
    MCF_FLAG_ADAPTIVE_BLOCK_SIZE = 1, MCF_FLAG_SMALL_BLOCKS_FOR_DENSE = 2, MCF_FLAG_REDUCED_COST_CACHING = 4,
    MCF_FLAG_CANDIDATE_LIST_PIVOT = 8, MCF_FLAG_HOT_COLD_SPLITTING = 16, MCF_FLAG_EARLY_TERMINATION = 32
};

/* OptimizationConfig (OptimizationTypes.cs:25-38), same defaults via mcf_default_options. */
typedef struct mcf_optimization_config {
    int32_t flags;
    int32_t max_block_size;
    int32_t min_block_size;
    int32_t dense_network_threshold;
    int32_t consecutive_hits_before_adapt;
    int32_t reserved0;
    double candidate_list_ratio;
    double block_size_growth_factor;
    double block_size_shrink_factor;
    double low_hit_rate_threshold;
    double high_hit_rate_threshold;
    double min_block_size_ratio;
} mcf_optimization_config;

typedef struct mcf_options {
    int32_t supply_type;            /* SetSupplyType, NetworkSimplex.cs:197; default MCF_GEQ (:38) */
    int32_t pivot_rule;             /* SetPivotRule, NetworkSimplex.cs:206; default MCF_BLOCK_SEARCH (:77) */
    int32_t auto_configuration;     /* SetAutoConfiguration, NetworkSimplex.cs:567; default 1 (:90) */
    int32_t optimized_pivot;        /* EnableOptimizedPivot, NetworkSimplex.cs:532: the rules of Internal/BlockSearchPivotOptimized.cs.
                                       First/Best Eligible: same pivots as the managed rules.  Block Search: its own cursor
                                       and wrap rules (:39-110), see simd_width */
    int32_t device;                 /* CUDA device ordinal */
    int32_t max_ctas;               /* 0 = one CTA per SM (cooperative-launch limit) */
    int32_t lookahead_blocks;       /* flat engine: blocks priced in the first pricing round (0 = 2); team engine: pricing CTAs (0 = auto) */
    int32_t engine;                 /* 0 = automatic; 1 = flat engine (mcf_kernels.cu); 2 = team engine (mcf_team.cu, Block Search); 3 = team engine
                                       with the tree-arc flows in global memory (what 0 / 2 fall back to when the resident slices do not fit) */
    int32_t simd_width;             /* optimized Block Search only: Vector<long>.Count of the host whose pivot sequence is to be
                                       reproduced (BlockSearchPivotOptimized.cs:74, :119): 4 = x64 AVX2 (default), 2 = SSE2 / NEON,
                                       0 = Vector.IsHardwareAccelerated false.  With a non-zero width the reference's scalar loop
                                       resumes after the vector loop's early return with its block counter at 0 and scans the
                                       rest of the range; the engine reproduces exactly that */
    int32_t warm_start;             /* SURVEY.md 8f-3 (README.md:17-18 "warm start" roadmap item; LEMON re-run semantics network_simplex.h:836-884):
                                       1 = when the previous mcf_solve on this handle ended Optimal and only arc COSTS changed since (same
                                       topology, bounds, supplies, supply type), start from that solve's optimal basis - tree, arc states and
                                       flows kept, potentials recomputed for the new costs - instead of the artificial star basis.  Anything
                                       else falls back to a cold start; mcf_metrics.warm_started says which one ran.  0 (default) = always cold */
    int64_t stop_after_pivots;      /* >0: stop after this many pivots with Status = NotSolved (bounded samples) */
    double barrier_timeout_s;       /* 0 = default 10 s */
    mcf_optimization_config config; /* SetOptimizationConfig, NetworkSimplex.cs:557-561 (used when auto_configuration == 0) */
} mcf_options;

/* SolverMetrics (OptimizationTypes.cs:43-69) + engine counters. */
typedef struct mcf_metrics {
    int64_t iterations;                 /* Iterations */
    int64_t total_arcs_checked;         /* TotalArcsChecked (Block Search rules only, like the reference) */
    int32_t initial_block_size;         /* InitialBlockSize */
    int32_t final_block_size;           /* FinalBlockSize */
    int32_t baseline_iterations;        /* BaselineIterations = (int)(sqrt(S) * n * 0.5), NetworkSimplex.cs:276 */
    int32_t pricing_kind;               /* 0 First, 1 Best, 2 Block, 3 cached Block (NetworkSimplex.cs:879-885), 4 optimized Block (:851-856) */
    double average_arcs_checked_per_pivot;
    double iteration_ratio;
    double pivot_search_time_us;        /* PivotSearchTimeMicros  (device time in phase A) */
    double tree_update_time_us;         /* TreeUpdateTimeMicros + PotentialUpdateTimeMicros: phase C is fused */
    double cycle_time_us;               /* join + leaving-arc search (phase B); the reference leaves this untimed */
    double total_solve_time_us;         /* TotalSolveTimeMicros: host wall time of mcf_solve */
    double kernel_time_us;              /* device time of the persistent kernel (CUDA events) */
    double h2d_time_us, d2h_time_us, host_prepass_time_us;
    int64_t h2d_bytes, d2h_bytes;
    int64_t arcs_priced;                /* arcs whose reduced cost was evaluated on the device */
    int64_t pricing_bytes;              /* 16 B * arcs_priced (int32 source, target, cost, state) */
    int64_t degenerate_pivots, cycle_nodes, moved_nodes, max_cycle, max_stem, pricing_rounds;
    int32_t config_flags;               /* OptimizationFlags actually used (after auto-configuration) */
    int32_t grid_ctas;
    double degree_cv;                   /* ProblemCharacteristics.DegreeCV when auto-configured, else 0 */
    int32_t engine;                     /* 1 = flat engine, 2 = team engine */
    int32_t pricer_ctas;                /* team engine: CTAs that price (the rest of grid_ctas own node slices) */
    int64_t stem_exchanges;             /* team engine: pivots whose re-hung stem was longer than one node */
    double hop_wait_done_us;            /* team engine: pricing CTA waiting for the owners' updates to become visible */
    double stem_exchange_us;
    double ns_per_clock;                /* team engine: measured SM clock period */
    double phase_us[16];                /* team engine: sub-phase times (0-7 pricing CTA, 8-15 first owner CTA), see DESIGN.md */
    int32_t wide_flows;                 /* team engine: bit 0 = tree-arc flows kept as int64 (else int32); bit 1 = kept in global memory (else in the slices) */
    int32_t warm_started;               /* 1 = this solve started from the previous optimal basis (mcf_options.warm_start) */
} mcf_metrics;

typedef struct mcf_handle mcf_handle;

int mcf_api_version(void);
/* number of usable sm_100 devices (0 when none) */
int mcf_device_count(void);
void mcf_default_options(mcf_options* out);

/* new NetworkSimplex(graph): NetworkSimplex.cs:119-148 + InitializeGraphStructure :605-622.
 * source/target: arc endpoints in arc-id order (CompactDigraph.cs:110-126). */
int mcf_create(int32_t n, int32_t m, const int32_t* source, const int32_t* target, mcf_handle** out);
void mcf_destroy(mcf_handle* h);

/* SetArcBounds / SetArcCost for all arcs at once (NetworkSimplex.cs:153-178).  NULL keeps the defaults
 * lower = 0, upper = MCF_INF, cost = 0 (NetworkSimplex.cs:615-617). */
int mcf_set_arcs(mcf_handle* h, const int64_t* lower, const int64_t* upper, const int64_t* cost);
/* SetNodeSupply for all nodes (NetworkSimplex.cs:183-192). */
int mcf_set_supply(mcf_handle* h, const int64_t* supply);
/* SetSupplyType / SetPivotRule / EnableOptimizedPivot / SetOptimizationConfig / SetAutoConfiguration. */
int mcf_set_options(mcf_handle* h, const mcf_options* opt);

/* Solve(): NetworkSimplex.cs:215-411.  *status_out receives a mcf_status. */
int mcf_solve(mcf_handle* h, int32_t* status_out);
int mcf_get_status(mcf_handle* h, int32_t* status_out);           /* Status, NetworkSimplex.cs:470 */

/* GetFlow / GetPotential / GetTotalCost (NetworkSimplex.cs:416-465): MCF_ERR_NOT_OPTIMAL unless Optimal. */
int mcf_get_flows(mcf_handle* h, int64_t* out_m);
int mcf_get_potentials(mcf_handle* h, int64_t* out_n);
int mcf_get_flow(mcf_handle* h, int32_t arc, int64_t* out);
int mcf_get_potential(mcf_handle* h, int32_t node, int64_t* out);
int mcf_get_total_cost(mcf_handle* h, int64_t* out);
/* GetNodeSupply / GetArcCost / GetArcLowerBound / GetArcUpperBound (NetworkSimplex.cs:480-527), including the
 * reference's post-solve quirk that the upper bound is returned shifted by the lower bound (:647-651). */
int mcf_get_node_supply(mcf_handle* h, int32_t node, int64_t* out);
int mcf_get_arc_cost(mcf_handle* h, int32_t arc, int64_t* out);
int mcf_get_arc_lower_bound(mcf_handle* h, int32_t arc, int64_t* out);
int mcf_get_arc_upper_bound(mcf_handle* h, int32_t arc, int64_t* out);

int mcf_get_metrics(mcf_handle* h, mcf_metrics* out);             /* GetMetrics, NetworkSimplex.cs:584 */

/* Bounded solves (mcf_options.stop_after_pivots > 0: samples, tests, warm-up): the arc flows (before the lower-bound restore of
 * NetworkSimplex.cs:375-388) and node potentials of the basis the solve stopped at, i.e. the solver state after exactly that many
 * pivots of NetworkSimplex.cs:282-341.  MCF_ERR_NOT_SOLVED unless the last solve ended that way.  Pointers may be NULL. */
int mcf_get_state_after_stop(mcf_handle* h, int64_t* flow_out_m, int64_t* potential_out_n);

/* The result arrays of the last Optimal solve where the solve left them in HBM (flow[m] in arc-id order, potential[n] in
 * node-id order, both int64) and the device they live on: what GetFlow / GetPotential serve from the host copies.  For the
 * multi-GPU result gather of a batch, which sends the records GPU -> GPU over NCCL without a host bounce.
 * Valid until the next call that solves, probes or destroys the handle.  MCF_ERR_NOT_OPTIMAL unless Optimal. */
int mcf_get_device_results(mcf_handle* h, const int64_t** flow_dev_out, const int64_t** potential_dev_out, int32_t* device_out);

/* Batches of independent instances (BASELINE.json config 5): instance i is solved on devices[i % n_devices],
 * one host thread per device.  statuses_out[count] receives each instance's mcf_status. */
int mcf_solve_batch(mcf_handle** hs, int32_t count, const int32_t* devices, int32_t n_devices, int32_t* statuses_out);
/* The same with `per_device` solves side by side on every GPU, each on its own stream over a 1/per_device share of the SMs
 * (instance i runs on devices[i % n_devices]).  Worth it when an instance does not need the whole GPU: a 2^18-node
 * instance occupies 37 of 148 SMs. */
int mcf_solve_batch_concurrent(mcf_handle** hs, int32_t count, const int32_t* devices, int32_t n_devices, int32_t per_device,
                               int32_t* statuses_out);

/* Roofline probe: uploads the handle's initial basis and runs the stand-alone Best Eligible pricing sweep
 * (one full pass over all S = m + n arcs, 16 B of arc data per arc) `reps` times, timing each launch with CUDA
 * events on the handle's stream.  ms_out[reps] receives the per-launch times, *entering_arc_out the arg-min arc
 * (lowest id among ties, -1 if none).  flush_l2 != 0: a 256 MB buffer (L2 is 126 MB) is overwritten between launches; with
 * flush_l2 == 1 a second 256 MB buffer is then read, so that the write-back of the first one's dirty lines does not fall into the
 * timed launch (2 = overwrite only). */
int mcf_pricing_probe(mcf_handle* h, int32_t reps, int32_t flush_l2, float* ms_out, int32_t* entering_arc_out,
                      int64_t* arcs_per_launch_out);

/* SolutionValidator.Validate() (Lemon/Validation/SolutionValidator.cs:20-53) on the device, over the arrays the solve left
 * in HBM: *failed_checks_out is a bit set - 1 flow conservation (:55-100), 2 capacity bounds (:102-125), 4 complementary
 * slackness (:135-176), 8 dual feasibility of the supply form (:191-231), 16 objective != sum flow*cost (:234-262),
 * 32 dual objective != primal (:268-342); 0 = IsValid.  MCF_ERR_NOT_OPTIMAL unless Status == Optimal (:24-33). */
int mcf_validate(mcf_handle* h, int32_t* failed_checks_out, int64_t* primal_objective_out, int64_t* dual_objective_out);

/* ---- DIMACS bulk I/O: the step before and after the path (host code, mcf_io.cpp) -------------------------------------
 * DimacsReader.ReadFromStream (src/MinCostFlow.Problems/Loaders/DimacsReader.cs:36-147) as one parallel pass over the
 * text into flat arrays in arc-id order (= order of the `a` lines), replacing the reader + the per-element setter loop
 * of Benchmarks/NetworkSimplexBenchmarks.cs:166-189.  Same grammar and errors: `p min N M` (:67-81), `n ID SUPPLY`
 * (:83-93, 1-based ids, a later line overrides), `a FROM TO LOWER UPPER COST` (:95-109), `c` and unknown lines skipped,
 * wrong token counts / non-numbers -> MCF_ERR_FORMAT with the reference's message.  Deviation: an arc count different
 * from the `p` line is refused (the reference throws IndexOutOfRange or leaves arrays and graph of different sizes). */
typedef struct mcf_dimacs mcf_dimacs;
int mcf_dimacs_open(const char* path, mcf_dimacs** out);
int mcf_dimacs_parse(const char* text, int64_t length, mcf_dimacs** out);
int mcf_dimacs_dims(const mcf_dimacs* d, int32_t* n_out, int32_t* m_out);
/* any pointer may be NULL; source/target/lower/upper/cost hold m entries, supply n */
int mcf_dimacs_copy(const mcf_dimacs* d, int32_t* source, int32_t* target, int64_t* lower, int64_t* upper, int64_t* cost, int64_t* supply);
void mcf_dimacs_close(mcf_dimacs* d);
/* DimacsReader.ReadFromFile + new NetworkSimplex(graph) + SetArcBounds / SetArcCost / SetNodeSupply for every element */
int mcf_create_from_dimacs(const char* path, mcf_handle** out);
/* SolutionLoader.SaveToFile (Loaders/SolutionLoader.cs:186-214): `s COST`, one `f` line per non-zero flow, optional
 * `p NODE POTENTIAL` lines.  format 0: `f ARC_ID FLOW` (0-based arc id, what SaveToFile writes); format 1: `f SRC DST FLOW`
 * (1-based node ids, the form of the .sol fixtures under Resources/, :124-129).  MCF_ERR_NOT_OPTIMAL unless Optimal. */
int mcf_write_solution(mcf_handle* h, const char* path, int32_t format, int32_t with_potentials);
/* SolutionLoader.LoadFromStream (:69-173): objective (INT64_MIN = "not specified", :165-170) and the `f` lines: up to
 * `capacity` of them into a_out / b_out / flow_out (arc id and -1, or 0-based source and target); *flow_lines_out = how many
 * the file holds; *endpoint_form_out = 1 if any line had the SRC DST form.  Output pointers may be NULL. */
int mcf_read_solution(const char* path, int64_t* cost_out, int32_t capacity, int32_t* a_out, int32_t* b_out, int64_t* flow_out,
                      int32_t* flow_lines_out, int32_t* endpoint_form_out);
/* message of the last failed I/O call on this thread */
const char* mcf_io_last_error(void);

/* number of nodes / arcs and the arc endpoints the handle was created with */
int mcf_get_dims(mcf_handle* h, int32_t* n_out, int32_t* m_out);
int mcf_get_endpoints(mcf_handle* h, int32_t* source_out, int32_t* target_out);

const char* mcf_last_error(mcf_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MCFGPU_H */
