/*
 * ns_oracle.c - CPU restatement of the reference's primal network simplex.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mincostflow_b200/ may include, link
 * or call this file; it exists so tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs can check and time the GPU
 * engine against the reference algorithm.
 *
 * It restates, function by function, the reference C# solver
 *   NS.cs = /root/reference/src/MinCostFlow.Core/Lemon/Algorithms/NetworkSimplex.cs
 * (which cannot be executed here: no dotnet/mono in the image) with the same
 * array layout (thread / rev_thread / succ_num / last_succ spanning tree), the
 * same scan orders and the same tie-breaks, so that the *pivot sequence* - and
 * therefore every arc flow and node potential - is the reference's.
 *
 * Parity pinning (tests/test_oracle_golden.py): the reference's published pivot
 * counts 96258 / 124916 / 142905 / 144041 on circulation_1000_0_05
 * (docs/performance-optimization-final-results.md:50-53), every fixture .sol
 * objective, the flow vectors of NetworkSimplexTests.cs, and cost agreement with
 * the vendored LEMON 1.3.1 build in oracle/_ref.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ns_oracle.h"

/* SpanningTree.cs:53-71 */
#define STATE_UPPER (-1)
#define STATE_TREE 0
#define STATE_LOWER 1
#define DIR_DOWN (-1)
#define DIR_UP 1

#define I64_MAX INT64_MAX
#define NS_INF (INT64_MAX / 2)          /* NS.cs:126-127 */

typedef struct {
    int n, m;                    /* _nodeCount, _arcCount */
    int all_arc_num, search_arc_num, root;
    int64_t *lower, *upper, *cost, *supply, *flow, *pi, *orig_lower;
    int *source, *target;
    int *parent, *pred, *thread, *rev_thread, *succ_num, *last_succ;
    int8_t *pred_dir, *state;
    int *dirty_revs;             /* NS.cs:1085 stackalloc, hoisted */
    int64_t sum_supply, art_cost;
    int in_arc, join, u_in, v_in, u_out, v_out;
    int64_t delta;
    /* reduced-cost cache (NS.cs:63-74); the dirty-node machinery is unreachable, see find_cached */
    int64_t *reduced_costs; int reduced_costs_dirty;
    ns_oracle_config cfg;
    /* pivot rule state */
    int block_size, next_arc, consecutive_low, consecutive_high, dyn_min_block;
    int64_t arcs_checked_pivot;
    /* Candidate List / Altering List rules (LEMON network_simplex.h:415-635; the C# port declares them, PivotRule.cs:33-40,
     * and throws NotImplementedException, NS.cs:884) */
    int *candidates; int64_t *cand_cost;
    int list_length, minor_limit, curr_length, minor_count, head_length, alt_block;
    ns_oracle_result *res;
    const ns_oracle_options *opt;
} ns_t;

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------------------------------- */
/* ProblemAnalyzer.cs:21-62 + OptimizationSelector.cs:14-96                   */

void ns_oracle_default_config(ns_oracle_config *c)
{   /* OptimizationTypes.cs:25-38 */
    c->flags = 0; c->max_block_size = 100; c->min_block_size = 25; c->dense_network_threshold = 10000;
    c->candidate_list_ratio = 0.1; c->block_size_growth_factor = 1.2; c->block_size_shrink_factor = 0.8;
    c->low_hit_rate_threshold = 0.05; c->high_hit_rate_threshold = 0.3; c->consecutive_hits_before_adapt = 3;
    c->min_block_size_ratio = 0.125;
}

void ns_oracle_analyze(int n, int m, const int32_t *src, const int32_t *tgt, const int64_t *lower,
                       const int64_t *upper, const int64_t *cost, const int64_t *supply,
                       ns_oracle_characteristics *ch)
{
    memset(ch, 0, sizeof(*ch));
    ch->node_count = n; ch->arc_count = m;
    int64_t max_possible = (int64_t)n * (n - 1);                         /* ProblemAnalyzer.cs:35-36 */
    ch->density = max_possible > 0 ? (double)m / (double)max_possible : 0;

    /* AnalyzeNodeDegrees, ProblemAnalyzer.cs:64-100 */
    int *deg = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    int *outd = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    int *ind = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int e = 0; e < m; e++) { outd[src[e]]++; ind[tgt[e]]++; }
    int total = 0, maxd = 0;
    for (int i = 0; i < n; i++) { deg[i] = outd[i] + ind[i]; total += deg[i]; if (deg[i] > maxd) maxd = deg[i]; }
    double avg = n > 0 ? (double)total / n : 0, var = 0;
    if (n > 0) { for (int i = 0; i < n; i++) { double d = deg[i] - avg; var += d * d; } var /= n; }
    ch->average_degree = avg; ch->max_degree = maxd; ch->degree_variance = var;
    ch->degree_cv = avg > 0 ? sqrt(var) / avg : 0;

    /* AnalyzeSupplyDistribution, :102-135 */
    int64_t max_abs = 0;
    for (int i = 0; i < n; i++) {
        int64_t s = supply[i];
        if (s > 0) { ch->source_count++; ch->total_supply += s; }
        else if (s < 0) ch->sink_count++;
        else ch->transshipment_count++;
        int64_t a = s < 0 ? -s : s; if (a > max_abs) max_abs = a;
    }
    ch->max_absolute_supply = max_abs;

    /* AnalyzeCosts, :137-176 */
    if (m == 0) { ch->has_uniform_costs = 1; }
    else {
        int64_t mn = I64_MAX, mx = INT64_MIN, tot = 0;
        for (int i = 0; i < m; i++) { if (cost[i] < mn) mn = cost[i]; if (cost[i] > mx) mx = cost[i]; tot += cost[i]; }
        double ac = (double)tot / m, v = 0;
        for (int i = 0; i < m; i++) { double d = cost[i] - ac; v += d * d; }
        v /= m;
        ch->cost_range = mx - mn; ch->average_cost = ac; ch->cost_variance = v;
        ch->cost_cv = fabs(ac) > 0 ? sqrt(v) / fabs(ac) : 0;
        ch->has_uniform_costs = ch->cost_cv < 0.01;
    }

    /* AnalyzeCapacities, :178-222 */
    if (m == 0) { ch->has_uniform_capacities = 1; }
    else {
        int64_t mn = I64_MAX, mx = INT64_MIN, tot = 0; int finite = 0;
        for (int i = 0; i < m; i++) {
            int64_t c = upper[i] - lower[i];
            if (c < I64_MAX / 2) { if (c < mn) mn = c; if (c > mx) mx = c; tot += c; finite++; }
        }
        if (finite > 0) { ch->capacity_range = mx - mn; ch->average_capacity = (double)tot / finite; ch->has_uniform_capacities = (mx - mn) == 0; }
        else ch->has_uniform_capacities = 1;
        ch->finite_capacity_ratio = (double)finite / m;
    }

    /* DetectProblemType + CheckBipartite, :224-289 */
    if (ch->source_count == 0 && ch->sink_count == 0) ch->detected_type = NS_TYPE_CIRCULATION;
    else {
        int only_out = 0, only_in = 0;
        for (int i = 0; i < n; i++) {
            if (outd[i] > 0 && ind[i] == 0) only_out++;
            else if (outd[i] == 0 && ind[i] > 0) only_in++;
        }
        int bip = ((double)(only_out + only_in) / n) > 0.8;
        ch->detected_type = NS_TYPE_GENERAL;
        if (bip && max_abs == 1 && ch->source_count == ch->sink_count) ch->detected_type = NS_TYPE_ASSIGNMENT;
        else if (bip && ch->transshipment_count == 0) ch->detected_type = NS_TYPE_TRANSPORTATION;
        else if (ch->transshipment_count > 0) ch->detected_type = NS_TYPE_TRANSSHIPMENT;
    }
    /* CheckForLayeredStructure runs before IsSparse is assigned (:54 vs :59) => IsSparse is still false */
    ch->is_layered = 0;
    ch->is_dense = ch->density > 0.01 || m > 10000;
    ch->is_sparse = ch->density < 0.005;
    free(deg); free(outd); free(ind);
}

void ns_oracle_select_config(const ns_oracle_characteristics *ch, ns_oracle_config *cfg)
{   /* OptimizationSelector.cs:14-96 */
    ns_oracle_default_config(cfg);
    int flags = 0;
    if (ch->is_dense) { flags |= NS_FLAG_SMALL_BLOCKS_FOR_DENSE; cfg->min_block_size = 10; cfg->max_block_size = 50; cfg->dense_network_threshold = 5000; }
    else { cfg->min_block_size = 25; cfg->max_block_size = 100; }
    if (ch->degree_cv > 0.5) {
        flags |= NS_FLAG_ADAPTIVE_BLOCK_SIZE;
        cfg->block_size_growth_factor = 1.3; cfg->block_size_shrink_factor = 0.7; cfg->consecutive_hits_before_adapt = 2;
    } else if (ch->degree_cv > 0.3) flags |= NS_FLAG_ADAPTIVE_BLOCK_SIZE;
    if (ch->is_sparse && ch->arc_count < 50000) flags |= NS_FLAG_REDUCED_COST_CACHING;
    /* ShouldUseCandidateList, :98-113 */
    if (ch->arc_count >= 1000 &&
        ((ch->is_sparse && ch->arc_count > 5000) || ch->has_uniform_costs ||
         ch->detected_type == NS_TYPE_ASSIGNMENT || ch->detected_type == NS_TYPE_TRANSPORTATION)) {
        flags |= NS_FLAG_CANDIDATE_LIST_PIVOT;
        cfg->candidate_list_ratio = ch->has_uniform_costs ? 0.2 : (ch->arc_count > 100000 ? 0.05 : 0.1);
    }
    if (ch->node_count > 5000 && ch->degree_cv > 1.0) flags |= NS_FLAG_HOT_COLD_SPLITTING;
    if (ch->detected_type == NS_TYPE_ASSIGNMENT || ch->detected_type == NS_TYPE_TRANSPORTATION) flags |= NS_FLAG_EARLY_TERMINATION;
    cfg->low_hit_rate_threshold = ch->arc_count > 10000 ? 0.03 : 0.05;
    cfg->high_hit_rate_threshold = ch->arc_count > 10000 ? 0.25 : 0.3;
    if (ch->arc_count > 100000) cfg->min_block_size_ratio = 0.0625;
    else if (ch->arc_count > 10000) cfg->min_block_size_ratio = 0.125;
    else cfg->min_block_size_ratio = 0.25;
    cfg->flags = flags;
}

/* ------------------------------------------------------------------------- */
/* NS.cs:624-669                                                               */

static int check_bounds(ns_t *s)
{
    for (int i = 0; i < s->m; i++) if (s->upper[i] < s->lower[i]) return 0;
    return 1;
}

static void transform_to_standard_form(ns_t *s)
{
    for (int i = 0; i < s->m; i++) {
        if (s->lower[i] != 0) {
            int u = s->source[i], v = s->target[i];
            s->supply[u] -= s->lower[i]; s->supply[v] += s->lower[i];
            s->upper[i] -= s->lower[i]; s->lower[i] = 0;
        }
    }
    s->sum_supply = 0;
    for (int i = 0; i < s->n; i++) s->sum_supply += s->supply[i];
    int64_t max_cost = 0;
    for (int i = 0; i < s->m; i++) { int64_t a = s->cost[i] < 0 ? -s->cost[i] : s->cost[i]; if (a > max_cost) max_cost = a; }
    s->art_cost = (max_cost + 1) * s->n;
}

/* NS.cs:671-845 */
static void initialize(ns_t *s)
{
    int n = s->n, m = s->m, root = n;
    s->root = root;
    s->parent[root] = -1; s->pred[root] = -1; s->thread[root] = 0; s->rev_thread[0] = root;
    s->succ_num[root] = n + 1; s->last_succ[root] = n - 1; s->pred_dir[root] = 0;

    int geq = s->opt->supply_type == NS_SUPPLY_GEQ;
    for (int i = 0; i < m; i++) { s->state[i] = STATE_LOWER; s->flow[i] = 0; }
    s->search_arc_num = m + n;
    int f = m + n;
    for (int u = 0; u < n; u++) s->thread[u] = u + 1;
    if (n > 0) s->thread[n - 1] = root;
    for (int u = 0; u < n; u++) s->rev_thread[s->thread[u]] = u;
    for (int u = 0, e = m; u < n; u++, e++) {
        s->parent[u] = root; s->succ_num[u] = 1; s->last_succ[u] = u;
        if (geq) {                                           /* InitializeGEQ, NS.cs:713-778 */
            if (s->supply[u] <= 0) {
                s->pred_dir[u] = DIR_DOWN; s->pi[u] = 0; s->pred[u] = e;
                s->source[e] = root; s->target[e] = u; s->upper[e] = NS_INF; s->flow[e] = -s->supply[u];
                s->cost[e] = 0; s->state[e] = STATE_TREE;
            } else {
                s->pred_dir[u] = DIR_UP; s->pi[u] = -s->art_cost; s->pred[u] = f;
                s->source[f] = u; s->target[f] = root; s->upper[f] = NS_INF; s->flow[f] = s->supply[u];
                s->state[f] = STATE_TREE; s->cost[f] = s->art_cost;
                s->source[e] = root; s->target[e] = u; s->upper[e] = NS_INF; s->flow[e] = 0;
                s->cost[e] = 0; s->state[e] = STATE_LOWER;
                f++;
            }
        } else {                                             /* InitializeLEQ, NS.cs:780-845 */
            if (s->supply[u] >= 0) {
                s->pred_dir[u] = DIR_UP; s->pi[u] = 0; s->pred[u] = e;
                s->source[e] = u; s->target[e] = root; s->upper[e] = NS_INF; s->flow[e] = s->supply[u];
                s->cost[e] = 0; s->state[e] = STATE_TREE;
            } else {
                s->pred_dir[u] = DIR_DOWN; s->pi[u] = s->art_cost; s->pred[u] = f;
                s->source[f] = root; s->target[f] = u; s->upper[f] = NS_INF; s->flow[f] = -s->supply[u];
                s->state[f] = STATE_TREE; s->cost[f] = s->art_cost;
                s->source[e] = u; s->target[e] = root; s->upper[e] = NS_INF; s->flow[e] = 0;
                s->cost[e] = 0; s->state[e] = STATE_LOWER;
                f++;
            }
        }
    }
    if (n > 0) { s->thread[n - 1] = root; s->rev_thread[root] = n - 1; }
    s->all_arc_num = s->search_arc_num;                      /* NS.cs:689 (overwrites f) */
    if (s->cfg.flags & NS_FLAG_REDUCED_COST_CACHING) {       /* NS.cs:692-696 */
        s->reduced_costs = (int64_t *)calloc((size_t)(s->all_arc_num > 0 ? s->all_arc_num : 1), sizeof(int64_t));
        s->reduced_costs_dirty = 1;
    }
}

/* ------------------------------------------------------------------------- */
/* pricing                                                                     */

static inline int64_t red_cost(const ns_t *s, int e)
{
    return s->state[e] * (s->cost[e] + s->pi[s->source[e]] - s->pi[s->target[e]]);
}

static void block_ctor(ns_t *s)
{   /* NS.cs:1304-1337 (identical in CachedBlockSearchPivot :1457-1490) */
    int base = (int)sqrt((double)s->search_arc_num);
    int r = (int)(base * s->cfg.min_block_size_ratio);
    s->dyn_min_block = s->cfg.min_block_size > r ? s->cfg.min_block_size : r;
    if (s->cfg.flags & NS_FLAG_SMALL_BLOCKS_FOR_DENSE) {
        double density = (double)s->search_arc_num / s->n;
        if (density > 10) s->block_size = 50 < base / 4 ? 50 : base / 4;
        else s->block_size = base;
    } else s->block_size = base;
    if (s->block_size < s->dyn_min_block) s->block_size = s->dyn_min_block;
    s->next_arc = 0; s->consecutive_low = s->consecutive_high = 0;
}

static void adapt_block(ns_t *s, int arcs_checked)
{   /* NS.cs:1399-1438 */
    if (!(s->cfg.flags & NS_FLAG_ADAPTIVE_BLOCK_SIZE)) return;
    double hit = arcs_checked > 0 ? 1.0 / arcs_checked : 0;
    if (hit < s->cfg.low_hit_rate_threshold) {
        s->consecutive_high = 0; s->consecutive_low++;
        if (s->consecutive_low >= s->cfg.consecutive_hits_before_adapt) {
            int ns = (int)(s->block_size * s->cfg.block_size_shrink_factor);
            s->block_size = s->dyn_min_block > ns ? s->dyn_min_block : ns;
            s->consecutive_low = 0;
        }
    } else if (hit > s->cfg.high_hit_rate_threshold) {
        s->consecutive_low = 0; s->consecutive_high++;
        if (s->consecutive_high >= s->cfg.consecutive_hits_before_adapt) {
            int ns = (int)(s->block_size * s->cfg.block_size_growth_factor);
            s->block_size = s->cfg.max_block_size < ns ? s->cfg.max_block_size : ns;
            s->consecutive_high = 0;
        }
    } else { s->consecutive_low = 0; s->consecutive_high = 0; }
}

/* NS.cs:1339-1441.  cached != 0 reads the reduced-cost cache instead (NS.cs:1492-1598). */
static int find_block(ns_t *s, int cached)
{
    int64_t min = 0; int cnt = s->block_size, e, arcs_checked = 0;
    const int S = s->search_arc_num;
    const int64_t *rc = s->reduced_costs;
    for (e = s->next_arc; e < S; e++) {
        arcs_checked++;
        int64_t c = cached ? rc[e] : red_cost(s, e);
        if (c < min) { min = c; s->in_arc = e; }
        if (--cnt == 0) { if (min < 0) goto search_end; cnt = s->block_size; }
    }
    for (e = 0; e < s->next_arc; e++) {
        arcs_checked++;
        int64_t c = cached ? rc[e] : red_cost(s, e);
        if (c < min) { min = c; s->in_arc = e; }
        if (--cnt == 0) { if (min < 0) goto search_end; cnt = s->block_size; }
    }
    if (min >= 0) { s->arcs_checked_pivot += arcs_checked; return 0; }
search_end:
    s->arcs_checked_pivot += arcs_checked;
    s->next_arc = e;
    adapt_block(s, arcs_checked);
    return 1;
}

/* UpdateReducedCosts, NS.cs:1211-1270.  _dirtyNodes is never allocated: Initialize() (NS.cs:692-696)
 * creates _reducedCosts first, so CreatePivotRuleFinder's `useCache && _reducedCosts == null`
 * branch (NS.cs:858-877) is dead and only the "full update when dirty" arm can run. */
static void update_reduced_costs(ns_t *s)
{
    if (!s->reduced_costs || !s->reduced_costs_dirty) return;
    int max_arc = s->search_arc_num < s->m ? s->search_arc_num : s->m;
    for (int e = 0; e < max_arc; e++)
        s->reduced_costs[e] = s->state[e] != STATE_TREE ? red_cost(s, e) : 0;
    s->reduced_costs_dirty = 0;
}

static int find_cached(ns_t *s) { update_reduced_costs(s); return find_block(s, 1); }

static int find_first(ns_t *s)
{   /* NS.cs:1607-1636 */
    const int S = s->search_arc_num;
    for (int e = s->next_arc; e < S; e++) if (red_cost(s, e) < 0) { s->in_arc = e; s->next_arc = e + 1; return 1; }
    for (int e = 0; e < s->next_arc; e++) if (red_cost(s, e) < 0) { s->in_arc = e; s->next_arc = e + 1; return 1; }
    return 0;
}

static int find_best(ns_t *s)
{   /* NS.cs:1644-1667 */
    int64_t min = 0; int best = -1; const int S = s->search_arc_num;
    for (int e = 0; e < S; e++) { int64_t c = red_cost(s, e); if (c < min) { min = c; best = e; } }
    s->arcs_checked_pivot += S;
    if (min < 0) { s->in_arc = best; return 1; }
    return 0;
}

/* CandidateListPivotRule, lemon-1.3.1/lemon/network_simplex.h:413-518 (constructor :441-458, findEnteringArc :461-516),
 * on the C# port's arrays and arc order (the port has no implementation: NS.cs:884 throws). */
static void candidate_ctor(ns_t *s)
{
    int l = (int)(0.25 * sqrt((double)s->search_arc_num));
    s->list_length = l > 10 ? l : 10;
    int ml = (int)(0.1 * s->list_length);
    s->minor_limit = ml > 3 ? ml : 3;
    s->curr_length = s->minor_count = 0; s->next_arc = 0;
    s->candidates = (int *)calloc((size_t)s->list_length, 4);
}

static int find_candidate_list(ns_t *s)
{
    int64_t min, c; int e; const int S = s->search_arc_num;
    if (s->curr_length > 0 && s->minor_count < s->minor_limit) {
        /* minor iteration: best eligible arc of the list; arcs that are no longer eligible are replaced by the last entry */
        s->minor_count++;
        min = 0;
        for (int i = 0; i < s->curr_length; ++i) {
            e = s->candidates[i];
            c = red_cost(s, e);
            s->arcs_checked_pivot++;
            if (c < min) { min = c; s->in_arc = e; }
            else if (c >= 0) s->candidates[i--] = s->candidates[--s->curr_length];
        }
        if (min < 0) return 1;
    }
    /* major iteration: a new list from the cyclic scan */
    min = 0; s->curr_length = 0;
    for (e = s->next_arc; e != S; ++e) {
        c = red_cost(s, e); s->arcs_checked_pivot++;
        if (c < 0) {
            s->candidates[s->curr_length++] = e;
            if (c < min) { min = c; s->in_arc = e; }
            if (s->curr_length == s->list_length) goto search_end;
        }
    }
    for (e = 0; e != s->next_arc; ++e) {
        c = red_cost(s, e); s->arcs_checked_pivot++;
        if (c < 0) {
            s->candidates[s->curr_length++] = e;
            if (c < min) { min = c; s->in_arc = e; }
            if (s->curr_length == s->list_length) goto search_end;
        }
    }
    if (s->curr_length == 0) return 0;
search_end:
    s->minor_count = 1;
    s->next_arc = e;
    return 1;
}

/* AlteringListPivotRule, network_simplex.h:521-635 (constructor :563-580, findEnteringArc :583-633).  std::partial_sort
 * leaves the order of entries with equal cost unspecified; this restatement - and the GPU engine with it - orders ties
 * by their position in the list when the sort starts (a stable partial sort, one of the orders the standard allows). */
static void altering_ctor(ns_t *s)
{
    int b = (int)(1.0 * sqrt((double)s->search_arc_num));
    s->alt_block = b > 10 ? b : 10;
    int h = (int)(0.01 * s->alt_block);
    s->head_length = h > 3 ? h : 3;
    s->candidates = (int *)calloc((size_t)s->head_length + s->alt_block, 4);
    s->cand_cost = (int64_t *)calloc((size_t)s->search_arc_num + 1, 8);
    s->curr_length = 0; s->next_arc = 0;
}

static int find_altering_list(ns_t *s)
{
    int e; int64_t c; const int S = s->search_arc_num;
    for (int i = 0; i != s->curr_length; ++i) {          /* check the current list */
        e = s->candidates[i];
        c = red_cost(s, e); s->arcs_checked_pivot++;
        if (c < 0) s->cand_cost[e] = c;
        else s->candidates[i--] = s->candidates[--s->curr_length];
    }
    int cnt = s->alt_block, limit = s->head_length;      /* extend it */
    for (e = s->next_arc; e != S; ++e) {
        c = red_cost(s, e); s->arcs_checked_pivot++;
        if (c < 0) { s->cand_cost[e] = c; s->candidates[s->curr_length++] = e; }
        if (--cnt == 0) { if (s->curr_length > limit) goto search_end; limit = 0; cnt = s->alt_block; }
    }
    for (e = 0; e != s->next_arc; ++e) {
        c = red_cost(s, e); s->arcs_checked_pivot++;
        if (c < 0) { s->cand_cost[e] = c; s->candidates[s->curr_length++] = e; }
        if (--cnt == 0) { if (s->curr_length > limit) goto search_end; limit = 0; cnt = s->alt_block; }
    }
    if (s->curr_length == 0) return 0;
search_end:;
    /* partial sort: the new_length cheapest entries in front, ascending by (cost, position before the sort) */
    int new_length = s->head_length + 1 < s->curr_length ? s->head_length + 1 : s->curr_length;
    int *top = (int *)malloc((size_t)new_length * 4);     /* positions, kept sorted */
    int nt = 0;
    for (int i = 0; i < s->curr_length; ++i) {
        const int64_t ci = s->cand_cost[s->candidates[i]];
        if (nt == new_length && !(ci < s->cand_cost[s->candidates[top[nt - 1]]])) continue;
        int j = nt < new_length ? nt++ : nt - 1;
        while (j > 0 && ci < s->cand_cost[s->candidates[top[j - 1]]]) { top[j] = top[j - 1]; --j; }
        top[j] = i;
    }
    for (int j = 0; j < nt; ++j) top[j] = s->candidates[top[j]];
    memcpy(s->candidates, top, (size_t)nt * 4);
    free(top);
    s->in_arc = s->candidates[0];
    s->next_arc = e;
    s->candidates[0] = s->candidates[new_length - 1];
    s->curr_length = new_length - 1;
    return 1;
}

/* Internal/BlockSearchPivotOptimized.cs:39-157.  ProcessArcRange falls through into its scalar loop
 * after ProcessArcRangeSIMD returns early (:74-80): cnt is then 0, `--cnt == 0` cannot fire again, and the
 * rest of the range is scanned to its end.  opt->simd_width (Vector<long>.Count; 4 on AVX2, 0 = not
 * hardware accelerated) selects that behaviour. */
static int opt_range(ns_t *s, int start, int end, int64_t *min, int *cnt, int *best)
{
    int e = start, vc = s->opt->simd_width;
    if (vc > 0 && end - start >= vc * 2) {
        int done = 0;
        for (; e <= end - vc && !done; e += vc) {
            for (int i = 0; i < vc; i++) {
                int idx = e + i;
                int64_t c = red_cost(s, idx);
                if (c < *min) { *min = c; *best = idx; }
                if (--*cnt == 0) { if (*min < 0) { e = idx + 1 - vc; done = 1; break; } *cnt = s->block_size; }
            }
        }
    }
    for (; e < end; e++) {
        int64_t c = red_cost(s, e);
        if (c < *min) { *min = c; *best = e; }
        if (--*cnt == 0) { if (*min < 0) return e + 1; *cnt = s->block_size; }
    }
    return e;
}

static int find_block_optimized(ns_t *s)
{
    int64_t min = 0; int cnt = s->block_size, best = -1, S = s->search_arc_num;
    int e = opt_range(s, s->next_arc, S, &min, &cnt, &best);
    if (e >= S && min >= 0) e = opt_range(s, 0, s->next_arc, &min, &cnt, &best);
    if (min >= 0) return 0;
    s->next_arc = e; s->in_arc = best;
    return 1;
}

static int find_first_optimized(ns_t *s)
{   /* BlockSearchPivotOptimized.cs:176-232: skips tree arcs, otherwise as find_first */
    const int S = s->search_arc_num;
    for (int e = s->next_arc; e < S; e++) { if (s->state[e] == 0) continue; if (red_cost(s, e) < 0) { s->in_arc = e; s->next_arc = e + 1; return 1; } }
    for (int e = 0; e < s->next_arc; e++) { if (s->state[e] == 0) continue; if (red_cost(s, e) < 0) { s->in_arc = e; s->next_arc = e + 1; return 1; } }
    return 0;
}

static int find_best_optimized(ns_t *s)
{   /* BlockSearchPivotOptimized.cs:251-289 */
    int64_t min = 0; int best = -1; const int S = s->search_arc_num;
    for (int e = 0; e < S; e++) { if (s->state[e] == 0) continue; int64_t c = red_cost(s, e); if (c < min) { min = c; best = e; } }
    if (min >= 0) return 0;
    s->in_arc = best; return 1;
}

/* ------------------------------------------------------------------------- */
/* NS.cs:925-1209                                                              */

static void find_join_node(ns_t *s)
{
    int u = s->source[s->in_arc], v = s->target[s->in_arc];
    int64_t steps = 0;
    while (u != v) {
        if (s->succ_num[u] < s->succ_num[v]) u = s->parent[u]; else v = s->parent[v];
        steps++;
    }
    s->join = u;
    s->res->join_steps += steps;
    if (steps > s->res->max_join_steps) s->res->max_join_steps = steps;
}

static int find_leaving_arc(ns_t *s)
{
    int first, second;
    if (s->state[s->in_arc] == STATE_LOWER) { first = s->source[s->in_arc]; second = s->target[s->in_arc]; }
    else { first = s->target[s->in_arc]; second = s->source[s->in_arc]; }
    s->delta = s->upper[s->in_arc];
    int result = 0; int64_t d; int e;
    for (int u = first; u != s->join; u = s->parent[u]) {
        e = s->pred[u]; d = s->flow[e];
        if (s->pred_dir[u] == DIR_DOWN) { int64_t c = s->upper[e]; d = c >= I64_MAX ? NS_INF : c - d; }
        if (d < s->delta) { s->delta = d; s->u_out = u; result = 1; }
    }
    for (int u = second; u != s->join; u = s->parent[u]) {
        e = s->pred[u]; d = s->flow[e];
        if (s->pred_dir[u] == DIR_UP) { int64_t c = s->upper[e]; d = c >= I64_MAX ? NS_INF : c - d; }
        if (d <= s->delta) { s->delta = d; s->u_out = u; result = 2; }
    }
    if (result == 1) { s->u_in = first; s->v_in = second; }
    else { s->u_in = second; s->v_in = first; }
    return result != 0;
}

static void change_flow(ns_t *s, int change)
{
    if (s->delta > 0) {
        int64_t val = s->state[s->in_arc] * s->delta;
        s->flow[s->in_arc] += val;
        for (int u = s->source[s->in_arc]; u != s->join; u = s->parent[u]) s->flow[s->pred[u]] -= s->pred_dir[u] * val;
        for (int u = s->target[s->in_arc]; u != s->join; u = s->parent[u]) s->flow[s->pred[u]] += s->pred_dir[u] * val;
    }
    if (change) {
        s->state[s->in_arc] = STATE_TREE;
        int le = s->pred[s->u_out];
        s->state[le] = s->flow[le] == 0 ? STATE_LOWER : STATE_UPPER;
    } else s->state[s->in_arc] = (int8_t)-s->state[s->in_arc];
}

static void update_tree_structure(ns_t *s)
{
    int *parent = s->parent, *pred = s->pred, *thread = s->thread, *rev_thread = s->rev_thread;
    int *succ_num = s->succ_num, *last_succ = s->last_succ; int8_t *pred_dir = s->pred_dir;
    int u_in = s->u_in, v_in = s->v_in, u_out = s->u_out, join = s->join, in_arc = s->in_arc;
    int old_rev_thread = rev_thread[u_out], old_succ_num = succ_num[u_out], old_last_succ = last_succ[u_out];
    int v_out = parent[u_out]; s->v_out = v_out;
    int64_t stem_len = 1;

    if (u_in == u_out) {
        parent[u_in] = v_in; pred[u_in] = in_arc;
        pred_dir[u_in] = u_in == s->source[in_arc] ? DIR_UP : DIR_DOWN;
        if (thread[v_in] != u_out) {
            int after = thread[old_last_succ];
            thread[old_rev_thread] = after; rev_thread[after] = old_rev_thread;
            after = thread[v_in];
            thread[v_in] = u_out; rev_thread[u_out] = v_in;
            thread[old_last_succ] = after; rev_thread[after] = old_last_succ;
        }
    } else {
        int thread_continue = old_rev_thread == v_in ? thread[old_last_succ] : thread[v_in];
        int stem = u_in, par_stem = v_in, next_stem, last = last_succ[u_in], before, after = thread[last];
        thread[v_in] = u_in;
        int *dirty = s->dirty_revs;
        if (s->opt->emulate_stackalloc) memset(dirty, 0, (size_t)s->n * sizeof(int));   /* NS.cs:1085: stackalloc is zero-initialised */
        dirty[0] = v_in; int dirty_count = 1;
        while (stem != u_out) {
            next_stem = parent[stem];
            thread[last] = next_stem; dirty[dirty_count++] = last;
            before = rev_thread[stem];
            thread[before] = after; rev_thread[after] = before;
            parent[stem] = par_stem; par_stem = stem; stem = next_stem;
            last = last_succ[stem] == last_succ[par_stem] ? rev_thread[par_stem] : last_succ[stem];
            after = thread[last];
            stem_len++;
        }
        parent[u_out] = par_stem;
        thread[last] = thread_continue; rev_thread[thread_continue] = last;
        last_succ[u_out] = last;
        if (old_rev_thread != v_in) { thread[old_rev_thread] = after; rev_thread[after] = old_rev_thread; }
        for (int i = 0; i < dirty_count; ++i) { int u = dirty[i]; rev_thread[thread[u]] = u; }
        int tmp_sc = 0, tmp_ls = last_succ[u_out];
        for (int u = u_out, p = parent[u]; u != u_in; u = p, p = parent[u]) {
            pred[u] = pred[p]; pred_dir[u] = (int8_t)-pred_dir[p];
            tmp_sc += succ_num[u] - succ_num[p]; succ_num[u] = tmp_sc;
            last_succ[p] = tmp_ls;
        }
        pred[u_in] = in_arc;
        pred_dir[u_in] = u_in == s->source[in_arc] ? DIR_UP : DIR_DOWN;
        succ_num[u_in] = old_succ_num;
    }

    int up_limit_out = last_succ[join] == v_in ? join : -1;
    int last_succ_out = last_succ[u_out];
    for (int u = v_in; u != -1 && last_succ[u] == v_in; u = parent[u]) last_succ[u] = last_succ_out;
    if (join != old_rev_thread && v_in != old_rev_thread) {
        for (int u = v_out; u != up_limit_out && last_succ[u] == old_last_succ; u = parent[u]) last_succ[u] = old_rev_thread;
    } else if (last_succ_out != old_last_succ) {
        for (int u = v_out; u != up_limit_out && last_succ[u] == old_last_succ; u = parent[u]) last_succ[u] = last_succ_out;
    }
    for (int u = v_in; u != join; u = parent[u]) succ_num[u] += old_succ_num;
    for (int u = v_out; u != join; u = parent[u]) succ_num[u] -= old_succ_num;
    s->res->stem_nodes += stem_len;
}

static void update_potentials(ns_t *s)
{
    int64_t sigma = s->pi[s->v_in] - s->pi[s->u_in] - s->pred_dir[s->u_in] * s->cost[s->in_arc];
    int end = s->thread[s->last_succ[s->u_in]];
    int64_t cnt = 0;
    for (int u = s->u_in; u != end; u = s->thread[u]) { s->pi[u] += sigma; cnt++; }
    s->res->subtree_nodes += cnt;
    if (cnt > s->res->max_subtree_nodes) s->res->max_subtree_nodes = cnt;
    if (s->reduced_costs != NULL) s->reduced_costs_dirty = 1;   /* NS.cs:1205-1208 (_dirtyNodes == null) */
}

static int check_feasibility(ns_t *s)
{   /* NS.cs:1272-1283 */
    for (int e = s->m; e < s->all_arc_num; e++) if (s->flow[e] != 0) return 0;
    return 1;
}

/* ------------------------------------------------------------------------- */

int ns_oracle_solve(int n, int m, const int32_t *src, const int32_t *tgt, const int64_t *lower,
                    const int64_t *upper, const int64_t *cost, const int64_t *supply,
                    const ns_oracle_options *opt, ns_oracle_result *res, int64_t *flow_out, int64_t *pi_out)
{
    ns_t S_; ns_t *s = &S_; memset(s, 0, sizeof(*s));
    memset(res, 0, sizeof(*res));
    s->n = n; s->m = m; s->opt = opt; s->res = res;
    double t_total0 = now_s();
    size_t A = (size_t)m + 2 * (size_t)n + 1, N1 = (size_t)n + 1;
    s->lower = (int64_t *)calloc(A, 8); s->upper = (int64_t *)calloc(A, 8); s->cost = (int64_t *)calloc(A, 8);
    s->flow = (int64_t *)calloc(A, 8); s->supply = (int64_t *)calloc(N1, 8); s->pi = (int64_t *)calloc(N1, 8);
    s->orig_lower = (int64_t *)calloc((size_t)m + 1, 8);
    s->source = (int *)calloc(A, 4); s->target = (int *)calloc(A, 4); s->state = (int8_t *)calloc(A, 1);
    s->parent = (int *)calloc(N1, 4); s->pred = (int *)calloc(N1, 4); s->thread = (int *)calloc(N1, 4);
    s->rev_thread = (int *)calloc(N1, 4); s->succ_num = (int *)calloc(N1, 4); s->last_succ = (int *)calloc(N1, 4);
    s->pred_dir = (int8_t *)calloc(N1, 1); s->dirty_revs = (int *)calloc(N1 + 1, 4);
    for (int i = 0; i < m; i++) {
        s->source[i] = src[i]; s->target[i] = tgt[i];
        s->lower[i] = lower ? lower[i] : 0; s->upper[i] = upper ? upper[i] : NS_INF; s->cost[i] = cost ? cost[i] : 0;
        s->orig_lower[i] = s->lower[i];
    }
    for (int i = 0; i < n; i++) s->supply[i] = supply ? supply[i] : 0;

    int status = NS_STATUS_NOT_SOLVED;
    int iterations = 0;
    if (!check_bounds(s)) { status = NS_STATUS_INFEASIBLE; goto done; }      /* NS.cs:227-231 */
    transform_to_standard_form(s);
    s->cfg = opt->config;
    if (opt->auto_config) {                                                   /* NS.cs:237-250 */
        ns_oracle_analyze(n, m, s->source, s->target, s->lower, s->upper, s->cost, s->supply, &res->characteristics);
        ns_oracle_select_config(&res->characteristics, &s->cfg);
    }
    res->config_used = s->cfg;
    initialize(s);

    /* CreatePivotRuleFinder, NS.cs:847-886 */
    int kind;
    if (opt->optimized_pivot) {
        int b = (int)sqrt((double)s->search_arc_num);
        s->block_size = b > 10 ? b : 10; s->next_arc = 0;                     /* BlockSearchPivotOptimized.cs:27-29 */
        kind = 10 + opt->pivot_rule;
    } else {
        int use_cache = (s->cfg.flags & NS_FLAG_REDUCED_COST_CACHING) != 0;
        kind = opt->pivot_rule == NS_PIVOT_BLOCK_SEARCH ? (use_cache ? 3 : 2) : opt->pivot_rule;
        if (opt->pivot_rule == NS_PIVOT_CANDIDATE_LIST) { kind = 4; candidate_ctor(s); }
        else if (opt->pivot_rule == NS_PIVOT_ALTERING_LIST) { kind = 5; altering_ctor(s); }
        else if (opt->pivot_rule == NS_PIVOT_BLOCK_SEARCH) block_ctor(s); else s->next_arc = 0;
    }
    res->pivot_kind = kind;
    if (kind == 2 || kind == 3) res->initial_block_size = s->block_size;

    if (opt->resume) {                                                        /* continue a solve this code checkpointed */
        const ns_oracle_state *r = opt->resume;
        size_t AA = (size_t)m + 2 * (size_t)n;
        memcpy(s->parent, r->parent, N1 * 4); memcpy(s->pred, r->pred, N1 * 4); memcpy(s->thread, r->thread, N1 * 4);
        memcpy(s->rev_thread, r->rev_thread, N1 * 4); memcpy(s->succ_num, r->succ_num, N1 * 4); memcpy(s->last_succ, r->last_succ, N1 * 4);
        memcpy(s->pred_dir, r->pred_dir, N1); memcpy(s->state, r->state, AA); memcpy(s->flow, r->flow, AA * 8); memcpy(s->pi, r->pi, N1 * 8);
        iterations = (int)r->iterations; s->next_arc = r->next_arc; s->block_size = r->block_size;
        s->consecutive_low = r->consecutive_low; s->consecutive_high = r->consecutive_high;
    }
    if (opt->warm) {                                                          /* warm start: previous optimal basis, new costs */
        const ns_oracle_state *r = opt->warm;
        size_t AA = (size_t)m + 2 * (size_t)n;
        memcpy(s->parent, r->parent, N1 * 4); memcpy(s->pred, r->pred, N1 * 4); memcpy(s->thread, r->thread, N1 * 4);
        memcpy(s->rev_thread, r->rev_thread, N1 * 4); memcpy(s->succ_num, r->succ_num, N1 * 4); memcpy(s->last_succ, r->last_succ, N1 * 4);
        memcpy(s->pred_dir, r->pred_dir, N1); memcpy(s->state, r->state, AA); memcpy(s->flow, r->flow, AA * 8);
        /* reduced cost 0 on every tree arc (the invariant UpdatePotentials maintains, NS.cs:1185-1209), root potential 0,
         * in thread order so that a parent is set before its children; initialize() above has put the current costs - and the
         * current artificial cost - on the arcs */
        s->pi[s->root] = 0;
        for (int u = s->thread[s->root]; u != s->root; u = s->thread[u]) {
            int e = s->pred[u], p = s->parent[u];
            s->pi[u] = s->pred_dir[u] == DIR_UP ? s->pi[p] - s->cost[e] : s->pi[p] + s->cost[e];
        }
    }
    int64_t max_iterations = (int64_t)n * m; if (max_iterations < 1000000) max_iterations = 1000000;  /* NS.cs:280 */
    double t_price = 0, t_tree = 0, t_pot = 0;
    const int timing = opt->collect_phase_times;
    double t_loop0 = now_s();
    for (;;) {
        double t0 = timing ? now_s() : 0;
        s->arcs_checked_pivot = 0;
        int found;
        switch (kind) {
            case 0: found = find_first(s); break;
            case 1: found = find_best(s); break;
            case 2: found = find_block(s, 0); break;
            case 3: found = find_cached(s); break;
            case 4: found = find_candidate_list(s); break;
            case 5: found = find_altering_list(s); break;
            case 10: found = find_first_optimized(s); break;
            case 11: found = find_best_optimized(s); break;
            default: found = find_block_optimized(s); break;
        }
        if (timing) t_price += now_s() - t0;
        res->total_arcs_checked += s->arcs_checked_pivot;
        if (!found) break;
        iterations++;
        if (iterations > max_iterations) { status = NS_STATUS_INFEASIBLE; res->iterations = iterations; goto done; }
        if (opt->trace_capacity > 0 && iterations <= opt->trace_capacity && opt->trace_in_arc) opt->trace_in_arc[iterations - 1] = s->in_arc;
        find_join_node(s);
        int change = find_leaving_arc(s);
        if (!change && s->delta == 0) { status = NS_STATUS_UNBOUNDED; res->iterations = iterations; goto done; }   /* NS.cs:321-325 */
        if (s->delta == 0) res->degenerate_pivots++;
        change_flow(s, change);
        if (opt->trace_capacity > 0 && iterations <= opt->trace_capacity && opt->trace_u_out) opt->trace_u_out[iterations - 1] = change ? s->u_out : -1;
        if (change) {
            double t1 = timing ? now_s() : 0;
            update_tree_structure(s);
            double t2 = timing ? now_s() : 0;
            update_potentials(s);
            if (timing) { t_tree += t2 - t1; t_pot += now_s() - t2; }
        }
        if (opt->max_pivots > 0 && iterations >= opt->max_pivots) { res->stopped_early = 1; break; }
    }
    res->loop_seconds = now_s() - t_loop0;
    res->iterations = iterations;
    if (kind == 2 || kind == 3) res->final_block_size = s->block_size;
    res->pricing_seconds = t_price; res->tree_seconds = t_tree; res->potential_seconds = t_pot;

    if (opt->save) {
        ns_oracle_state *w = opt->save;
        size_t AA = (size_t)m + 2 * (size_t)n;
        memcpy(w->parent, s->parent, N1 * 4); memcpy(w->pred, s->pred, N1 * 4); memcpy(w->thread, s->thread, N1 * 4);
        memcpy(w->rev_thread, s->rev_thread, N1 * 4); memcpy(w->succ_num, s->succ_num, N1 * 4); memcpy(w->last_succ, s->last_succ, N1 * 4);
        memcpy(w->pred_dir, s->pred_dir, N1); memcpy(w->state, s->state, AA); memcpy(w->flow, s->flow, AA * 8); memcpy(w->pi, s->pi, N1 * 8);
        w->iterations = iterations; w->next_arc = s->next_arc; w->block_size = s->block_size;
        w->consecutive_low = s->consecutive_low; w->consecutive_high = s->consecutive_high;
    }
    if (res->stopped_early) { status = NS_STATUS_NOT_SOLVED; }
    else if (check_feasibility(s)) {
        status = NS_STATUS_OPTIMAL;
        for (int i = 0; i < m; i++) if (s->orig_lower[i] != 0) s->flow[i] += s->orig_lower[i];  /* NS.cs:365-388 */
    } else status = NS_STATUS_INFEASIBLE;

done:
    res->status = status;
    res->total_cost = 0;
    for (int i = 0; i < m; i++) res->total_cost += s->flow[i] * s->cost[i];
    if (flow_out) memcpy(flow_out, s->flow, (size_t)m * 8);
    if (pi_out) memcpy(pi_out, s->pi, (size_t)n * 8);
    res->art_cost = s->art_cost; res->sum_supply = s->sum_supply;
    res->total_seconds = now_s() - t_total0;
    free(s->lower); free(s->upper); free(s->cost); free(s->flow); free(s->supply); free(s->pi); free(s->orig_lower);
    free(s->source); free(s->target); free(s->state); free(s->parent); free(s->pred); free(s->thread);
    free(s->rev_thread); free(s->succ_num); free(s->last_succ); free(s->pred_dir); free(s->dirty_revs);
    free(s->reduced_costs); free(s->candidates); free(s->cand_cost);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Validation/SolutionValidator.cs:20-342 restated over flat arrays.  Returns a bit mask of failed checks. */

int ns_oracle_validate(int n, int m, const int32_t *src, const int32_t *tgt, const int64_t *lower,
                       const int64_t *upper, const int64_t *cost, const int64_t *supply, int supply_type,
                       const int64_t *flow, const int64_t *pi, int64_t reported_cost, int64_t *dual_cost_out)
{
    int bad = 0;
    int64_t *net = (int64_t *)calloc((size_t)n + 1, 8);
    for (int e = 0; e < m; e++) { net[src[e]] += flow[e]; net[tgt[e]] -= flow[e]; }
    for (int i = 0; i < n; i++) {                                            /* :55-100 */
        int ok = supply_type == NS_SUPPLY_GEQ ? net[i] >= supply[i] : net[i] <= supply[i];
        if (!ok) bad |= 1;
    }
    for (int e = 0; e < m; e++) {                                            /* :102-125 */
        int64_t lo = lower ? lower[e] : 0, up = upper ? upper[e] : NS_INF;
        if (flow[e] < lo || flow[e] > up) bad |= 2;
    }
    for (int e = 0; e < m; e++) {                                            /* :135-176 */
        int64_t lo = lower ? lower[e] : 0, up = upper ? upper[e] : NS_INF;
        int64_t rc = cost[e] + pi[src[e]] - pi[tgt[e]];
        if (rc > 0 && flow[e] != lo) bad |= 4;
        if (rc < 0 && flow[e] != up) bad |= 4;
    }
    for (int i = 0; i < n; i++) {                                            /* :191-231 */
        if (supply_type == NS_SUPPLY_GEQ) { if (pi[i] > 0) bad |= 8; else if (pi[i] < 0 && net[i] != supply[i]) bad |= 8; }
        else { if (pi[i] < 0) bad |= 8; else if (pi[i] > 0 && net[i] != supply[i]) bad |= 8; }
    }
    int64_t calc = 0;                                                        /* :234-262 */
    for (int e = 0; e < m; e++) calc += flow[e] * cost[e];
    if (calc != reported_cost) bad |= 16;
    int64_t dual = 0;                                                        /* :268-342 */
    int64_t *adj = (int64_t *)calloc((size_t)n + 1, 8);
    for (int i = 0; i < n; i++) adj[i] = supply[i];
    for (int e = 0; e < m; e++) {
        int64_t lo = lower ? lower[e] : 0;
        if (lo != 0) { dual += lo * cost[e]; adj[src[e]] -= lo; adj[tgt[e]] += lo; }
    }
    for (int i = 0; i < n; i++) dual -= adj[i] * pi[i];
    for (int e = 0; e < m; e++) {
        int64_t rc = cost[e] + pi[src[e]] - pi[tgt[e]];
        if (rc < 0) { int64_t lo = lower ? lower[e] : 0, up = upper ? upper[e] : NS_INF; dual -= (up - lo) * -rc; }
    }
    if (dual != reported_cost) bad |= 32;
    if (dual_cost_out) *dual_cost_out = dual;
    free(net); free(adj);
    return bad;
}
