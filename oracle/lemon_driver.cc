// lemon_driver.cc - flat-array entry point around the vendored LEMON 1.3.1 NetworkSimplex.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile (target `ref`) from the upstream sources where they
// lie under /root/reference/lemon-1.3.1 into oracle/_ref/liblemon_ns.so; no reference source is copied.
// LEMON is the C++ code the reference's C# solver was ported from (NetworkSimplex.cs:12-15); it is used as
// an independent optimal-cost / status oracle and as the "reference" CPU baseline.  It is NOT a
// pivot-sequence oracle: arc mixing, the EQ-form initial basis and initialPivots() differ.
#include <chrono>
#include <cstdint>
#include <limits>
#include <vector>

#include <lemon/list_graph.h>
#include <lemon/network_simplex.h>

using namespace lemon;

extern "C" int lemon_ns_solve(int n, int m, const int32_t* src, const int32_t* tgt, const int64_t* lower,
                              const int64_t* upper, const int64_t* cost, const int64_t* supply, int pivot_rule,
                              int supply_type, int64_t* total_cost, double* run_seconds, int64_t* flow_out,
                              int64_t* pi_out)
{
    typedef NetworkSimplex<ListDigraph, int64_t, int64_t> NS;
    ListDigraph g;
    std::vector<ListDigraph::Node> nodes(n);
    std::vector<ListDigraph::Arc> arcs(m);
    g.reserveNode(n); g.reserveArc(m);
    for (int i = 0; i < n; ++i) nodes[i] = g.addNode();
    for (int e = 0; e < m; ++e) arcs[e] = g.addArc(nodes[src[e]], nodes[tgt[e]]);
    ListDigraph::ArcMap<int64_t> lo(g), up(g), co(g);
    ListDigraph::NodeMap<int64_t> su(g);
    const int64_t ref_inf = std::numeric_limits<int64_t>::max() / 2;      // NetworkSimplex.cs:127
    for (int e = 0; e < m; ++e) {
        lo[arcs[e]] = lower ? lower[e] : 0;
        int64_t u = upper ? upper[e] : ref_inf;
        up[arcs[e]] = u >= ref_inf ? std::numeric_limits<int64_t>::max() : u;
        co[arcs[e]] = cost ? cost[e] : 0;
    }
    for (int i = 0; i < n; ++i) su[nodes[i]] = supply ? supply[i] : 0;
    NS ns(g);
    ns.lowerMap(lo).upperMap(up).costMap(co).supplyMap(su);
    ns.supplyType(supply_type == 0 ? NS::GEQ : NS::LEQ);
    NS::PivotRule rule = pivot_rule == 0 ? NS::FIRST_ELIGIBLE : pivot_rule == 1 ? NS::BEST_ELIGIBLE : pivot_rule == 3 ? NS::CANDIDATE_LIST
                       : pivot_rule == 4 ? NS::ALTERING_LIST : NS::BLOCK_SEARCH;       // PivotRule.cs:7-40 values
    auto t0 = std::chrono::steady_clock::now();
    NS::ProblemType st = ns.run(rule);
    auto t1 = std::chrono::steady_clock::now();
    *run_seconds = std::chrono::duration<double>(t1 - t0).count();
    int status = st == NS::OPTIMAL ? 1 : st == NS::INFEASIBLE ? 2 : 3;    // SolverStatus.cs:7-34 values
    *total_cost = 0;
    if (st == NS::OPTIMAL) {
        *total_cost = ns.totalCost<int64_t>();
        if (flow_out) for (int e = 0; e < m; ++e) flow_out[e] = ns.flow(arcs[e]);
        if (pi_out) for (int i = 0; i < n; ++i) pi_out[i] = ns.potential(nodes[i]);
    }
    return status;
}
