// The reference's own solver tests, restated against the C++ mirror of its API (include/mcf_network_simplex.hpp) over libmcfgpu.so:
//   src/MinCostFlow.Tests/Lemon/NetworkSimplexTests.cs:28-79   SimpleTransportationProblem_SolvesCorrectly
//                                                      :80-123  MinimumCostCirculation_SolvesCorrectly
//                                                      :125-180 ExcessSupply_FeasibleWithGEQ
//                                                      :182-203 LowerBounds_RespectedInSolution
//                                                      :205-245 ComplementarySlackness_ValidatedCorrectly
//   src/MinCostFlow.Tests/Lemon/OptimizationTests.cs:14-70       optimized pivot == baseline (status, cost, flows)
// plus the reference's error behaviour (NetworkSimplex.cs:155-158, :418-421, :884) and the DIMACS path.
// Usage: test_network_simplex [--no-device] <dir with grid_5x5.min and netgen_8_08a.min>      exit code 0 = all passed.
// (tests/test_cpp_mirror.py writes the directory: conftest.dimacs_dir)
#include <cstdio>
#include <cstring>
#include <string>

#include "mcf_network_simplex.hpp"

using namespace mcf;

static int g_failed = 0, g_checked = 0;
#define CHECK(cond) do { ++g_checked; if (!(cond)) { ++g_failed; std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); } } while (0)
#define CHECK_THROWS(expr, Ex) do { ++g_checked; bool ok__ = false; try { expr; } catch (const Ex&) { ok__ = true; } catch (...) {} \
    if (!ok__) { ++g_failed; std::printf("FAILED %s:%d  %s does not throw %s\n", __FILE__, __LINE__, #expr, #Ex); } } while (0)

static void ValidateSolution(const NetworkSimplex& solver)
{
    const ValidationResult r = solver.Validate();
    CHECK(r.IsValid);
    CHECK(r.Errors.empty());
    CHECK(r.ObjectiveValue == solver.GetTotalCost());
    CHECK(r.DualCost == solver.GetTotalCost());
}

static void SimpleTransportationProblem_SolvesCorrectly()
{
    GraphBuilder builder;
    builder.AddNodes(4);
    builder.AddArc(0, 2).AddArc(0, 3).AddArc(1, 2).AddArc(1, 3);
    const CompactDigraph& graph = builder.Build();
    NetworkSimplex solver(graph);
    solver.SetNodeSupply(builder.GetNode(0), 10).SetNodeSupply(builder.GetNode(1), 15);
    solver.SetNodeSupply(builder.GetNode(2), -12).SetNodeSupply(builder.GetNode(3), -13);
    solver.SetArcCost(Arc(0), 3).SetArcCost(Arc(1), 5).SetArcCost(Arc(2), 4).SetArcCost(Arc(3), 2);
    CHECK(solver.Solve() == SolverStatus::Optimal);
    CHECK(solver.GetTotalCost() == 64);
    CHECK(solver.GetFlow(Arc(0)) == 10 && solver.GetFlow(Arc(1)) == 0 && solver.GetFlow(Arc(2)) == 2 && solver.GetFlow(Arc(3)) == 13);
    CHECK(solver.Status() == SolverStatus::Optimal && solver.SupplyType() == SupplyType::Geq);
    ValidateSolution(solver);
}

static void MinimumCostCirculation_SolvesCorrectly()
{
    GraphBuilder builder;
    builder.AddNodes(3);
    builder.AddArc(0, 1).AddArc(1, 2).AddArc(2, 0);
    NetworkSimplex solver(builder.Build());
    for (int u = 0; u < 3; ++u) solver.SetNodeSupply(builder.GetNode(u), 0);
    solver.SetArcCost(Arc(0), 2).SetArcCost(Arc(1), 3).SetArcCost(Arc(2), -6);
    for (int a = 0; a < 3; ++a) solver.SetArcBounds(Arc(a), 0, 10);
    CHECK(solver.Solve() == SolverStatus::Optimal);
    CHECK(solver.GetTotalCost() == -10);
    CHECK(solver.GetFlow(Arc(0)) == 10 && solver.GetFlow(Arc(1)) == 10 && solver.GetFlow(Arc(2)) == 10);
    ValidateSolution(solver);
}

static void ExcessSupply_FeasibleWithGEQ()
{
    GraphBuilder builder;
    builder.AddNodes(3);
    builder.AddArc(0, 1).AddArc(0, 2).AddArc(1, 2);
    NetworkSimplex solver(builder.Build());
    solver.SetNodeSupply(builder.GetNode(0), 10).SetNodeSupply(builder.GetNode(1), 0).SetNodeSupply(builder.GetNode(2), -10);
    solver.SetArcCost(Arc(0), 3).SetArcCost(Arc(1), 1).SetArcCost(Arc(2), 2);
    CHECK(solver.Solve() == SolverStatus::Optimal);
    CHECK(solver.SupplyType() == SupplyType::Geq);
    ValidateSolution(solver);
    CHECK(solver.GetFlow(Arc(1)) == 10 && solver.GetFlow(Arc(0)) == 0 && solver.GetFlow(Arc(2)) == 0);
    CHECK(solver.GetTotalCost() == 10);
}

static void LowerBounds_RespectedInSolution()
{
    GraphBuilder builder;
    builder.AddNodes(2);
    builder.AddArc(0, 1);
    NetworkSimplex solver(builder.Build());
    solver.SetNodeSupply(builder.GetNode(0), 10).SetNodeSupply(builder.GetNode(1), -10);
    solver.SetArcBounds(Arc(0), 5, 15).SetArcCost(Arc(0), 1);
    CHECK(solver.Solve() == SolverStatus::Optimal);
    CHECK(solver.GetFlow(Arc(0)) == 10 && solver.GetTotalCost() == 10);
    CHECK(solver.GetArcLowerBound(Arc(0)) == 5 && solver.GetArcCost(Arc(0)) == 1 && solver.GetNodeSupply(Node(0)) == 10);
    ValidateSolution(solver);
}

static void ComplementarySlackness_ValidatedCorrectly()
{
    GraphBuilder builder;
    builder.AddNodes(3);
    builder.AddArc(0, 1).AddArc(1, 2);
    NetworkSimplex solver(builder.Build());
    solver.SetNodeSupply(builder.GetNode(0), 10).SetNodeSupply(builder.GetNode(1), 0).SetNodeSupply(builder.GetNode(2), -10);
    solver.SetArcCost(Arc(0), 1).SetArcCost(Arc(1), 1);
    solver.SetArcBounds(Arc(0), 0, 20).SetArcBounds(Arc(1), 0, 20);
    CHECK(solver.Solve() == SolverStatus::Optimal);
    const ValidationResult r = solver.Validate();
    CHECK(r.IsValid && r.Errors.empty());
}

// OptimizationTests.cs:14-70 with SetupProblem :122-144; here for every rule and every Vector<long>.Count a host can have
static void OptimizedPivot_ProducesSameResults()
{
    GraphBuilder builder;
    builder.AddNodes(5);
    builder.AddArc(0, 1).AddArc(0, 2).AddArc(1, 3).AddArc(2, 3).AddArc(2, 4).AddArc(3, 4);
    const int64_t supply[5] = {50, 20, -10, -30, -30}, cost[6] = {10, 20, 30, 15, 25, 35};
    auto setup = [&](NetworkSimplex& s) {
        for (int u = 0; u < 5; ++u) s.SetNodeSupply(builder.GetNode(u), supply[u]);
        for (int a = 0; a < 6; ++a) s.SetArcCost(Arc(a), cost[a]).SetArcBounds(Arc(a), 0, 100);
    };
    const PivotRule rules[3] = {PivotRule::FirstEligible, PivotRule::BestEligible, PivotRule::BlockSearch};
    for (PivotRule rule : rules) {
        NetworkSimplex baseline(builder.Build());
        setup(baseline); baseline.SetPivotRule(rule);
        const SolverStatus st = baseline.Solve();
        CHECK(st == SolverStatus::Optimal);
        const int widths[3] = {4, 2, 0};
        for (int w : widths) {
            NetworkSimplex optimized(builder.Build());
            setup(optimized); optimized.SetPivotRule(rule); optimized.EnableOptimizedPivot(true, w);
            CHECK(optimized.Solve() == st);
            CHECK(optimized.GetTotalCost() == baseline.GetTotalCost());
            for (int a = 0; a < 6; ++a) CHECK(optimized.GetFlow(Arc(a)) == baseline.GetFlow(Arc(a)));
            ValidateSolution(optimized);
        }
        ValidateSolution(baseline);
    }
}

static void ErrorBehaviour()
{
    GraphBuilder builder;
    builder.AddNodes(2).AddArc(0, 1);
    CHECK_THROWS(builder.AddArc(0, 5), ArgumentException);                        // GraphBuilder.cs:64-67
    CHECK_THROWS(builder.AddNode(1), ArgumentException);                          // :32-35
    CHECK_THROWS(builder.GetNode(9), ArgumentException);                          // :78-81
    NetworkSimplex solver(builder.Build());
    CHECK_THROWS(solver.GetFlow(Arc(0)), InvalidOperationException);              // NetworkSimplex.cs:418-421
    CHECK_THROWS(solver.SetArcCost(Arc(3), 1), ArgumentException);                // :171-174
    CHECK_THROWS(solver.SetNodeSupply(Node(2), 1), ArgumentException);            // :185-188
    solver.SetNodeSupply(Node(0), 4).SetNodeSupply(Node(1), -4).SetArcCost(Arc(0), 7);
    CHECK(solver.Solve() == SolverStatus::Optimal && solver.GetTotalCost() == 28 && solver.GetFlow(Arc(0)) == 4);
    CHECK_THROWS(solver.GetFlow(Arc(1)), ArgumentException);
    CHECK_THROWS(solver.GetPotential(Node(-1)), ArgumentException);
    solver.SetArcBounds(Arc(0), 0, 3);                                            // capacity too small: infeasible, getters throw again
    CHECK(solver.Solve() == SolverStatus::Infeasible);
    CHECK_THROWS(solver.GetTotalCost(), InvalidOperationException);
    CHECK(!solver.Validate().IsValid);
    solver.SetPivotRule(PivotRule::CandidateList);
    solver.EnableOptimizedPivot(true);
    CHECK_THROWS(solver.Solve(), NotImplementedException);                        // :1694 (without the optimized wrapper the rule runs, defined as LEMON's)
    solver.EnableOptimizedPivot(false);
    solver.SetArcBounds(Arc(0), 0, 10);
    CHECK(solver.Solve() == SolverStatus::Optimal && solver.GetTotalCost() == 28);
    solver.SetPivotRule(PivotRule::AlteringList);
    CHECK(solver.Solve() == SolverStatus::Optimal && solver.GetTotalCost() == 28);
}

static void DimacsPath(const std::string& dir)
{
    auto ns = NetworkSimplex::FromDimacsFile(dir + "/grid_5x5.min");
    CHECK(ns->Graph().NodeCount() == 25 && ns->Graph().ArcCount() == 80);
    CHECK(ns->Solve() == SolverStatus::Optimal);
    CHECK(ns->GetTotalCost() == 27000);                                           // Resources/grid/grid_5x5.sol: `s 27000`
    ValidateSolution(*ns);
    auto big = NetworkSimplex::FromDimacsFile(dir + "/netgen_8_08a.min");
    CHECK(big->Solve() == SolverStatus::Optimal && big->GetTotalCost() == 142274536);   // Resources/netgen/netgen_8_08a.sol
    const std::string out = "/tmp/mcf_cpp_test.sol";
    big->SaveSolution(out, true, false);
    int64_t cost = 0; int32_t lines = 0, form = 0;
    CHECK(mcf_read_solution(out.c_str(), &cost, 0, nullptr, nullptr, nullptr, &lines, &form) == MCF_OK);
    CHECK(cost == 142274536 && form == 1 && lines > 0);
    CHECK_THROWS(NetworkSimplex::FromDimacsFile(dir + "/missing.min"), EngineException);
}

int main(int argc, char** argv)
{
    bool no_device = false; std::string dir = ".";
    for (int i = 1; i < argc; ++i) { if (!std::strcmp(argv[i], "--no-device")) no_device = true; else dir = argv[i]; }
    if (no_device) {
        // no CPU fallback: without an sm_100 GPU the constructor must fail loudly (mcf_create -> MCF_ERR_NO_DEVICE)
        GraphBuilder b; b.AddNodes(2).AddArc(0, 1);
        CHECK(mcf_device_count() == 0);
        CHECK_THROWS(NetworkSimplex s(b.Build()), EngineException);
        CHECK_THROWS(b.AddArc(0, 7), ArgumentException);
        mcf_dimacs* d = nullptr;
        CHECK(mcf_dimacs_open((dir + "/grid_5x5.min").c_str(), &d) == MCF_OK);    // the reader is host code
        int32_t n = 0, m = 0; mcf_dimacs_dims(d, &n, &m); CHECK(n == 25 && m == 80);
        mcf_dimacs_close(d);
    } else {
        SimpleTransportationProblem_SolvesCorrectly();
        MinimumCostCirculation_SolvesCorrectly();
        ExcessSupply_FeasibleWithGEQ();
        LowerBounds_RespectedInSolution();
        ComplementarySlackness_ValidatedCorrectly();
        OptimizedPivot_ProducesSameResults();
        ErrorBehaviour();
        DimacsPath(dir);
    }
    std::printf("%d checks, %d failed\n", g_checked, g_failed);
    return g_failed == 0 ? 0 : 1;
}
