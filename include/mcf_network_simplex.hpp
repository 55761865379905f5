// mcf_network_simplex.hpp - the reference's solver surface in C++, over the C ABI of libmcfgpu.so (include/mcfgpu.h).
//
// The reference (pdegenhardt/MinCostFlow) is compiled C#; no .NET toolchain exists where this repository is built, so the
// host side above the C ABI is written here in C++ with the reference's names, argument meaning and error behaviour:
//
//   GraphBuilder, CompactDigraph          src/MinCostFlow.Core/Lemon/Graphs/GraphBuilder.cs:10-104, CompactDigraph.cs
//   NetworkSimplex : IMinCostFlowSolver   src/MinCostFlow.Core/Lemon/Algorithms/NetworkSimplex.cs:119-210, :215-587
//   OptimizationConfig / Flags / Metrics  src/MinCostFlow.Core/Lemon/Algorithms/OptimizationTypes.cs:8-69
//   SolutionValidator, ValidationResult   src/MinCostFlow.Core/Lemon/Validation/SolutionValidator.cs:12-53, :348-359
//   DimacsReader, SolutionLoader          src/MinCostFlow.Problems/Loaders/DimacsReader.cs:25-147, SolutionLoader.cs:60-214
//
// ArgumentException -> mcf::ArgumentException (std::invalid_argument), InvalidOperationException -> mcf::InvalidOperationException
// (std::logic_error), NotImplementedException -> mcf::NotImplementedException, FormatException -> mcf::FormatException.
// bindings/CudaNetworkSimplex.cs is the same thing for the C# side (P/Invoke); mincostflow_b200/solver.py for Python.
// Header-only; link with -lmcfgpu.  There is no CPU fallback: the constructor throws mcf::EngineException without an sm_100 GPU.
#ifndef MCF_NETWORK_SIMPLEX_HPP
#define MCF_NETWORK_SIMPLEX_HPP

#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mcfgpu.h"

namespace mcf {

struct ArgumentException : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct InvalidOperationException : std::logic_error { using std::logic_error::logic_error; };
struct NotImplementedException : std::logic_error { using std::logic_error::logic_error; };
struct FormatException : std::runtime_error { using std::runtime_error::runtime_error; };
struct EngineException : std::runtime_error {                 // CUDA / engine failure reported through the ABI (no reference counterpart)
    int Code;
    EngineException(int code, const std::string& what) : std::runtime_error(what), Code(code) {}
};

struct Node { int Id = -1; Node() = default; explicit Node(int id) : Id(id) {} };      // Types/Node.cs
struct Arc { int Id = -1; Arc() = default; explicit Arc(int id) : Id(id) {} };         // Types/Arc.cs

enum class SolverStatus { NotSolved = 0, Optimal = 1, Infeasible = 2, Unbounded = 3, Unbalanced = 4 };                  // SolverStatus.cs:7-34
enum class PivotRule { FirstEligible = 0, BestEligible = 1, BlockSearch = 2, CandidateList = 3, AlteringList = 4 };      // PivotRule.cs:7-40
enum class SupplyType { Geq = 0, Leq = 1 };                                                                              // SupplyType.cs:7-17
enum OptimizationFlags : int {                                                                                           // OptimizationTypes.cs:8-20
    None = 0, AdaptiveBlockSize = 1, SmallBlocksForDense = 2, ReducedCostCaching = 4, CandidateListPivot = 8, HotColdSplitting = 16,
    EarlyTermination = 32
};

struct OptimizationConfig {                                                                                              // OptimizationTypes.cs:25-38
    int Flags = None;
    int MaxBlockSize = 100, MinBlockSize = 25, DenseNetworkThreshold = 10, ConsecutiveHitsBeforeAdapt = 3;
    double CandidateListRatio = 0.25, BlockSizeGrowthFactor = 1.2, BlockSizeShrinkFactor = 0.8;
    double LowHitRateThreshold = 0.05, HighHitRateThreshold = 0.3, MinBlockSizeRatio = 0.125;
};

using SolverMetrics = mcf_metrics;                                                                                       // OptimizationTypes.cs:43-69 + engine counters

// CompactDigraph.cs: arc ids in insertion order
class CompactDigraph {
public:
    int NodeCount() const { return n_; }
    int ArcCount() const { return (int)src_.size(); }
    Node AddNode() { return Node(n_++); }
    Arc AddArc(Node s, Node t)
    {
        if (!IsValid(s) || !IsValid(t)) throw ArgumentException("Invalid node");
        src_.push_back(s.Id); tgt_.push_back(t.Id);
        return Arc((int)src_.size() - 1);
    }
    bool IsValid(Node u) const { return u.Id >= 0 && u.Id < n_; }
    bool IsValid(Arc a) const { return a.Id >= 0 && a.Id < (int)src_.size(); }
    Node Source(Arc a) const { return Node(src_.at((size_t)a.Id)); }
    Node Target(Arc a) const { return Node(tgt_.at((size_t)a.Id)); }
    const std::vector<int32_t>& Sources() const { return src_; }
    const std::vector<int32_t>& Targets() const { return tgt_; }
    static CompactDigraph FromArrays(int n, std::vector<int32_t> src, std::vector<int32_t> tgt)
    {
        CompactDigraph g; g.n_ = n; g.src_ = std::move(src); g.tgt_ = std::move(tgt); return g;
    }
private:
    int n_ = 0;
    std::vector<int32_t> src_, tgt_;
};

// GraphBuilder.cs:10-104
class GraphBuilder {
public:
    GraphBuilder& AddNode() { return AddNode(next_id_++); }
    GraphBuilder& AddNode(int externalId)
    {
        if (map_.count(externalId)) throw ArgumentException("Node with ID " + std::to_string(externalId) + " already exists");
        map_[externalId] = graph_.AddNode();
        return *this;
    }
    GraphBuilder& AddNodes(int count) { for (int i = 0; i < count; ++i) AddNode(); return *this; }
    GraphBuilder& AddArc(int sourceId, int targetId)
    {
        auto s = map_.find(sourceId), t = map_.find(targetId);
        if (s == map_.end()) throw ArgumentException("Source node " + std::to_string(sourceId) + " not found");
        if (t == map_.end()) throw ArgumentException("Target node " + std::to_string(targetId) + " not found");
        graph_.AddArc(s->second, t->second);
        return *this;
    }
    Node GetNode(int externalId) const
    {
        auto it = map_.find(externalId);
        if (it == map_.end()) throw ArgumentException("Node " + std::to_string(externalId) + " not found");
        return it->second;
    }
    const CompactDigraph& Build() const { return graph_; }
    const std::map<int, Node>& NodeMap() const { return map_; }
private:
    CompactDigraph graph_;
    std::map<int, Node> map_;
    int next_id_ = 0;
};

// IMinCostFlowSolver.cs:8-34
class IMinCostFlowSolver {
public:
    virtual ~IMinCostFlowSolver() = default;
    virtual SolverStatus Solve() = 0;
    virtual SolverStatus Status() const = 0;
    virtual int64_t GetFlow(Arc arc) const = 0;
    virtual int64_t GetPotential(Node node) const = 0;
    virtual int64_t GetTotalCost() const = 0;
};

struct ValidationResult {                                                                                                // SolutionValidator.cs:348-359
    bool IsValid = false;
    mcf::SolverStatus SolverStatus = mcf::SolverStatus::NotSolved;
    mcf::SupplyType SupplyType = mcf::SupplyType::Geq;
    int64_t ObjectiveValue = 0, DualCost = 0;
    std::vector<std::string> Errors;
};

// NetworkSimplex.cs:119-587, backed by the CUDA engine
class NetworkSimplex : public IMinCostFlowSolver {
public:
    static constexpr int64_t INF = MCF_INF;                                                                              // NetworkSimplex.cs:127

    explicit NetworkSimplex(const CompactDigraph& graph, int device = 0) : graph_(graph)
    {
        const int n = graph.NodeCount(), m = graph.ArcCount();
        const int rc = mcf_create(n, m, graph.Sources().data(), graph.Targets().data(), &h_);
        if (rc != MCF_OK)
            throw EngineException(rc, rc == MCF_ERR_NO_DEVICE ? "no sm_100 CUDA device (the engine has no CPU fallback)" : "mcf_create failed");
        lower_.assign((size_t)m, 0); upper_.assign((size_t)m, INF); cost_.assign((size_t)m, 0); supply_.assign((size_t)n, 0);   // :615-617
        mcf_default_options(&opt_);
        opt_.device = device;
    }
    // DimacsReader.ReadFromFile + the constructor + every setter in one native call (mcf_create_from_dimacs)
    static std::unique_ptr<NetworkSimplex> FromDimacsFile(const std::string& path, int device = 0)
    {
        mcf_dimacs* d = nullptr;
        int rc = mcf_dimacs_open(path.c_str(), &d);
        if (rc == MCF_ERR_FORMAT) throw FormatException(mcf_io_last_error());
        if (rc != MCF_OK) throw EngineException(rc, mcf_io_last_error());
        int32_t n = 0, m = 0;
        mcf_dimacs_dims(d, &n, &m);
        std::vector<int32_t> s((size_t)m), t((size_t)m);
        mcf_dimacs_copy(d, s.data(), t.data(), nullptr, nullptr, nullptr, nullptr);
        auto graph = std::make_shared<CompactDigraph>(CompactDigraph::FromArrays(n, std::move(s), std::move(t)));
        std::unique_ptr<NetworkSimplex> ns(new NetworkSimplex(*graph, device));
        ns->owned_graph_ = graph;
        mcf_dimacs_copy(d, nullptr, nullptr, ns->lower_.data(), ns->upper_.data(), ns->cost_.data(), ns->supply_.data());
        mcf_dimacs_close(d);
        return ns;
    }
    ~NetworkSimplex() override { if (h_) mcf_destroy(h_); }
    NetworkSimplex(const NetworkSimplex&) = delete;
    NetworkSimplex& operator=(const NetworkSimplex&) = delete;

    // ---- fluent setters, NetworkSimplex.cs:153-210
    NetworkSimplex& SetArcBounds(Arc arc, int64_t lower, int64_t upper)
    {
        if (!graph_.IsValid(arc)) throw ArgumentException("Invalid arc");
        lower_[(size_t)arc.Id] = lower; upper_[(size_t)arc.Id] = upper; dirty_ = true;
        return *this;
    }
    NetworkSimplex& SetArcCost(Arc arc, int64_t cost)
    {
        if (!graph_.IsValid(arc)) throw ArgumentException("Invalid arc");
        cost_[(size_t)arc.Id] = cost; dirty_ = true;
        return *this;
    }
    NetworkSimplex& SetNodeSupply(Node node, int64_t supply)
    {
        if (!graph_.IsValid(node)) throw ArgumentException("Invalid node");
        supply_[(size_t)node.Id] = supply; dirty_ = true;
        return *this;
    }
    NetworkSimplex& SetSupplyType(mcf::SupplyType type) { opt_.supply_type = (int)type; return *this; }
    NetworkSimplex& SetPivotRule(PivotRule rule) { opt_.pivot_rule = (int)rule; return *this; }
    // :532; simdWidth = Vector<long>.Count of the host whose optimized Block Search sequence is to be reproduced (4 = x64 AVX2)
    void EnableOptimizedPivot(bool enable = true, int simdWidth = 4) { opt_.optimized_pivot = enable ? 1 : 0; opt_.simd_width = simdWidth; }
    // SURVEY.md 8f-3 (README.md:17-18 roadmap): re-Solve() after SetArcCost edits starts from the previous optimal basis (mcf_options.warm_start)
    void EnableWarmStart(bool enable = true) { opt_.warm_start = enable ? 1 : 0; }
    void EnableOptimizations(int flags) { opt_.config.flags = flags; }                                                  // :549-552
    void SetOptimizationConfig(const OptimizationConfig& c)                                                             // :557-561
    {
        mcf_optimization_config& o = opt_.config;
        o.flags = c.Flags; o.max_block_size = c.MaxBlockSize; o.min_block_size = c.MinBlockSize; o.dense_network_threshold = c.DenseNetworkThreshold;
        o.consecutive_hits_before_adapt = c.ConsecutiveHitsBeforeAdapt; o.candidate_list_ratio = c.CandidateListRatio;
        o.block_size_growth_factor = c.BlockSizeGrowthFactor; o.block_size_shrink_factor = c.BlockSizeShrinkFactor;
        o.low_hit_rate_threshold = c.LowHitRateThreshold; o.high_hit_rate_threshold = c.HighHitRateThreshold; o.min_block_size_ratio = c.MinBlockSizeRatio;
        opt_.auto_configuration = 0;                                                                                    // :560
    }
    void SetAutoConfiguration(bool enable) { opt_.auto_configuration = enable ? 1 : 0; }                                // :567
    mcf_options& EngineOptions() { return opt_; }                                                                       // device, engine, stop_after_pivots ...

    // ---- Solve(), :215-411
    SolverStatus Solve() override
    {
        // CandidateList / AlteringList: the reference throws (:884); here they run, defined as LEMON's (network_simplex.h:413-635).
        // The optimized wrapper still knows only the first three rules (:1689-1695).
        if (opt_.pivot_rule > (int)PivotRule::BlockSearch && opt_.optimized_pivot) throw NotImplementedException("Optimized pivot rule not implemented");
        Push();
        int32_t st = 0;
        Check(mcf_solve(h_, &st));
        return (SolverStatus)st;
    }
    SolverStatus Status() const override { int32_t st = 0; Check(mcf_get_status(h_, &st)); return (SolverStatus)st; }  // :470
    mcf::SupplyType SupplyType() const { return (mcf::SupplyType)opt_.supply_type; }

    // ---- results, :416-465 (InvalidOperationException unless Optimal, ArgumentException for a bad id)
    int64_t GetFlow(Arc arc) const override { int64_t v = 0; Check(mcf_get_flow(h_, arc.Id, &v)); return v; }
    int64_t GetPotential(Node node) const override { int64_t v = 0; Check(mcf_get_potential(h_, node.Id, &v)); return v; }
    int64_t GetTotalCost() const override { int64_t v = 0; Check(mcf_get_total_cost(h_, &v)); return v; }
    std::vector<int64_t> Flows() const { std::vector<int64_t> f((size_t)graph_.ArcCount()); if (!f.empty()) Check(mcf_get_flows(h_, f.data())); return f; }
    std::vector<int64_t> Potentials() const { std::vector<int64_t> p((size_t)graph_.NodeCount()); if (!p.empty()) Check(mcf_get_potentials(h_, p.data())); return p; }
    // ---- :480-527 (what SolutionValidator reads)
    int64_t GetNodeSupply(Node node) const { int64_t v = 0; Check(mcf_get_node_supply(h_, node.Id, &v)); return v; }
    int64_t GetArcCost(Arc arc) const { int64_t v = 0; Check(mcf_get_arc_cost(h_, arc.Id, &v)); return v; }
    int64_t GetArcLowerBound(Arc arc) const { int64_t v = 0; Check(mcf_get_arc_lower_bound(h_, arc.Id, &v)); return v; }
    int64_t GetArcUpperBound(Arc arc) const { int64_t v = 0; Check(mcf_get_arc_upper_bound(h_, arc.Id, &v)); return v; }
    SolverMetrics GetMetrics() const { SolverMetrics m; Check(mcf_get_metrics(h_, &m)); return m; }                     // :584

    // SolutionValidator(graph, solver).Validate(), SolutionValidator.cs:20-53, on the device
    ValidationResult Validate() const
    {
        ValidationResult r;
        r.SolverStatus = Status(); r.SupplyType = SupplyType();
        if (r.SolverStatus != SolverStatus::Optimal) { r.Errors.push_back("Solver status is not optimal"); return r; } // :24-33
        int32_t bits = 0; int64_t primal = 0, dual = 0;
        Check(mcf_validate(h_, &bits, &primal, &dual));
        r.ObjectiveValue = primal; r.DualCost = dual;
        static const char* const what[6] = {"Flow conservation violated", "Capacity bounds violated", "Complementary slackness violated",
                                            "Dual feasibility violated", "Objective value mismatch", "Strong duality violated"};
        for (int b = 0; b < 6; ++b) if (bits & (1 << b)) r.Errors.push_back(what[b]);
        r.IsValid = bits == 0;
        return r;
    }
    // SolutionLoader.SaveToFile, SolutionLoader.cs:186-214
    void SaveSolution(const std::string& path, bool byEndpoints = false, bool withPotentials = false) const
    {
        const int rc = mcf_write_solution(h_, path.c_str(), byEndpoints ? 1 : 0, withPotentials ? 1 : 0);
        if (rc == MCF_ERR_NOT_OPTIMAL) throw InvalidOperationException("Solution not optimal");
        if (rc != MCF_OK) throw EngineException(rc, mcf_io_last_error());
    }
    const CompactDigraph& Graph() const { return graph_; }

private:
    void Push()
    {
        if (dirty_) {
            Check(mcf_set_arcs(h_, lower_.data(), upper_.data(), cost_.data()));
            Check(mcf_set_supply(h_, supply_.data()));
            dirty_ = false;
        }
        Check(mcf_set_options(h_, &opt_));
    }
    void Check(int rc) const
    {
        if (rc == MCF_OK) return;
        const char* e = mcf_last_error(h_);
        const std::string msg = e && *e ? e : "";
        if (rc == MCF_ERR_INVALID_ARGUMENT) throw ArgumentException(msg.empty() ? "invalid argument" : msg);
        if (rc == MCF_ERR_NOT_OPTIMAL) throw InvalidOperationException(msg.empty() ? "Solution not optimal" : msg);
        throw EngineException(rc, msg);
    }

    const CompactDigraph& graph_;
    std::shared_ptr<CompactDigraph> owned_graph_;
    mcf_handle* h_ = nullptr;
    mcf_options opt_{};
    std::vector<int64_t> lower_, upper_, cost_, supply_;
    bool dirty_ = true;
};

}  // namespace mcf
#endif  // MCF_NETWORK_SIMPLEX_HPP
