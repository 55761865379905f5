// mcf_capi.cu - host layer of libmcfgpu.so: the C ABI declared in include/mcfgpu.h.
//
// Mirrors the host-side part of NetworkSimplex.Solve() (NS.cs = src/MinCostFlow.Core/Lemon/Algorithms/
// NetworkSimplex.cs in the reference): CheckBounds (:624), TransformToStandardForm (:636), the automatic
// configuration (ProblemAnalyzer.cs:21-62 -> OptimizationSelector.cs:14-96), Initialize (:671-845), the choice of
// pricing rule (:847-886), and after the pivot loop the status / result plumbing (:359-410).  The pivot loop itself
// runs in mcf_kernels.cu.  There is no CPU solver in this library.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/mcfgpu.h"
#include "mcf_device.cuh"

extern "C" int mcfk_pivot_smem_bytes();
extern "C" int mcfk_max_grid(int device, int* sm_count);
extern "C" int mcfk_launch_pivot(const mcf::Params* p, int grid, cudaStream_t stream);
extern "C" int mcfk_launch_l2_read(const void* buf, size_t bytes, long long* sink, int sms, cudaStream_t stream);
extern "C" int mcfk_launch_price_sweep(const mcf::Params* p, mcf::PriceRec* out, int grid, cudaStream_t stream);
extern "C" int mcfk_launch_validate(const mcf::ValidateParams* v, int sms, cudaStream_t stream);
extern "C" cudaError_t mcfk_device_props(int device, cudaDeviceProp* out);
extern "C" size_t mcfk_team_smem_bytes(int slice, int wide, int spill);
extern "C" int mcfk_team_max_slice(int device, int wide, int spill);
extern "C" int mcfk_team_max_ctas(int device, int slice, int wide, int spill);
extern "C" int mcfk_launch_team(const mcf::TeamParams* p, cudaStream_t stream);
extern "C" void mcfk_team_replicas(int* ent, int* cyc);

namespace {

using clk = std::chrono::steady_clock;
inline double us_since(clk::time_point t0) { return std::chrono::duration<double, std::micro>(clk::now() - t0).count(); }

constexpr int64_t kInf = std::numeric_limits<int64_t>::max() / 2;   // NS.cs:127

struct Characteristics {            // the subset of ProblemCharacteristics the selector reads
    int node_count = 0, arc_count = 0;
    double density = 0, degree_cv = 0, cost_cv = 0;
    bool is_dense = false, is_sparse = false, has_uniform_costs = false;
    int source_count = 0, sink_count = 0, transshipment_count = 0;
    int64_t max_abs_supply = 0;
    int detected_type = 0;          // 0 General, 1 Circulation, 2 Assignment, 3 Transportation, 4 Transshipment
};

template <class T> struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t count) {
        if (count <= cap && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) cap = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct mcf_handle {
    int n = 0, m = 0;
    std::vector<int32_t> source, target;
    std::vector<int64_t> lower, upper, cost, supply, orig_lower;      // NS.cs:42-48 (mutated by Solve like the reference)
    mcf_options opt{};
    int status = MCF_NOT_SOLVED;
    bool solved_once = false;
    bool stopped_early = false;                                       // the last solve ended at opt.stop_after_pivots
    std::vector<int64_t> flow, pi;                                    // results, host side
    int64_t total_cost = 0;
    mcf_metrics metrics{};
    std::string err;
    // device
    int device_bound = -1;
    cudaStream_t stream = nullptr;
    DevBuf<int> d_src, d_tgt, d_cost, d_state, d_in, d_sz, d_parent, d_pd;
    DevBuf<long long> d_flow, d_upper, d_lower, d_pi, d_rc;
    DevBuf<mcf::PriceRec> d_part;
    DevBuf<mcf::CycEnt> d_list;
    DevBuf<int> d_scratch;                                            // flat engine: stem scratch (mcf_device.cuh)
    DevBuf<int> d_cand, d_cand_scratch;                               // Candidate List / Altering List rules (mcf_device.cuh)
    DevBuf<long long> d_cand_cost;
    DevBuf<mcf::Ctl> d_ctl;
    DevBuf<unsigned char> d_flush;
    DevBuf<long long> d_val;                                          // validator scratch
    const long long* d_pi_final = nullptr;                            // potentials of the last solve, device side
    int supply_type_solved = 0;
    // team engine (mcf_team.cu)
    DevBuf<mcf::NodeRec> d_node;
    DevBuf<int4> d_mail;                                              // ent0 | prc | late | cyc | stemseg
    DevBuf<unsigned> d_done;
    DevBuf<long long> d_piout, d_flg, d_upg;
    std::vector<mcf::NodeRec> h_node;
    // warm start (mcf_options.warm_start): what identifies the problem the device arrays of the last optimal solve belong to
    bool warm_valid = false;
    int warm_supply_type = 0, warm_engine = 0;
    std::vector<int64_t> warm_upper, warm_supply, warm_lower, cur_upper, cur_supply;   // standard form (after the lower-bound shift) + the shift itself
    // host staging for the initial basis
    std::vector<int> h_depth;
    std::vector<int> h_src, h_tgt, h_cost, h_state, h_in, h_sz, h_parent, h_pd;
    std::vector<long long> h_flow, h_upper, h_pi;
};

namespace {

// The persisting-L2 set-aside is a device-wide limit, and changing it waits for the kernels that are running on the device:
// with several solves side by side (mcf_solve_batch_concurrent) a per-solve cudaDeviceSetLimit serialised them.  It is
// therefore only ever raised, under a lock, and the batch entry point raises it once before its workers start.
std::mutex g_persist_mu;
size_t g_persist_limit[64] = {0};
void ensure_persisting_l2(int device, size_t want)
{
    if (device < 0 || device >= 64) return;
    std::lock_guard<std::mutex> lk(g_persist_mu);
    if (g_persist_limit[device] >= want) return;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) g_persist_limit[device] = want;
    else cudaGetLastError();
}

int fail(mcf_handle* h, int code, const char* fmt, ...)
{
    if (h) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
        h->err = buf;
    }
    return code;
}

#define CUDA_TRY(h, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return fail((h), e__ == cudaErrorMemoryAllocation ? MCF_ERR_OUT_OF_MEMORY : MCF_ERR_CUDA,       \
                        "%s failed: %s", #call, cudaGetErrorString(e__));                                   \
    } while (0)

struct EventPair {                  // two timing events that are destroyed on every exit path
    cudaEvent_t a = nullptr, b = nullptr;
    cudaError_t create() { cudaError_t e = cudaEventCreate(&a); return e != cudaSuccess ? e : cudaEventCreate(&b); }
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

bool device_ok(int dev)
{
    cudaDeviceProp p;
    if (mcfk_device_props(dev, &p) != cudaSuccess) return false;
    return p.major == 10;           // sm_100a code only
}

// ProblemAnalyzer.Analyze (ProblemAnalyzer.cs:21-62); only what OptimizationSelector consumes.
Characteristics analyze(const mcf_handle& h)
{
    Characteristics ch;
    const int n = h.n, m = h.m;
    ch.node_count = n; ch.arc_count = m;
    const int64_t max_possible = (int64_t)n * (n - 1);
    ch.density = max_possible > 0 ? (double)m / (double)max_possible : 0;
    std::vector<int> outd(n > 0 ? n : 1, 0), ind(n > 0 ? n : 1, 0);
    for (int e = 0; e < m; ++e) { outd[h.source[e]]++; ind[h.target[e]]++; }
    int total = 0;
    for (int i = 0; i < n; ++i) total += outd[i] + ind[i];
    const double avg = n > 0 ? (double)total / n : 0;
    double var = 0;
    if (n > 0) { for (int i = 0; i < n; ++i) { const double d = (outd[i] + ind[i]) - avg; var += d * d; } var /= n; }   // :86-95
    ch.degree_cv = avg > 0 ? std::sqrt(var) / avg : 0;
    for (int i = 0; i < n; ++i) {
        const int64_t s = h.supply[i];
        if (s > 0) ch.source_count++; else if (s < 0) ch.sink_count++; else ch.transshipment_count++;
        const int64_t a = s < 0 ? -s : s; if (a > ch.max_abs_supply) ch.max_abs_supply = a;
    }
    if (m == 0) ch.has_uniform_costs = true;
    else {
        int64_t tot = 0;
        for (int i = 0; i < m; ++i) tot += h.cost[i];
        const double ac = (double)tot / m;
        double v = 0;
        for (int i = 0; i < m; ++i) { const double d = h.cost[i] - ac; v += d * d; }
        v /= m;
        ch.cost_cv = std::fabs(ac) > 0 ? std::sqrt(v) / std::fabs(ac) : 0;
        ch.has_uniform_costs = ch.cost_cv < 0.01;
    }
    if (ch.source_count == 0 && ch.sink_count == 0) ch.detected_type = 1;
    else {
        int only_out = 0, only_in = 0;
        for (int i = 0; i < n; ++i) {
            if (outd[i] > 0 && ind[i] == 0) only_out++; else if (outd[i] == 0 && ind[i] > 0) only_in++;
        }
        const bool bip = ((double)(only_out + only_in) / n) > 0.8;
        if (bip && ch.max_abs_supply == 1 && ch.source_count == ch.sink_count) ch.detected_type = 2;
        else if (bip && ch.transshipment_count == 0) ch.detected_type = 3;
        else if (ch.transshipment_count > 0) ch.detected_type = 4;
    }
    ch.is_dense = ch.density > 0.01 || m > 10000;
    ch.is_sparse = ch.density < 0.005;
    return ch;
}

void default_config(mcf_optimization_config* c)
{   // OptimizationTypes.cs:25-38
    c->flags = 0; c->max_block_size = 100; c->min_block_size = 25; c->dense_network_threshold = 10000;
    c->consecutive_hits_before_adapt = 3; c->reserved0 = 0; c->candidate_list_ratio = 0.1;
    c->block_size_growth_factor = 1.2; c->block_size_shrink_factor = 0.8; c->low_hit_rate_threshold = 0.05;
    c->high_hit_rate_threshold = 0.3; c->min_block_size_ratio = 0.125;
}

// OptimizationSelector.SelectConfiguration (OptimizationSelector.cs:14-96)
mcf_optimization_config select_config(const Characteristics& ch)
{
    mcf_optimization_config c; default_config(&c);
    int flags = 0;
    if (ch.is_dense) { flags |= MCF_FLAG_SMALL_BLOCKS_FOR_DENSE; c.min_block_size = 10; c.max_block_size = 50; c.dense_network_threshold = 5000; }
    if (ch.degree_cv > 0.5) {
        flags |= MCF_FLAG_ADAPTIVE_BLOCK_SIZE;
        c.block_size_growth_factor = 1.3; c.block_size_shrink_factor = 0.7; c.consecutive_hits_before_adapt = 2;
    } else if (ch.degree_cv > 0.3) flags |= MCF_FLAG_ADAPTIVE_BLOCK_SIZE;
    if (ch.is_sparse && ch.arc_count < 50000) flags |= MCF_FLAG_REDUCED_COST_CACHING;
    const bool transp = ch.detected_type == 2 || ch.detected_type == 3;
    if (ch.arc_count >= 1000 && ((ch.is_sparse && ch.arc_count > 5000) || ch.has_uniform_costs || transp)) {
        flags |= MCF_FLAG_CANDIDATE_LIST_PIVOT;         // set but never read by the solver (inert)
        c.candidate_list_ratio = ch.has_uniform_costs ? 0.2 : (ch.arc_count > 100000 ? 0.05 : 0.1);
    }
    if (ch.node_count > 5000 && ch.degree_cv > 1.0) flags |= MCF_FLAG_HOT_COLD_SPLITTING;
    if (transp) flags |= MCF_FLAG_EARLY_TERMINATION;
    c.low_hit_rate_threshold = ch.arc_count > 10000 ? 0.03 : 0.05;
    c.high_hit_rate_threshold = ch.arc_count > 10000 ? 0.25 : 0.3;
    c.min_block_size_ratio = ch.arc_count > 100000 ? 0.0625 : (ch.arc_count > 10000 ? 0.125 : 0.25);
    c.flags = flags;
    return c;
}

int bind_device(mcf_handle* h)
{
    const int dev = h->opt.device;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return fail(h, MCF_ERR_NO_DEVICE, "no CUDA device visible");
    if (dev < 0 || dev >= count || !device_ok(dev)) return fail(h, MCF_ERR_NO_DEVICE, "device %d is not an sm_100 GPU", dev);
    CUDA_TRY(h, cudaSetDevice(dev));
    if (h->device_bound != dev) {
        if (h->stream) { cudaStreamDestroy(h->stream); h->stream = nullptr; }
        h->d_src.release(); h->d_tgt.release(); h->d_cost.release(); h->d_state.release(); h->d_in.release(); h->d_sz.release();
        h->d_parent.release(); h->d_pd.release(); h->d_flow.release(); h->d_upper.release(); h->d_lower.release(); h->d_pi.release();
        h->d_rc.release(); h->d_part.release(); h->d_list.release(); h->d_scratch.release(); h->d_ctl.release(); h->d_flush.release();
        h->d_node.release(); h->d_mail.release(); h->d_done.release(); h->d_piout.release(); h->d_flg.release(); h->d_upg.release(); h->d_cand.release(); h->d_cand_scratch.release(); h->d_cand_cost.release(); h->d_val.release(); h->d_pi_final = nullptr;
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->device_bound = dev;
    }
    return MCF_OK;
}

// Initialize() + InitializeGEQ / InitializeLEQ (NS.cs:671-845) into the staging arrays, in the engine's layout:
// the star tree rooted at node n, labelled in[u] = u + 1 (the reference's initial thread order root,0,1,..,n-1).
void build_initial_basis(mcf_handle* h, int64_t art_cost)
{
    const int n = h->n, m = h->m, S = m + n, A = m + 2 * n, root = n;
    h->h_src.assign(S, 0); h->h_tgt.assign(S, 0); h->h_cost.assign(S, 0);
    h->h_state.assign(A, mcf::STATE_LOWER); h->h_flow.assign(A, 0); h->h_upper.assign(A, kInf);
    h->h_in.resize(n + 1); h->h_sz.resize(n + 1); h->h_parent.resize(n + 1); h->h_pd.resize(n + 1); h->h_pi.assign(n + 1, 0);
    for (int e = 0; e < m; ++e) {
        h->h_src[e] = h->source[e]; h->h_tgt[e] = h->target[e]; h->h_cost[e] = (int)h->cost[e];
        h->h_upper[e] = h->upper[e];
    }
    const bool geq = h->opt.supply_type == MCF_GEQ;
    int f = S;
    for (int u = 0, e = m; u < n; ++u, ++e) {
        h->h_in[u] = u + 1; h->h_sz[u] = 1; h->h_parent[u] = root;
        const int64_t s = h->supply[u];
        if (geq) { h->h_src[e] = root; h->h_tgt[e] = u; } else { h->h_src[e] = u; h->h_tgt[e] = root; }
        const bool plain = geq ? s <= 0 : s >= 0;
        if (plain) {                                        // NS.cs:744-754 / :811-821
            h->h_pd[u] = e * 2 + (geq ? 0 : 1);
            h->h_pi[u] = 0; h->h_flow[e] = geq ? -s : s; h->h_state[e] = mcf::STATE_TREE;
        } else {                                            // NS.cs:755-771 / :822-838: artificial arc f carries the supply
            h->h_pd[u] = f * 2 + (geq ? 1 : 0);
            h->h_pi[u] = geq ? -art_cost : art_cost;
            h->h_flow[f] = geq ? s : -s; h->h_state[f] = mcf::STATE_TREE;
            h->h_state[e] = mcf::STATE_LOWER; h->h_flow[e] = 0;
            ++f;
        }
    }
    h->h_in[root] = 0; h->h_sz[root] = n + 1; h->h_parent[root] = -1; h->h_pd[root] = -2;
    h->h_depth.assign(n + 1, 1); h->h_depth[root] = 0;
}

// Warm start: the staging arrays from the basis the previous optimal solve left on the device.  The basis is
// the set of arcs in STATE_TREE plus the flows; parent / pred / direction follow from a traversal from the root, the interval
// labels are unobservable (any depth-first order is valid), and the potentials are recomputed for the CURRENT costs along the
// tree: pi[u] = pi[parent] - cost when the pred arc leaves u, + cost when it enters u (reduced cost 0 on every tree arc,
// NS.cs:1185-1209).  Returns false (caller starts cold) when the arrays do not describe a spanning tree.
bool build_warm_basis(mcf_handle* h, int64_t art_cost)
{
    const int n = h->n, m = h->m, S = m + n, A = m + 2 * n, root = n;
    std::vector<int> w_state(A);
    std::vector<long long> w_flow(A);
    if (cudaMemcpy(w_state.data(), h->d_state.p, (size_t)A * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaMemcpy(w_flow.data(), h->d_flow.p, (size_t)A * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return false; }
    h->metrics.d2h_bytes += (int64_t)A * 12;
    build_initial_basis(h, art_cost);                       // arc arrays (endpoints, new costs, capacities) and the second artificial arcs' owners
    // endpoints of the second set of artificial arcs [S, A): the cold basis hands them out in node order (NS.cs:755-771 / :822-838)
    const bool geq = h->opt.supply_type == MCF_GEQ;
    std::vector<int> f_node(A - S, -1);
    for (int u = 0; u < n; ++u) if ((h->h_pd[u] >> 1) >= S) f_node[(h->h_pd[u] >> 1) - S] = u;
    auto ends = [&](int a, int& s, int& t) {
        if (a < S) { s = h->h_src[a]; t = h->h_tgt[a]; return true; }
        const int u = f_node[a - S];
        if (u < 0) return false;
        if (geq) { s = u; t = root; } else { s = root; t = u; }
        return true;
    };
    // adjacency of the tree arcs
    std::vector<int> deg(n + 2, 0), tree_arcs;
    tree_arcs.reserve(n);
    for (int a = 0; a < A; ++a) if (w_state[a] == mcf::STATE_TREE) {
        int s, t; if (!ends(a, s, t)) return false;
        tree_arcs.push_back(a); ++deg[s + 1]; ++deg[t + 1];
    }
    if ((int)tree_arcs.size() != n) return false;
    for (int u = 0; u <= n; ++u) deg[u + 1] += deg[u];
    std::vector<int> adj(2 * (size_t)n), fill(deg.begin(), deg.end() - 1);
    for (int a : tree_arcs) { int s, t; ends(a, s, t); adj[fill[s]++] = a; adj[fill[t]++] = a; }
    // depth-first traversal from the root: labels, parent, pred word, depth, potentials
    std::vector<int> order; order.reserve(n + 1);
    std::vector<int> stack; stack.reserve(n + 1);
    std::vector<char> seen(n + 1, 0);
    h->h_parent[root] = -1; h->h_pd[root] = -2; h->h_depth[root] = 0; h->h_pi[root] = 0;
    stack.push_back(root); seen[root] = 1;
    while (!stack.empty()) {
        const int u = stack.back(); stack.pop_back();
        h->h_in[u] = (int)order.size(); order.push_back(u);
        for (int k = deg[u + 1] - 1; k >= deg[u]; --k) {            // pushed in reverse: children are visited in adjacency order
            const int a = adj[k];
            int s, t; ends(a, s, t);
            const int v = s == u ? t : s;
            if (seen[v]) continue;
            seen[v] = 1;
            const bool dir_up = s == v;                             // the arc leaves the child towards its parent
            const long long c = a < S ? (long long)h->h_cost[a] : (long long)art_cost;
            h->h_parent[v] = u; h->h_pd[v] = a * 2 + (dir_up ? 1 : 0); h->h_depth[v] = h->h_depth[u] + 1;
            h->h_pi[v] = dir_up ? h->h_pi[u] - c : h->h_pi[u] + c;
            stack.push_back(v);
        }
    }
    if ((int)order.size() != n + 1) return false;
    for (int u = 0; u <= n; ++u) h->h_sz[u] = 1;
    for (int i = n; i > 0; --i) { const int u = order[i]; h->h_sz[h->h_parent[u]] += h->h_sz[u]; }
    for (int a = 0; a < A; ++a) { h->h_state[a] = w_state[a]; h->h_flow[a] = w_flow[a]; }
    for (int e = 0; e < m; ++e) if (h->orig_lower[e] != 0) h->h_flow[e] -= h->orig_lower[e];      // the epilogue added the lower bounds (NS.cs:375-388)
    return true;
}

int upload_basis(mcf_handle* h, bool need_cache, int grid)
{
    const int n = h->n, m = h->m, S = m + n, A = m + 2 * n;
    CUDA_TRY(h, h->d_src.ensure(S + 4)); CUDA_TRY(h, h->d_tgt.ensure(S + 4)); CUDA_TRY(h, h->d_cost.ensure(S + 4));
    CUDA_TRY(h, h->d_state.ensure(A + 4)); CUDA_TRY(h, h->d_flow.ensure(A)); CUDA_TRY(h, h->d_upper.ensure(A));
    CUDA_TRY(h, h->d_in.ensure(n + 1)); CUDA_TRY(h, h->d_sz.ensure(n + 1)); CUDA_TRY(h, h->d_parent.ensure(n + 1));
    CUDA_TRY(h, h->d_pd.ensure(n + 1)); CUDA_TRY(h, h->d_pi.ensure(n + 1));
    CUDA_TRY(h, h->d_part.ensure((size_t)2 * grid)); CUDA_TRY(h, h->d_list.ensure((size_t)n + 1)); CUDA_TRY(h, h->d_scratch.ensure((size_t)6 * (n + 1))); CUDA_TRY(h, h->d_ctl.ensure(1));
    if (need_cache) { CUDA_TRY(h, h->d_rc.ensure(S)); CUDA_TRY(h, cudaMemsetAsync(h->d_rc.p, 0, (size_t)S * 8, h->stream)); }
    cudaStream_t st = h->stream;
    int64_t bytes = 0;
    auto up = [&](void* d, const void* s, size_t b) { bytes += (int64_t)b; return cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, st); };
    CUDA_TRY(h, up(h->d_src.p, h->h_src.data(), (size_t)S * 4)); CUDA_TRY(h, up(h->d_tgt.p, h->h_tgt.data(), (size_t)S * 4));
    CUDA_TRY(h, up(h->d_cost.p, h->h_cost.data(), (size_t)S * 4)); CUDA_TRY(h, up(h->d_state.p, h->h_state.data(), (size_t)A * 4));
    CUDA_TRY(h, up(h->d_flow.p, h->h_flow.data(), (size_t)A * 8)); CUDA_TRY(h, up(h->d_upper.p, h->h_upper.data(), (size_t)A * 8));
    CUDA_TRY(h, up(h->d_in.p, h->h_in.data(), (size_t)(n + 1) * 4)); CUDA_TRY(h, up(h->d_sz.p, h->h_sz.data(), (size_t)(n + 1) * 4));
    CUDA_TRY(h, up(h->d_parent.p, h->h_parent.data(), (size_t)(n + 1) * 4)); CUDA_TRY(h, up(h->d_pd.p, h->h_pd.data(), (size_t)(n + 1) * 4));
    CUDA_TRY(h, up(h->d_pi.p, h->h_pi.data(), (size_t)(n + 1) * 8));
    CUDA_TRY(h, cudaMemsetAsync(h->d_ctl.p, 0, sizeof(mcf::Ctl), st));
    h->metrics.h2d_bytes = bytes;
    return MCF_OK;
}

void fill_params(mcf_handle* h, mcf::Params* P)
{
    std::memset(P, 0, sizeof(*P));
    P->n = h->n; P->m = h->m; P->S = h->m + h->n; P->A = h->m + 2 * h->n;
    P->src = h->d_src.p; P->tgt = h->d_tgt.p; P->cost = h->d_cost.p; P->state = h->d_state.p;
    P->flow = h->d_flow.p; P->upper = h->d_upper.p; P->orig_lower = nullptr; P->rc_cache = nullptr;
    P->in = h->d_in.p; P->sz = h->d_sz.p; P->parent = h->d_parent.p; P->pd = h->d_pd.p; P->pi = h->d_pi.p;
    P->part = h->d_part.p; P->list = h->d_list.p; P->list_cap = h->n + 1; P->stem_scratch = h->d_scratch.p; P->ctl = h->d_ctl.p;
}

int choose_grid(mcf_handle* h, int* sms_out)
{
    int sms = 0;
    const int maxg = mcfk_max_grid(h->opt.device, &sms);
    if (maxg <= 0) return fail(h, MCF_ERR_CUDA, "cooperative occupancy query failed (%d): %s", maxg, cudaGetErrorString(cudaGetLastError()));
    int g = maxg;
    if (h->opt.max_ctas > 0 && h->opt.max_ctas < g) g = h->opt.max_ctas;
    else if (h->opt.max_ctas <= 0) {
        // small instances: fewer CTAs make every grid barrier cheaper; keep >= 2048 nodes or 16K arcs per CTA
        const int64_t S = (int64_t)h->m + h->n;
        int64_t want = std::max<int64_t>((h->n + 2047) / 2048, (S + 16383) / 16384);
        if (want < 4) want = 4;
        if (want < g) g = (int)want;
    }
    if (sms_out) *sms_out = sms;
    return g;
}


// Team engine (mcf_team.cu): the first `pricers` CTAs price, the others own node slices that stay in shared memory.
// Returns the team size, 0 when the instance does not fit (caller falls back to the flat engine).
int choose_team(mcf_handle* h, int wide, int spill, int* slice_out, int* pricers_out)
{
    cudaDeviceProp prop;
    if (mcfk_device_props(h->opt.device, &prop) != cudaSuccess) return 0;
    const int max_slice = mcfk_team_max_slice(h->opt.device, wide, spill);
    if (max_slice <= 0) return 0;
    int limit = prop.multiProcessorCount;
    if (limit > mcf::kTeamMax) limit = mcf::kTeamMax;
    if (h->opt.max_ctas > 1 && h->opt.max_ctas < limit) limit = h->opt.max_ctas;
    if (limit < 2) return 0;
    const long long nodes = (long long)h->n + 1;
    const long long S = (long long)h->m + h->n;
    const long long need = (nodes + max_slice - 1) / max_slice;                 // owners the slices need at least
    if (need > limit - 1) return 0;
    // pricers: one SM is gather-bound on a whole block; ~192 arcs of the first block per pricer (each arc end costs two L1TEX line look-ups)
    const long long B = (long long)std::sqrt((double)S);
    long long pricers = h->opt.lookahead_blocks > 0 ? h->opt.lookahead_blocks : (B + 191) / 192;
    if (pricers < 1) pricers = 1;
    if (pricers > mcf::kMaxPricers) pricers = mcf::kMaxPricers;
    if (pricers > limit - need) pricers = limit - need;
    // owners: ~1024 nodes each when SMs are to spare (fewer CTAs make every hop cheaper), never fewer than needed
    long long owners = (nodes + 1023) / 1024;
    if (owners < need) owners = need;
    if (owners > limit - pricers) owners = limit - pricers;
    long long slice = (nodes + owners - 1) / owners;
    slice = (slice + 7) & ~7LL;
    if (slice > max_slice) return 0;
    const int team = (int)(owners + pricers);
    if (mcfk_team_max_ctas(h->opt.device, (int)slice, wide, spill) < team) return 0;
    *slice_out = (int)slice; *pricers_out = (int)pricers;
    return team;
}

int upload_team(mcf_handle* h, int team, int pricers, int slice, int wide, int spill, mcf::TeamParams* P)
{
    const int n = h->n, m = h->m, S = m + n, A = m + 2 * n;
    CUDA_TRY(h, h->d_src.ensure(S + 4)); CUDA_TRY(h, h->d_tgt.ensure(S + 4)); CUDA_TRY(h, h->d_cost.ensure(S + 4));
    CUDA_TRY(h, h->d_state.ensure(A + 4)); CUDA_TRY(h, h->d_flow.ensure(A)); CUDA_TRY(h, h->d_upper.ensure(A));
    CUDA_TRY(h, h->d_sz.ensure(n + 1)); CUDA_TRY(h, h->d_pd.ensure(n + 1));
    // node mirror {pi, depth} and the dense label array live in ONE allocation so that a single L2 access-policy window can
    // keep both resident (the arc arrays stream through L2 and would otherwise evict them between two uses of a record)
    const size_t mirror_recs = (size_t)(n + 1) + ((size_t)(n + 1) * 4 + sizeof(mcf::NodeRec) - 1) / sizeof(mcf::NodeRec);
    CUDA_TRY(h, h->d_node.ensure(mirror_recs));
    int* const d_in_g = reinterpret_cast<int*>(h->d_node.p + (n + 1));
    CUDA_TRY(h, h->d_piout.ensure(n)); CUDA_TRY(h, h->d_ctl.ensure(1)); CUDA_TRY(h, h->d_done.ensure((size_t)(team + pricers) * 32));     // DONE flags of the owners + GATHERED flags of the pricers
    int rep_ent = 1, rep_cyc = 1;
    mcfk_team_replicas(&rep_ent, &rep_cyc);
    const size_t w_ent = (size_t)2 * rep_ent * pricers * mcf::kMailWords, w_pr = (size_t)2 * pricers * mcf::kMailWords, w_late = 2 * mcf::kMailWords;
    const size_t w_cyc = (size_t)2 * rep_cyc * team * mcf::kMailWords;
    const size_t w_seg = (size_t)4 * (n + 1);     // [2 parities][n+1 entries][2 words]
    const size_t seg_off = (w_ent + w_pr + w_late + w_cyc + 7) & ~(size_t)7;
    CUDA_TRY(h, h->d_mail.ensure(seg_off + w_seg + 8));
    h->h_node.resize(n + 1);
    for (int u = 0; u <= n; ++u) { h->h_node[u].pi = h->h_pi[u]; h->h_node[u].in = h->h_in[u]; h->h_node[u].dp = h->h_depth[u]; }
    cudaStream_t st = h->stream;
    int64_t bytes = 0;
    auto up = [&](void* d, const void* s, size_t b) { bytes += (int64_t)b; return cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, st); };
    CUDA_TRY(h, up(h->d_src.p, h->h_src.data(), (size_t)S * 4)); CUDA_TRY(h, up(h->d_tgt.p, h->h_tgt.data(), (size_t)S * 4));
    CUDA_TRY(h, up(h->d_cost.p, h->h_cost.data(), (size_t)S * 4)); CUDA_TRY(h, up(h->d_state.p, h->h_state.data(), (size_t)A * 4));
    CUDA_TRY(h, up(h->d_flow.p, h->h_flow.data(), (size_t)A * 8)); CUDA_TRY(h, up(h->d_upper.p, h->h_upper.data(), (size_t)A * 8));
    CUDA_TRY(h, up(h->d_sz.p, h->h_sz.data(), (size_t)(n + 1) * 4)); CUDA_TRY(h, up(h->d_pd.p, h->h_pd.data(), (size_t)(n + 1) * 4));
    CUDA_TRY(h, up(h->d_node.p, h->h_node.data(), (size_t)(n + 1) * sizeof(mcf::NodeRec)));
    CUDA_TRY(h, up(d_in_g, h->h_in.data(), (size_t)(n + 1) * 4));
    CUDA_TRY(h, cudaMemsetAsync(h->d_ctl.p, 0, sizeof(mcf::Ctl), st));
    CUDA_TRY(h, cudaMemsetAsync(h->d_done.p, 0, (size_t)(team + pricers) * 32 * sizeof(unsigned), st));
    CUDA_TRY(h, cudaMemsetAsync(h->d_mail.p, 0, (seg_off + w_seg + 8) * sizeof(int4), st));
    h->metrics.h2d_bytes = bytes;
    std::memset(P, 0, sizeof(*P));
    P->n = n; P->m = m; P->S = S; P->A = A;
    P->src = h->d_src.p; P->tgt = h->d_tgt.p; P->cost = h->d_cost.p; P->state = h->d_state.p; P->flow = h->d_flow.p; P->upper = h->d_upper.p;
    P->node = h->d_node.p; P->in_g = d_in_g; P->sz0 = h->d_sz.p; P->pd0 = h->d_pd.p; P->pi_out = h->d_piout.p;
    P->ent0 = h->d_mail.p; P->prc = P->ent0 + w_ent; P->late = P->prc + w_pr; P->cyc = P->late + w_late;
    P->stemseg = h->d_mail.p + seg_off;
    P->done = h->d_done.p; P->ctl = h->d_ctl.p; P->team = team; P->pricers = pricers; P->slice = slice; P->wide = wide;
    P->spill = spill;
    if (spill) {                                        // flows / capacities of the tree arcs, one entry per node, each touched by its owner only
        const size_t ents = (size_t)(team - pricers) * slice + 8;
        CUDA_TRY(h, h->d_flg.ensure(ents)); CUDA_TRY(h, h->d_upg.ensure(ents));
        P->fl_g = h->d_flg.p; P->up_g = h->d_upg.p;
    }
    return MCF_OK;
}


// after a solve: an Optimal result whose device arrays (state, flow) a later warm start may pick up
void note_warm(mcf_handle* h, int st)
{
    h->warm_valid = false;
    if (st != MCF_OPTIMAL || !h->opt.warm_start) return;
    h->warm_upper.swap(h->cur_upper); h->warm_supply.swap(h->cur_supply); h->warm_lower = h->orig_lower; h->warm_supply_type = h->opt.supply_type;
    h->warm_valid = true;
}

int solve_team(mcf_handle* h, int team, int pricers, int slice, int wide, int spill, int block, int dyn_min, const mcf_optimization_config& cfg, bool has_lower,
               clk::time_point t_total, int32_t* status_out, bool* needs_wide)
{
    const int n = h->n, m = h->m, S = m + n;
    auto done = [&](int st) { h->status = st; h->solved_once = true; if (status_out) *status_out = st; h->metrics.total_solve_time_us = us_since(t_total); return MCF_OK; };
    h->metrics.grid_ctas = team;
    const auto t_h2d = clk::now();
    mcf::TeamParams P;
    int rc = upload_team(h, team, pricers, slice, wide, spill, &P);
    if (rc != MCF_OK) return rc;
    if (has_lower) {
        CUDA_TRY(h, h->d_lower.ensure(m));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_lower.p, h->orig_lower.data(), (size_t)m * 8, cudaMemcpyHostToDevice, h->stream));
        h->metrics.h2d_bytes += (int64_t)m * 8;
        P.orig_lower = h->d_lower.p;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->metrics.h2d_time_us = us_since(t_h2d);

    P.block_size = block > 0 ? block : 1; P.dyn_min_block = dyn_min; P.max_block_size = cfg.max_block_size;
    P.adaptive = (cfg.flags & MCF_FLAG_ADAPTIVE_BLOCK_SIZE) ? 1 : 0; P.consecutive = cfg.consecutive_hits_before_adapt;
    P.low_thr = cfg.low_hit_rate_threshold; P.high_thr = cfg.high_hit_rate_threshold;
    P.shrink = cfg.block_size_shrink_factor; P.grow = cfg.block_size_growth_factor;
    P.max_iterations = std::max<int64_t>(1000000LL, (int64_t)n * m);                                  // NS.cs:280
    P.stop_after = h->opt.stop_after_pivots;
    const double tmo = h->opt.barrier_timeout_s > 0 ? h->opt.barrier_timeout_s : 10.0;
    P.timeout_cycles = (unsigned long long)(tmo * 1.9e9);

    {   // keep the node mirror L2-resident (persisting lines), everything else on this stream streams through
        int max_win = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, h->opt.device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->opt.device);
        const size_t mirror_bytes = (size_t)(n + 1) * (sizeof(mcf::NodeRec) + 4);
        if (max_win > 0 && max_persist > 0) {
            const size_t want = std::min<size_t>(mirror_bytes + (1u << 20), (size_t)max_persist);
            ensure_persisting_l2(h->opt.device, want);
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = h->d_node.p;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(mirror_bytes, (size_t)max_win);
            av.accessPolicyWindow.hitRatio = mirror_bytes <= want ? 1.0f : (float)((double)want / (double)mirror_bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
        }
    }
    EventPair evp;
    CUDA_TRY(h, evp.create());
    const cudaEvent_t ev0 = evp.a, ev1 = evp.b;
    CUDA_TRY(h, cudaEventRecord(ev0, h->stream));
    const int lrc = mcfk_launch_team(&P, h->stream);
    if (lrc != 0) { return fail(h, MCF_ERR_CUDA, "cooperative launch of the team kernel failed: %s", cudaGetErrorString((cudaError_t)lrc)); }
    CUDA_TRY(h, cudaEventRecord(ev1, h->stream));
    cudaError_t se = cudaStreamSynchronize(h->stream);
    if (se != cudaSuccess) { return fail(h, MCF_ERR_CUDA, "team kernel failed: %s", cudaGetErrorString(se)); }
    float kms = 0; cudaEventElapsedTime(&kms, ev0, ev1);
    h->metrics.kernel_time_us = kms * 1000.0;

    const auto t_d2h = clk::now();
    mcf::Ctl ctl;
    CUDA_TRY(h, cudaMemcpyAsync(&ctl, h->d_ctl.p, sizeof(ctl), cudaMemcpyDeviceToHost, h->stream));
    h->flow.resize(m); h->pi.resize(n);
    CUDA_TRY(h, cudaMemcpyAsync(h->flow.data(), h->d_flow.p, (size_t)m * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->pi.data(), h->d_piout.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->metrics.d2h_time_us = us_since(t_d2h);
    h->metrics.d2h_bytes = (int64_t)m * 8 + (int64_t)n * 8 + (int64_t)sizeof(ctl);

    if (ctl.needs_wide && !wide) { *needs_wide = true; return MCF_OK; }     // a flow left the int32 range: the caller re-runs wide
    mcf_metrics& M = h->metrics;
    M.iterations = ctl.iterations; M.total_arcs_checked = ctl.arcs_checked; M.final_block_size = ctl.final_block_size;
    M.average_arcs_checked_per_pivot = ctl.iterations > 0 ? (double)ctl.arcs_checked / ctl.iterations : 0;
    M.iteration_ratio = M.baseline_iterations > 0 ? (double)ctl.iterations / M.baseline_iterations : 1.0;
    M.pricer_ctas = pricers; M.wide_flows = wide | (spill ? 2 : 0);
    // phase accumulators are SM clock ticks (reading %globaltimer costs microseconds); scale by the kernel's own ns / tick
    const double ns_per_clk = ctl.clk_total > 0 ? (double)ctl.ns_total / (double)ctl.clk_total : 0.0;
    M.pivot_search_time_us = ctl.ns_price * ns_per_clk / 1000.0; M.cycle_time_us = ctl.ns_cycle * ns_per_clk / 1000.0;
    M.tree_update_time_us = (ctl.ns_update + ctl.ns_wait_done + ctl.ns_stem) * ns_per_clk / 1000.0;
    M.hop_wait_done_us = ctl.ns_wait_done * ns_per_clk / 1000.0; M.stem_exchange_us = ctl.ns_stem * ns_per_clk / 1000.0; M.stem_exchanges = ctl.stem_exchanges;
    M.ns_per_clock = ns_per_clk;
    for (int i = 0; i < 16; ++i) M.phase_us[i] = ctl.clk[i] * ns_per_clk / 1000.0;
    M.degenerate_pivots = ctl.degenerate; M.cycle_nodes = ctl.cycle_nodes; M.moved_nodes = ctl.moved_nodes;
    M.max_cycle = ctl.max_cycle; M.max_stem = ctl.max_stem; M.pricing_rounds = ctl.pricing_rounds;
    M.arcs_priced = ctl.arcs_checked; M.pricing_bytes = 16 * M.arcs_priced; M.engine = 2;
    h->total_cost = ctl.total_cost; h->d_pi_final = h->d_piout.p; h->supply_type_solved = h->opt.supply_type;

    if (ctl.abort || ctl.status == mcf::ST_ERR_BARRIER_TIMEOUT) { done(MCF_NOT_SOLVED); return fail(h, MCF_ERR_TIMEOUT, "team exchange timed out after %lld pivots", (long long)ctl.iterations); }
    if (ctl.status == mcf::ST_ERR_CYCLE_TOO_LONG || ctl.status == mcf::ST_ERR_STEM_TOO_LONG) {
        done(MCF_NOT_SOLVED);
        return fail(h, MCF_ERR_ENGINE_LIMIT, "pivot %lld: stem exceeds the in-kernel staging buffer (%d entries)", (long long)ctl.iterations, mcf::kTeamStemCap);
    }
    int st;
    h->stopped_early = ctl.status == mcf::ST_STOPPED_EARLY;
    switch (ctl.status) {
        case mcf::ST_OPTIMAL: st = ctl.infeasible ? MCF_INFEASIBLE : MCF_OPTIMAL; break;             // NS.cs:360-393
        case mcf::ST_INFEASIBLE: st = MCF_INFEASIBLE; break;
        case mcf::ST_UNBOUNDED: st = MCF_UNBOUNDED; break;
        default: st = MCF_NOT_SOLVED; break;
    }
    if (st == MCF_OPTIMAL && has_lower) {                                                            // NS.cs:375-388 (supplies)
        for (int i = 0; i < m; ++i) if (h->orig_lower[i] != 0) { h->supply[h->source[i]] += h->orig_lower[i]; h->supply[h->target[i]] -= h->orig_lower[i]; }
    }
    (void)S;
    note_warm(h, st);
    return done(st);
}

}  // namespace

// ================================================================================================= C ABI

extern "C" {

int mcf_api_version(void) { return MCF_API_VERSION; }

int mcf_device_count(void)
{
    int count = 0, ok = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (int d = 0; d < count; ++d) ok += device_ok(d) ? 1 : 0;
    return ok;
}

void mcf_default_options(mcf_options* o)
{
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->supply_type = MCF_GEQ; o->pivot_rule = MCF_BLOCK_SEARCH; o->auto_configuration = 1; o->optimized_pivot = 0; o->simd_width = 4;
    o->device = 0; o->max_ctas = 0; o->lookahead_blocks = 0; o->engine = 0; o->stop_after_pivots = 0; o->barrier_timeout_s = 0;
    default_config(&o->config);
}

int mcf_create(int32_t n, int32_t m, const int32_t* source, const int32_t* target, mcf_handle** out)
{
    if (!out) return MCF_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (n < 0 || m < 0 || (m > 0 && (!source || !target))) return MCF_ERR_INVALID_ARGUMENT;
    if ((int64_t)m + 2 * (int64_t)n >= (1LL << 30)) return MCF_ERR_RANGE;
    for (int e = 0; e < m; ++e)
        if (source[e] < 0 || source[e] >= n || target[e] < 0 || target[e] >= n) return MCF_ERR_INVALID_ARGUMENT;
    if (mcf_device_count() <= 0) return MCF_ERR_NO_DEVICE;          // no CPU fallback
    mcf_handle* h = new (std::nothrow) mcf_handle();
    if (!h) return MCF_ERR_OUT_OF_MEMORY;
    h->n = n; h->m = m;
    h->source.assign(source, source + m); h->target.assign(target, target + m);
    h->lower.assign(m, 0); h->upper.assign(m, kInf); h->cost.assign(m, 0); h->orig_lower.assign(m, 0);   // NS.cs:615-617
    h->supply.assign(n, 0);
    mcf_default_options(&h->opt);
    *out = h;
    return MCF_OK;
}

void mcf_destroy(mcf_handle* h)
{
    if (!h) return;
    if (h->device_bound >= 0) {
        cudaSetDevice(h->device_bound);
        h->d_src.release(); h->d_tgt.release(); h->d_cost.release(); h->d_state.release(); h->d_in.release(); h->d_sz.release();
        h->d_parent.release(); h->d_pd.release(); h->d_flow.release(); h->d_upper.release(); h->d_lower.release(); h->d_pi.release();
        h->d_rc.release(); h->d_part.release(); h->d_list.release(); h->d_scratch.release(); h->d_ctl.release(); h->d_flush.release();
        h->d_node.release(); h->d_mail.release(); h->d_done.release(); h->d_piout.release(); h->d_flg.release(); h->d_upg.release(); h->d_cand.release(); h->d_cand_scratch.release(); h->d_cand_cost.release(); h->d_val.release(); h->d_pi_final = nullptr;
        if (h->stream) cudaStreamDestroy(h->stream);
    }
    delete h;
}

int mcf_set_arcs(mcf_handle* h, const int64_t* lower, const int64_t* upper, const int64_t* cost)
{
    if (!h) return MCF_ERR_INVALID_ARGUMENT;
    const int m = h->m;
    if (lower) { h->lower.assign(lower, lower + m); h->orig_lower.assign(lower, lower + m); }       // NS.cs:160-162
    if (upper) h->upper.assign(upper, upper + m);
    if (cost) h->cost.assign(cost, cost + m);
    return MCF_OK;
}

int mcf_get_dims(mcf_handle* h, int32_t* n_out, int32_t* m_out)
{
    if (!h) return MCF_ERR_INVALID_ARGUMENT;
    if (n_out) *n_out = h->n;
    if (m_out) *m_out = h->m;
    return MCF_OK;
}

int mcf_get_endpoints(mcf_handle* h, int32_t* source_out, int32_t* target_out)
{
    if (!h) return MCF_ERR_INVALID_ARGUMENT;
    if (source_out && h->m > 0) std::memcpy(source_out, h->source.data(), (size_t)h->m * 4);
    if (target_out && h->m > 0) std::memcpy(target_out, h->target.data(), (size_t)h->m * 4);
    return MCF_OK;
}

int mcf_set_supply(mcf_handle* h, const int64_t* supply)
{
    if (!h || (!supply && h->n > 0)) return MCF_ERR_INVALID_ARGUMENT;
    h->supply.assign(supply, supply + h->n);
    return MCF_OK;
}

int mcf_set_options(mcf_handle* h, const mcf_options* opt)
{
    if (!h || !opt) return MCF_ERR_INVALID_ARGUMENT;
    if (opt->supply_type != MCF_GEQ && opt->supply_type != MCF_LEQ) return fail(h, MCF_ERR_INVALID_ARGUMENT, "unknown supply type %d", opt->supply_type);
    // The reference throws NotImplementedException for CandidateList / AlteringList (NS.cs:884); here they run, defined as LEMON's
    // (network_simplex.h:413-635).  The optimized wrapper knows the first three rules only (OptimizedPivotWrapper, NS.cs:851-856).
    if (opt->pivot_rule < MCF_FIRST_ELIGIBLE || opt->pivot_rule > MCF_ALTERING_LIST)
        return fail(h, MCF_ERR_INVALID_ARGUMENT, "pivot rule %d does not exist", opt->pivot_rule);
    if (opt->optimized_pivot && opt->pivot_rule > MCF_BLOCK_SEARCH)
        return fail(h, MCF_ERR_INVALID_ARGUMENT, "pivot rule %d has no optimized variant", opt->pivot_rule);
    if (opt->simd_width < 0 || opt->simd_width > 64) return fail(h, MCF_ERR_INVALID_ARGUMENT, "simd_width %d out of range", opt->simd_width);
    h->opt = *opt;
    return MCF_OK;
}

int mcf_solve(mcf_handle* h, int32_t* status_out)
{
    if (!h) return MCF_ERR_INVALID_ARGUMENT;
    const auto t_total = clk::now();
    h->metrics = mcf_metrics{};
    h->status = MCF_NOT_SOLVED;
    h->stopped_early = false;
    h->d_pi_final = nullptr;
    const int n = h->n, m = h->m, S = m + n;
    auto done = [&](int st) { h->status = st; h->solved_once = true; if (status_out) *status_out = st; h->metrics.total_solve_time_us = us_since(t_total); return MCF_OK; };

    int rc = bind_device(h);
    if (rc != MCF_OK) return rc;

    const auto t_pre = clk::now();
    for (int i = 0; i < m; ++i) if (h->upper[i] < h->lower[i]) return done(MCF_INFEASIBLE);          // CheckBounds, NS.cs:624-634
    bool has_lower = false;
    for (int i = 0; i < m; ++i) {                                                                     // NS.cs:639-653
        if (h->lower[i] != 0) {
            h->supply[h->source[i]] -= h->lower[i]; h->supply[h->target[i]] += h->lower[i];
            h->upper[i] -= h->lower[i]; h->lower[i] = 0;
        }
        has_lower |= h->orig_lower[i] != 0;
    }
    int64_t max_cost = 0;
    for (int i = 0; i < m; ++i) { const int64_t a = h->cost[i] < 0 ? -h->cost[i] : h->cost[i]; if (a > max_cost) max_cost = a; }
    if (max_cost > std::numeric_limits<int32_t>::max()) return fail(h, MCF_ERR_RANGE, "|cost| = %lld does not fit the device's int32 cost array", (long long)max_cost);
    const int64_t art_cost = (max_cost + 1) * n;                                                      // NS.cs:663-668

    mcf_optimization_config cfg = h->opt.config;
    if (h->opt.auto_configuration) {                                                                  // NS.cs:237-250
        const Characteristics ch = analyze(*h);
        cfg = select_config(ch);
        h->metrics.degree_cv = ch.degree_cv;
    }
    h->metrics.config_flags = cfg.flags;

    // CreatePivotRuleFinder, NS.cs:847-886 (+ BlockSearchPivot ctor :1304-1337)
    int kind = h->opt.pivot_rule;
    if (!h->opt.optimized_pivot && h->opt.pivot_rule == MCF_BLOCK_SEARCH && (cfg.flags & MCF_FLAG_REDUCED_COST_CACHING)) kind = mcf::PK_BLOCK_CACHED;
    if (h->opt.optimized_pivot && h->opt.pivot_rule == MCF_BLOCK_SEARCH) kind = mcf::PK_BLOCK_OPT;          // NS.cs:851-856
    if (h->opt.pivot_rule == MCF_CANDIDATE_LIST) kind = mcf::PK_CAND_LIST;
    if (h->opt.pivot_rule == MCF_ALTERING_LIST) kind = mcf::PK_ALT_LIST;
    int block = 0, dyn_min = 0, list_length = 0, minor_limit = 0, head_length = 0, cand_cap = 0;
    if (kind == mcf::PK_CAND_LIST) {                                                                  // network_simplex.h:441-458
        list_length = std::max((int)(0.25 * std::sqrt((double)S)), 10);
        minor_limit = std::max((int)(0.1 * list_length), 3);
        cand_cap = list_length;
        cfg.flags &= ~MCF_FLAG_ADAPTIVE_BLOCK_SIZE;
    } else if (kind == mcf::PK_ALT_LIST) {                                                            // network_simplex.h:563-580
        block = std::max((int)(1.0 * std::sqrt((double)S)), 10);
        head_length = std::max((int)(0.01 * block), 3);
        cand_cap = head_length + block;
        cfg.flags &= ~MCF_FLAG_ADAPTIVE_BLOCK_SIZE;
    } else if (kind == mcf::PK_BLOCK_OPT) {
        block = std::max((int)std::sqrt((double)S), 10);                                              // BlockSearchPivotOptimized.cs:27-28 (MIN_BLOCK_SIZE, NS.cs:100)
        cfg.flags &= ~MCF_FLAG_ADAPTIVE_BLOCK_SIZE;                                                   // the optimized rule never adapts
    } else if (kind == mcf::PK_BLOCK || kind == mcf::PK_BLOCK_CACHED) {
        const int base = (int)std::sqrt((double)S);
        const int r = (int)(base * cfg.min_block_size_ratio);
        dyn_min = cfg.min_block_size > r ? cfg.min_block_size : r;
        if (cfg.flags & MCF_FLAG_SMALL_BLOCKS_FOR_DENSE) {
            const double density = (double)S / n;
            block = density > 10 ? std::min(50, base / 4) : base;
        } else block = base;
        block = std::max(block, dyn_min);
        h->metrics.initial_block_size = block;
    }
    h->metrics.pricing_kind = kind;
    h->metrics.baseline_iterations = (int)(std::sqrt((double)S) * n * 0.5);                           // NS.cs:276

    if (n == 0) { h->flow.assign(m, 0); h->pi.clear(); h->total_cost = 0; return done(MCF_OPTIMAL); }

    bool warm = false;
    if (h->opt.warm_start) {
        h->cur_upper = h->upper; h->cur_supply = h->supply;                                           // standard form: what the basis belongs to
        if (h->warm_valid && h->warm_supply_type == h->opt.supply_type && h->warm_upper == h->cur_upper && h->warm_supply == h->cur_supply && h->warm_lower == h->orig_lower)
            warm = build_warm_basis(h, art_cost);
    }
    h->warm_valid = false;
    if (!warm) build_initial_basis(h, art_cost);
    h->metrics.warm_started = warm ? 1 : 0;
    h->metrics.host_prepass_time_us = us_since(t_pre);

    // engine: the team engine (node slices resident in shared memory) runs plain Block Search; everything else, and
    // instances whose slices do not fit, run on the flat engine (mcf_kernels.cu).  opt.engine: 0 auto, 1 flat, 2 team.
    if (kind == mcf::PK_BLOCK && h->opt.engine != 1) {
        // narrow mode keeps tree-arc flows / capacities as int32 in shared memory: every capacity is "infinite" (== INF) or < 2^31 - 1
        int wide = 0;
        for (int i = 0; i < m && !wide; ++i) if (h->upper[i] != kInf && h->upper[i] >= (int64_t)std::numeric_limits<int32_t>::max()) wide = 1;
        // narrow mode also needs every tree-arc flow below 2^31 - 1: a flow never exceeds the sum of the positive supplies (after
        // the lower-bound shift) unless it runs around a cycle of uncapacitated arcs, which the overflow flag of the kernel catches
        if (!wide) {
            int64_t pos = 0;
            for (int u = 0; u < n; ++u) { const int64_t a = h->supply[u] < 0 ? -h->supply[u] : h->supply[u]; pos += a; if (pos >= (int64_t)std::numeric_limits<int32_t>::max()) { wide = 1; break; } }
        }
        // per flow width: the slices with the tree-arc flows resident, else with the flows in global memory (16 B per node resident:
        // n up to ~1.5 M, and 2^20 nodes with 64-bit flows); opt.engine == 3 asks for the second form (tests)
        // The shared memory of the slices comes out of the SM's L1, and the pricing CTAs (same launch, same carve-out) gather the
        // next block's node records through it: with the flows resident a 2^20-node instance leaves 34 KB of L1, with the flows in
        // global memory 97 KB.  Over FULL solves (profiles/r02_team_tuning.txt): NETGEN-8 2^20 7.94 -> 7.68 us per pivot, but the
        // 1024^2 time-expanded grid 11.8 -> 13.0 - its cycles are thousands of nodes long and every cycle node's flow then comes
        // from global memory.  So the spilled form is chosen when the resident slices would leave less than ~64 KB of L1 AND the
        // graph is dense enough (m >= 6 n) for shallow basis trees and short cycles to be the rule - and whenever the resident
        // form does not fit at all.
        constexpr size_t kResidentMax = 160u << 10;
        const bool short_cycles = (int64_t)m >= 6 * (int64_t)n;
        for (; wide < 2; ++wide) {
            int slice = 0, pricers = 0, spill = h->opt.engine == 3 ? 1 : 0;
            int team = choose_team(h, wide, spill, &slice, &pricers);
            const bool force_resident = std::getenv("MCF_FORCE_RESIDENT") != nullptr;                     // tuning aid
            if (!spill && (team <= 0 || (mcfk_team_smem_bytes(slice, wide, 0) > kResidentMax && short_cycles && !force_resident))) {
                int s2 = 0, p2 = 0;
                const int t2 = choose_team(h, wide, 1, &s2, &p2);
                if (t2 > 0) { team = t2; slice = s2; pricers = p2; spill = 1; }
            }
            if (team <= 0) break;
            bool needs_wide = false;
            rc = solve_team(h, team, pricers, slice, wide, spill, block, dyn_min, cfg, has_lower, t_total, status_out, &needs_wide);
            if (rc != MCF_OK || !needs_wide) return rc;
        }
        if (h->opt.engine >= 2) return fail(h, MCF_ERR_ENGINE_LIMIT, "team engine requested but the instance does not fit (n = %d)", n);
    } else if (h->opt.engine >= 2) return fail(h, MCF_ERR_ENGINE_LIMIT, "team engine requested but not applicable (pricing kind %d)", kind);

    int sms = 0;
    const int grid = choose_grid(h, &sms);
    if (grid <= 0) return grid;
    h->metrics.grid_ctas = grid;

    const auto t_h2d = clk::now();
    rc = upload_basis(h, kind == mcf::PK_BLOCK_CACHED, grid);
    if (rc != MCF_OK) return rc;
    mcf::Params P; fill_params(h, &P);
    if (has_lower) {
        CUDA_TRY(h, h->d_lower.ensure(m));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_lower.p, h->orig_lower.data(), (size_t)m * 8, cudaMemcpyHostToDevice, h->stream));
        h->metrics.h2d_bytes += (int64_t)m * 8;
        P.orig_lower = h->d_lower.p;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->metrics.h2d_time_us = us_since(t_h2d);

    if (cand_cap > 0) {
        CUDA_TRY(h, h->d_cand.ensure(2 * (size_t)cand_cap)); CUDA_TRY(h, h->d_cand_scratch.ensure(2 * (size_t)cand_cap));
        CUDA_TRY(h, h->d_cand_cost.ensure(2 * (size_t)cand_cap));
        P.cand = h->d_cand.p; P.cand_scratch = h->d_cand_scratch.p; P.cand_cost = h->d_cand_cost.p; P.cand_cap = cand_cap;
        P.list_length = list_length; P.minor_limit = minor_limit; P.head_length = head_length;
    }
    P.rc_cache = kind == mcf::PK_BLOCK_CACHED ? h->d_rc.p : nullptr;
    P.kind = kind; P.block_size = block > 0 ? block : 1; P.dyn_min_block = dyn_min; P.max_block_size = cfg.max_block_size;
    P.adaptive = (cfg.flags & MCF_FLAG_ADAPTIVE_BLOCK_SIZE) ? 1 : 0; P.consecutive = cfg.consecutive_hits_before_adapt;
    P.low_thr = cfg.low_hit_rate_threshold; P.high_thr = cfg.high_hit_rate_threshold;
    P.shrink = cfg.block_size_shrink_factor; P.grow = cfg.block_size_growth_factor;
    P.lookahead0 = h->opt.lookahead_blocks > 0 ? h->opt.lookahead_blocks : 2;
    P.simd_width = h->opt.simd_width;
    P.max_iterations = std::max<int64_t>(1000000LL, (int64_t)n * m);                                  // NS.cs:280
    P.stop_after = h->opt.stop_after_pivots;
    const double tmo = h->opt.barrier_timeout_s > 0 ? h->opt.barrier_timeout_s : 10.0;
    P.barrier_timeout_cycles = (unsigned long long)(tmo * 1.9e9);

    EventPair evp;
    CUDA_TRY(h, evp.create());
    const cudaEvent_t ev0 = evp.a, ev1 = evp.b;
    CUDA_TRY(h, cudaEventRecord(ev0, h->stream));
    const int lrc = mcfk_launch_pivot(&P, grid, h->stream);
    if (lrc != 0) { return fail(h, MCF_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString((cudaError_t)lrc)); }
    CUDA_TRY(h, cudaEventRecord(ev1, h->stream));
    cudaError_t se = cudaStreamSynchronize(h->stream);
    if (se != cudaSuccess) { return fail(h, MCF_ERR_CUDA, "pivot kernel failed: %s", cudaGetErrorString(se)); }
    float kms = 0; cudaEventElapsedTime(&kms, ev0, ev1);
    h->metrics.kernel_time_us = kms * 1000.0;

    const auto t_d2h = clk::now();
    mcf::Ctl ctl;
    CUDA_TRY(h, cudaMemcpyAsync(&ctl, h->d_ctl.p, sizeof(ctl), cudaMemcpyDeviceToHost, h->stream));
    h->flow.resize(m); h->pi.resize(n);
    CUDA_TRY(h, cudaMemcpyAsync(h->flow.data(), h->d_flow.p, (size_t)m * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->pi.data(), h->d_pi.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->metrics.d2h_time_us = us_since(t_d2h);
    h->metrics.d2h_bytes = (int64_t)m * 8 + (int64_t)n * 8 + (int64_t)sizeof(ctl);

    mcf_metrics& M = h->metrics;
    M.iterations = ctl.iterations; M.total_arcs_checked = ctl.arcs_checked;
    M.final_block_size = (kind >= 2 && kind != mcf::PK_CAND_LIST && kind != mcf::PK_ALT_LIST) ? ctl.final_block_size : 0;
    M.average_arcs_checked_per_pivot = ctl.iterations > 0 ? (double)ctl.arcs_checked / ctl.iterations : 0;
    M.iteration_ratio = M.baseline_iterations > 0 ? (double)ctl.iterations / M.baseline_iterations : 1.0;
    M.pivot_search_time_us = ctl.ns_price / 1000.0; M.cycle_time_us = ctl.ns_cycle / 1000.0; M.tree_update_time_us = ctl.ns_update / 1000.0;
    M.degenerate_pivots = ctl.degenerate; M.cycle_nodes = ctl.cycle_nodes; M.moved_nodes = ctl.moved_nodes;
    M.max_cycle = ctl.max_cycle; M.max_stem = ctl.max_stem; M.pricing_rounds = ctl.pricing_rounds;
    if (kind == mcf::PK_BEST) M.arcs_priced = ctl.pricing_rounds * (int64_t)S;
    else if (kind == mcf::PK_FIRST) M.arcs_priced = 0;       // not tracked for First Eligible
    else if (kind == mcf::PK_BLOCK_OPT) { M.arcs_priced = ctl.arcs_priced_opt; M.final_block_size = 0; M.average_arcs_checked_per_pivot = 0; }
    else M.arcs_priced = ctl.arcs_checked;
    M.pricing_bytes = 16 * M.arcs_priced; M.engine = 1;
    h->total_cost = ctl.total_cost; h->d_pi_final = h->d_pi.p; h->supply_type_solved = h->opt.supply_type;

    if (ctl.abort || ctl.status == mcf::ST_ERR_BARRIER_TIMEOUT) { done(MCF_NOT_SOLVED); return fail(h, MCF_ERR_TIMEOUT, "grid barrier timed out after %lld pivots", (long long)ctl.iterations); }
    if (ctl.status == mcf::ST_ERR_CYCLE_TOO_LONG || ctl.status == mcf::ST_ERR_STEM_TOO_LONG) {
        done(MCF_NOT_SOLVED);
        return fail(h, MCF_ERR_ENGINE_LIMIT, "pivot %lld: cycle/stem exceeds the in-kernel staging buffer (%d / %d entries)", (long long)ctl.iterations, mcf::kListSmem, mcf::kStemCap);
    }
    int st;
    h->stopped_early = ctl.status == mcf::ST_STOPPED_EARLY;
    switch (ctl.status) {
        case mcf::ST_OPTIMAL: st = ctl.infeasible ? MCF_INFEASIBLE : MCF_OPTIMAL; break;             // NS.cs:360-393
        case mcf::ST_INFEASIBLE: st = MCF_INFEASIBLE; break;
        case mcf::ST_UNBOUNDED: st = MCF_UNBOUNDED; break;
        default: st = MCF_NOT_SOLVED; break;
    }
    if (st == MCF_OPTIMAL && has_lower) {                                                            // NS.cs:375-388 (supplies)
        for (int i = 0; i < m; ++i) if (h->orig_lower[i] != 0) { h->supply[h->source[i]] += h->orig_lower[i]; h->supply[h->target[i]] -= h->orig_lower[i]; }
    }
    note_warm(h, st);
    return done(st);
}

int mcf_get_status(mcf_handle* h, int32_t* out) { if (!h || !out) return MCF_ERR_INVALID_ARGUMENT; *out = h->status; return MCF_OK; }

int mcf_get_flows(mcf_handle* h, int64_t* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (h->status != MCF_OPTIMAL) return fail(h, MCF_ERR_NOT_OPTIMAL, "Solution not optimal");
    std::memcpy(out, h->flow.data(), (size_t)h->m * 8);
    return MCF_OK;
}
int mcf_get_potentials(mcf_handle* h, int64_t* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (h->status != MCF_OPTIMAL) return fail(h, MCF_ERR_NOT_OPTIMAL, "Solution not optimal");
    std::memcpy(out, h->pi.data(), (size_t)h->n * 8);
    return MCF_OK;
}
int mcf_get_flow(mcf_handle* h, int32_t arc, int64_t* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (h->status != MCF_OPTIMAL) return fail(h, MCF_ERR_NOT_OPTIMAL, "Solution not optimal");       // NS.cs:418-421
    if (arc < 0 || arc >= h->m) return fail(h, MCF_ERR_INVALID_ARGUMENT, "Invalid arc");              // NS.cs:423-426
    *out = h->flow[arc];
    return MCF_OK;
}
int mcf_get_potential(mcf_handle* h, int32_t node, int64_t* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (h->status != MCF_OPTIMAL) return fail(h, MCF_ERR_NOT_OPTIMAL, "Solution not optimal");
    if (node < 0 || node >= h->n) return fail(h, MCF_ERR_INVALID_ARGUMENT, "Invalid node");
    *out = h->pi[node];
    return MCF_OK;
}
int mcf_get_total_cost(mcf_handle* h, int64_t* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (h->status != MCF_OPTIMAL) return fail(h, MCF_ERR_NOT_OPTIMAL, "Solution not optimal");
    *out = h->total_cost;
    return MCF_OK;
}
int mcf_get_node_supply(mcf_handle* h, int32_t node, int64_t* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (node < 0 || node >= h->n) return fail(h, MCF_ERR_INVALID_ARGUMENT, "Invalid node");
    *out = h->supply[node]; return MCF_OK;
}
int mcf_get_arc_cost(mcf_handle* h, int32_t arc, int64_t* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (arc < 0 || arc >= h->m) return fail(h, MCF_ERR_INVALID_ARGUMENT, "Invalid arc");
    *out = h->cost[arc]; return MCF_OK;
}
int mcf_get_arc_lower_bound(mcf_handle* h, int32_t arc, int64_t* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (arc < 0 || arc >= h->m) return fail(h, MCF_ERR_INVALID_ARGUMENT, "Invalid arc");
    *out = h->orig_lower[arc]; return MCF_OK;                                                         // NS.cs:513
}
int mcf_get_arc_upper_bound(mcf_handle* h, int32_t arc, int64_t* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (arc < 0 || arc >= h->m) return fail(h, MCF_ERR_INVALID_ARGUMENT, "Invalid arc");
    *out = h->upper[arc]; return MCF_OK;                                                              // NS.cs:526 (shifted after Solve)
}

int mcf_get_metrics(mcf_handle* h, mcf_metrics* out)
{
    if (!h || !out) return MCF_ERR_INVALID_ARGUMENT;
    if (!h->solved_once) return fail(h, MCF_ERR_NOT_SOLVED, "Solve() has not been called");
    *out = h->metrics; return MCF_OK;
}

int mcf_get_state_after_stop(mcf_handle* h, int64_t* flow_out, int64_t* potential_out)
{
    if (!h) return MCF_ERR_INVALID_ARGUMENT;
    if (!h->stopped_early) return fail(h, MCF_ERR_NOT_SOLVED, "the last solve did not end at stop_after_pivots");
    if (flow_out) std::memcpy(flow_out, h->flow.data(), (size_t)h->m * 8);
    if (potential_out) std::memcpy(potential_out, h->pi.data(), (size_t)h->n * 8);
    return MCF_OK;
}

int mcf_get_device_results(mcf_handle* h, const int64_t** flow_dev_out, const int64_t** potential_dev_out, int32_t* device_out)
{
    if (!h) return MCF_ERR_INVALID_ARGUMENT;
    if (h->status != MCF_OPTIMAL || !h->d_pi_final || !h->d_flow.p) return fail(h, MCF_ERR_NOT_OPTIMAL, "Solution not optimal");
    if (flow_dev_out) *flow_dev_out = reinterpret_cast<const int64_t*>(h->d_flow.p);
    if (potential_dev_out) *potential_dev_out = reinterpret_cast<const int64_t*>(h->d_pi_final);
    if (device_out) *device_out = h->device_bound;
    return MCF_OK;
}

int mcf_solve_batch_concurrent(mcf_handle** hs, int32_t count, const int32_t* devices, int32_t n_devices, int32_t per_device, int32_t* statuses_out)
{
    if (!hs || count < 0 || !devices || n_devices <= 0 || per_device <= 0) return MCF_ERR_INVALID_ARGUMENT;
    // `per_device` solves share one GPU: each is its own cooperative launch on its own stream over a 1 / per_device share of
    // the SMs (a 2^18-node instance needs 37 CTAs, four run side by side on 148 SMs).  The solves are independent - no kernel
    // ever waits for another one - so they may also simply run one after the other if the SMs are not free.
    const int lanes = n_devices * per_device;
    if (per_device > 1) {                           // raise the persisting-L2 set-aside once, before any kernel runs (see ensure_persisting_l2)
        int max_n = 0;
        for (int i = 0; i < count; ++i) if (hs[i] && hs[i]->n > max_n) max_n = hs[i]->n;
        for (int d = 0; d < n_devices; ++d) {
            int max_persist = 0;
            if (cudaSetDevice(devices[d]) != cudaSuccess || cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, devices[d]) != cudaSuccess) { cudaGetLastError(); continue; }
            const size_t one = (size_t)(max_n + 1) * (sizeof(mcf::NodeRec) + 4) + (1u << 20);
            ensure_persisting_l2(devices[d], std::min<size_t>(one * (size_t)per_device, (size_t)max_persist));
        }
    }
    std::vector<int> rcs(lanes, MCF_OK);
    std::vector<std::thread> workers;
    for (int w = 0; w < lanes; ++w) {
        workers.emplace_back([&, w]() {
            const int d = w % n_devices;
            int sms = 0;
            if (per_device > 1) { cudaDeviceProp prop; if (mcfk_device_props(devices[d], &prop) == cudaSuccess) sms = prop.multiProcessorCount; }
            for (int i = w; i < count; i += lanes) {
                if (!hs[i]) { rcs[w] = MCF_ERR_INVALID_ARGUMENT; continue; }
                const mcf_options saved = hs[i]->opt;                       // device and CTA share are per-call choices, not handle state
                hs[i]->opt.device = devices[d];
                if (per_device > 1 && sms > 0 && hs[i]->opt.max_ctas <= 0) hs[i]->opt.max_ctas = sms / per_device;
                int32_t st = MCF_NOT_SOLVED;
                const int rc = mcf_solve(hs[i], &st);
                hs[i]->opt = saved;
                if (statuses_out) statuses_out[i] = st;
                if (rc != MCF_OK && rcs[w] == MCF_OK) rcs[w] = rc;
            }
        });
    }
    for (auto& w : workers) w.join();
    for (int w = 0; w < lanes; ++w) if (rcs[w] != MCF_OK) return rcs[w];
    return MCF_OK;
}

int mcf_solve_batch(mcf_handle** hs, int32_t count, const int32_t* devices, int32_t n_devices, int32_t* statuses_out)
{
    return mcf_solve_batch_concurrent(hs, count, devices, n_devices, 1, statuses_out);
}

int mcf_pricing_probe(mcf_handle* h, int32_t reps, int32_t flush_l2, float* ms_out, int32_t* entering_arc_out, int64_t* arcs_out)
{
    if (!h || reps <= 0 || !ms_out) return MCF_ERR_INVALID_ARGUMENT;
    int rc = bind_device(h);
    if (rc != MCF_OK) return rc;
    h->d_pi_final = nullptr;                                    // the probe re-uploads the initial basis over the solve's arrays
    h->warm_valid = false;
    const int n = h->n, m = h->m, S = m + n;
    // same pre-pass as mcf_solve, on copies: the probe must not disturb the handle's problem data
    std::vector<int64_t> sv_supply = h->supply, sv_upper = h->upper;
    for (int i = 0; i < m; ++i) if (h->lower[i] != 0) { h->supply[h->source[i]] -= h->lower[i]; h->supply[h->target[i]] += h->lower[i]; h->upper[i] -= h->lower[i]; }
    int64_t max_cost = 0;
    for (int i = 0; i < m; ++i) { const int64_t a = h->cost[i] < 0 ? -h->cost[i] : h->cost[i]; if (a > max_cost) max_cost = a; }
    if (max_cost > std::numeric_limits<int32_t>::max()) { h->supply = sv_supply; h->upper = sv_upper; return fail(h, MCF_ERR_RANGE, "cost range"); }
    build_initial_basis(h, (max_cost + 1) * n);
    h->supply = sv_supply; h->upper = sv_upper;
    int sms = 0;
    const int maxg = mcfk_max_grid(h->opt.device, &sms);
    if (maxg <= 0) return fail(h, MCF_ERR_CUDA, "occupancy query failed");
    const int grid = sms;                           // one CTA per SM
    rc = upload_basis(h, false, grid);
    if (rc != MCF_OK) return rc;
    mcf::Params P; fill_params(h, &P);
    P.kind = mcf::PK_BEST;
    // L2 flush between launches: overwrite a buffer larger than L2 (126 MB), then (flush_l2 == 1) read a second one, so that
    // the kernel starts on a cold L2 without having to write the first buffer's dirty lines back inside the timed region
    // (tools/micro/stream.cu, profiles/r01_micro_stream.txt: +4 us on a 151 MB read-only stream otherwise)
    const size_t flush_bytes = 256u << 20;
    if (flush_l2) { CUDA_TRY(h, h->d_flush.ensure(2 * flush_bytes)); CUDA_TRY(h, cudaMemsetAsync(h->d_flush.p + flush_bytes, 0x5a, flush_bytes, h->stream)); }
    EventPair evp;
    CUDA_TRY(h, evp.create());
    const cudaEvent_t ev0 = evp.a, ev1 = evp.b;
    for (int r = 0; r < reps; ++r) {
        if (flush_l2) {
            CUDA_TRY(h, cudaMemsetAsync(h->d_flush.p, r & 0xff, flush_bytes, h->stream));
            if (flush_l2 == 1) {
                const int frc = mcfk_launch_l2_read(h->d_flush.p + flush_bytes, flush_bytes, reinterpret_cast<long long*>(h->d_part.p), sms, h->stream);
                if (frc != 0) return fail(h, MCF_ERR_CUDA, "flush launch failed: %s", cudaGetErrorString((cudaError_t)frc));
            }
        }
        CUDA_TRY(h, cudaEventRecord(ev0, h->stream));
        const int lrc = mcfk_launch_price_sweep(&P, h->d_part.p, grid, h->stream);
        if (lrc != 0) return fail(h, MCF_ERR_CUDA, "sweep launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
        CUDA_TRY(h, cudaEventRecord(ev1, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        CUDA_TRY(h, cudaEventElapsedTime(&ms_out[r], ev0, ev1));
    }
    std::vector<mcf::PriceRec> recs(grid);
    CUDA_TRY(h, cudaMemcpy(recs.data(), h->d_part.p, sizeof(mcf::PriceRec) * grid, cudaMemcpyDeviceToHost));
    long long bc = 0; int ba = -1;
    for (int g = 0; g < grid; ++g) if (recs[g].c < bc || (recs[g].c == bc && recs[g].c < 0 && recs[g].arc < ba)) { bc = recs[g].c; ba = recs[g].arc; }
    if (entering_arc_out) *entering_arc_out = ba;
    if (arcs_out) *arcs_out = S;
    return MCF_OK;
}

int mcf_validate(mcf_handle* h, int32_t* failed_checks_out, int64_t* primal_out, int64_t* dual_out)
{
    if (!h || !failed_checks_out) return MCF_ERR_INVALID_ARGUMENT;
    if (h->status != MCF_OPTIMAL) return fail(h, MCF_ERR_NOT_OPTIMAL, "Solution not optimal");        // SolutionValidator.cs:24-33
    const int n = h->n, m = h->m;
    if (n == 0) { *failed_checks_out = 0; if (primal_out) *primal_out = 0; if (dual_out) *dual_out = 0; return MCF_OK; }
    if (!h->d_pi_final) return fail(h, MCF_ERR_NOT_SOLVED, "the solve's arrays are no longer resident on the device (a pricing probe ran since)");
    CUDA_TRY(h, cudaSetDevice(h->device_bound));
    // device-resident from the solve: src, tgt, cost (int32), flow (final, lower bounds restored), pi.  Uploaded here: the
    // caller's bounds and supplies (h->upper is shifted by the lower bound after Solve(), NS.cs:647-651; undo that).
    const size_t words = (size_t)3 * m + (size_t)3 * n + 8;
    CUDA_TRY(h, h->d_val.ensure(words));
    long long* d_lower = h->d_val.p; long long* d_upper = d_lower + m; long long* d_supply = d_upper + m;
    long long* d_net = d_supply + n; long long* d_adj = d_net + n; long long* d_out = d_adj + n;
    std::vector<int64_t> up(m);
    for (int i = 0; i < m; ++i) up[i] = h->upper[i] >= kInf ? h->upper[i] : h->upper[i] + h->orig_lower[i];
    CUDA_TRY(h, cudaMemcpyAsync(d_lower, h->orig_lower.data(), (size_t)m * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_upper, up.data(), (size_t)m * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_supply, h->supply.data(), (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(d_net, 0, ((size_t)2 * n + 8) * 8, h->stream));
    mcf::ValidateParams V{};
    V.n = n; V.m = m; V.supply_type = h->supply_type_solved;
    V.src = h->d_src.p; V.tgt = h->d_tgt.p; V.cost = h->d_cost.p; V.flow = h->d_flow.p; V.lower = d_lower; V.upper = d_upper;
    V.supply = d_supply; V.pi = h->d_pi_final; V.net = d_net; V.adj = d_adj; V.out = d_out;
    cudaDeviceProp prop;
    CUDA_TRY(h, mcfk_device_props(h->device_bound, &prop));
    const int lrc = mcfk_launch_validate(&V, prop.multiProcessorCount, h->stream);
    if (lrc != 0) return fail(h, MCF_ERR_CUDA, "validator launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
    long long out[3] = {0, 0, 0};
    CUDA_TRY(h, cudaMemcpyAsync(out, d_out, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    int bad = (int)out[0];
    if (out[1] != h->total_cost) bad |= 16;                                                          // SolutionValidator.cs:234-262
    if (out[2] != h->total_cost) bad |= 32;                                                          // :268-342
    *failed_checks_out = bad;
    if (primal_out) *primal_out = out[1];
    if (dual_out) *dual_out = out[2];
    return MCF_OK;
}

const char* mcf_last_error(mcf_handle* h) { return h ? h->err.c_str() : "invalid handle"; }

}  // extern "C"
