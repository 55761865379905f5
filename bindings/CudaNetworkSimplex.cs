// CudaNetworkSimplex.cs - the reference-side binding of libmcfgpu.so: a drop-in IMinCostFlowSolver
// (src/MinCostFlow.Core/IMinCostFlowSolver.cs:8-34) with the NetworkSimplex setter surface
// (src/MinCostFlow.Core/Lemon/Algorithms/NetworkSimplex.cs:153-210, :470-587).
// NOT compiled in this repository's image (no dotnet); it is the file a maintainer adds to MinCostFlow.Core.
// Blittable structs mirror include/mcfgpu.h field for field.
using System;
using System.Runtime.InteropServices;
using MinCostFlow.Core.Lemon.Algorithms;
using MinCostFlow.Core.Lemon.Graphs;
using MinCostFlow.Core.Lemon.Types;

namespace MinCostFlow.Core.Cuda
{
    [StructLayout(LayoutKind.Sequential)]
    public struct McfOptimizationConfig
    {
        public int Flags, MaxBlockSize, MinBlockSize, DenseNetworkThreshold, ConsecutiveHitsBeforeAdapt, Reserved0;
        public double CandidateListRatio, BlockSizeGrowthFactor, BlockSizeShrinkFactor, LowHitRateThreshold, HighHitRateThreshold, MinBlockSizeRatio;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct McfOptions
    {
        public int SupplyType, PivotRule, AutoConfiguration, OptimizedPivot, Device, MaxCtas, LookaheadBlocks, Engine, SimdWidth, WarmStart;
        public long StopAfterPivots;
        public double BarrierTimeoutSeconds;
        public McfOptimizationConfig Config;
    }

    internal static unsafe class Native
    {
        private const string Lib = "mcfgpu";   // libmcfgpu.so
        [DllImport(Lib)] internal static extern int mcf_create(int n, int m, int* source, int* target, out IntPtr handle);
        [DllImport(Lib)] internal static extern void mcf_destroy(IntPtr h);
        [DllImport(Lib)] internal static extern void mcf_default_options(out McfOptions o);
        [DllImport(Lib)] internal static extern int mcf_set_arcs(IntPtr h, long* lower, long* upper, long* cost);
        [DllImport(Lib)] internal static extern int mcf_set_supply(IntPtr h, long* supply);
        [DllImport(Lib)] internal static extern int mcf_set_options(IntPtr h, ref McfOptions o);
        [DllImport(Lib)] internal static extern int mcf_solve(IntPtr h, out int status);
        // Problems/Loaders: DimacsReader.cs:25-147, SolutionLoader.cs:69-214 (bulk, native)
        [DllImport(Lib)] internal static extern int mcf_create_from_dimacs([MarshalAs(UnmanagedType.LPUTF8Str)] string path, out IntPtr handle);
        [DllImport(Lib)] internal static extern int mcf_dimacs_open([MarshalAs(UnmanagedType.LPUTF8Str)] string path, out IntPtr dimacs);
        [DllImport(Lib)] internal static extern int mcf_dimacs_dims(IntPtr dimacs, out int n, out int m);
        [DllImport(Lib)] internal static extern int mcf_dimacs_copy(IntPtr dimacs, int* source, int* target, long* lower, long* upper, long* cost, long* supply);
        [DllImport(Lib)] internal static extern void mcf_dimacs_close(IntPtr dimacs);
        [DllImport(Lib)] internal static extern int mcf_write_solution(IntPtr h, [MarshalAs(UnmanagedType.LPUTF8Str)] string path, int format, int withPotentials);
        [DllImport(Lib)] internal static extern IntPtr mcf_io_last_error();
        [DllImport(Lib)] internal static extern int mcf_get_flows(IntPtr h, long* outM);
        [DllImport(Lib)] internal static extern int mcf_get_potentials(IntPtr h, long* outN);
        [DllImport(Lib)] internal static extern int mcf_get_total_cost(IntPtr h, out long cost);
        [DllImport(Lib)] internal static extern int mcf_validate(IntPtr h, out int failedChecks, out long primal, out long dual);
        [DllImport(Lib)] internal static extern IntPtr mcf_last_error(IntPtr h);
    }

    /// <summary>NetworkSimplex on a B200 through libmcfgpu.so.  Same results, bit for bit, as NetworkSimplex.Solve().</summary>
    public sealed unsafe class CudaNetworkSimplex : IMinCostFlowSolver, IDisposable
    {
        private readonly IGraph _graph;
        private readonly int _n, _m;
        private readonly long[] _lower, _upper, _cost, _supply;
        private long[] _flow = Array.Empty<long>(), _pi = Array.Empty<long>();
        private long _totalCost;
        private McfOptions _opt;
        private IntPtr _h;
        public SolverStatus Status { get; private set; } = SolverStatus.NotSolved;

        public CudaNetworkSimplex(IGraph graph)
        {
            _graph = graph ?? throw new ArgumentNullException(nameof(graph));          // NetworkSimplex.cs:121
            _n = graph.NodeCount; _m = graph.ArcCount;
            var src = new int[_m]; var tgt = new int[_m];
            for (int e = 0; e < _m; e++) { src[e] = graph.Source(new Arc(e)).Id; tgt[e] = graph.Target(new Arc(e)).Id; }   // NetworkSimplex.cs:605-613
            _lower = new long[_m]; _upper = new long[_m]; _cost = new long[_m]; _supply = new long[_n];
            Array.Fill(_upper, long.MaxValue / 2);                                         // INF, NetworkSimplex.cs:127, :616
            Native.mcf_default_options(out _opt);
            fixed (int* s = src, t = tgt) Check(Native.mcf_create(_n, _m, s, t, out _h));
        }

        public CudaNetworkSimplex SetArcBounds(Arc arc, long lower, long upper) { CheckArc(arc); _lower[arc.Id] = lower; _upper[arc.Id] = upper; return this; }
        public CudaNetworkSimplex SetArcCost(Arc arc, long cost) { CheckArc(arc); _cost[arc.Id] = cost; return this; }
        public CudaNetworkSimplex SetNodeSupply(Node node, long supply) { if (!_graph.IsValidNode(node)) throw new ArgumentException("Invalid node"); _supply[node.Id] = supply; return this; }
        public CudaNetworkSimplex SetSupplyType(SupplyType type) { _opt.SupplyType = (int)type; return this; }
        public CudaNetworkSimplex SetPivotRule(PivotRule rule) { _opt.PivotRule = (int)rule; return this; }
        /// <summary>README.md:17-18 roadmap item: a Solve() after SetArcCost edits starts from the previous optimal basis (mcf_options.warm_start).</summary>
        public void EnableWarmStart(bool enable = true) { _opt.WarmStart = enable ? 1 : 0; }

        public void EnableOptimizedPivot(bool enable = true)
        {
            _opt.OptimizedPivot = enable ? 1 : 0;
            // BlockSearchPivotOptimized.cs:74, :119 - the managed rule's pivot sequence depends on the vector width of this host
            _opt.SimdWidth = System.Numerics.Vector.IsHardwareAccelerated ? System.Numerics.Vector<long>.Count : 0;
        }
        public void SetAutoConfiguration(bool enable) => _opt.AutoConfiguration = enable ? 1 : 0;
        public void SetOptimizationConfig(OptimizationConfig c)
        {
            _opt.Config = new McfOptimizationConfig {
                Flags = (int)c.Flags, MaxBlockSize = c.MaxBlockSize, MinBlockSize = c.MinBlockSize, DenseNetworkThreshold = c.DenseNetworkThreshold,
                ConsecutiveHitsBeforeAdapt = c.ConsecutiveHitsBeforeAdapt, CandidateListRatio = c.CandidateListRatio,
                BlockSizeGrowthFactor = c.BlockSizeGrowthFactor, BlockSizeShrinkFactor = c.BlockSizeShrinkFactor,
                LowHitRateThreshold = c.LowHitRateThreshold, HighHitRateThreshold = c.HighHitRateThreshold, MinBlockSizeRatio = c.MinBlockSizeRatio };
            _opt.AutoConfiguration = 0;                                                    // NetworkSimplex.cs:560
        }

        public SolverStatus Solve()
        {
            // CandidateList / AlteringList run here (defined as LEMON's, network_simplex.h:413-635); NetworkSimplex.cs:884 throws on them.
            if (_opt.PivotRule > (int)PivotRule.BlockSearch && _opt.OptimizedPivot != 0) throw new NotImplementedException($"Optimized pivot rule {(PivotRule)_opt.PivotRule} not implemented");   // NetworkSimplex.cs:1694
            fixed (long* lo = _lower, up = _upper, co = _cost, su = _supply)               // pinned for the call only, like OptimizedPivotWrapper (NetworkSimplex.cs:1699-1722)
            {
                Check(Native.mcf_set_arcs(_h, lo, up, co));
                Check(Native.mcf_set_supply(_h, su));
            }
            Check(Native.mcf_set_options(_h, ref _opt));
            Check(Native.mcf_solve(_h, out int st));
            Status = (SolverStatus)st;
            if (Status == SolverStatus.Optimal)
            {
                _flow = new long[_m]; _pi = new long[_n];
                fixed (long* f = _flow, p = _pi) { Check(Native.mcf_get_flows(_h, f)); Check(Native.mcf_get_potentials(_h, p)); }
                Check(Native.mcf_get_total_cost(_h, out _totalCost));
            }
            return Status;
        }

        public long GetFlow(Arc arc) { RequireOptimal(); CheckArc(arc); return _flow[arc.Id]; }               // NetworkSimplex.cs:416-431
        public long GetPotential(Node node) { RequireOptimal(); if (!_graph.IsValidNode(node)) throw new ArgumentException("Invalid node"); return _pi[node.Id]; }
        public long GetTotalCost() { RequireOptimal(); return _totalCost; }

        /// <summary>SolutionValidator.Validate() on the device; 0 = valid, else bits 1 conservation, 2 bounds, 4 complementary
        /// slackness, 8 dual feasibility, 16 objective, 32 dual objective (SolutionValidator.cs:20-342).</summary>
        public int ValidateOnDevice(out long primal, out long dual) { RequireOptimal(); Check(Native.mcf_validate(_h, out int bad, out primal, out dual)); return bad; }

        private void RequireOptimal() { if (Status != SolverStatus.Optimal) throw new InvalidOperationException("Solution not optimal"); }
        private void CheckArc(Arc arc) { if (!_graph.IsValidArc(arc)) throw new ArgumentException("Invalid arc"); }
        private void Check(int rc)
        {
            if (rc == 0) return;
            string msg = _h != IntPtr.Zero ? Marshal.PtrToStringAnsi(Native.mcf_last_error(_h)) ?? "" : "";
            throw rc switch { -1 => new ArgumentException(msg), -5 => new InvalidOperationException("Solution not optimal"),
                              -2 => new PlatformNotSupportedException("no sm_100 CUDA device (libmcfgpu has no CPU fallback)"),
                              _ => new ExternalException($"libmcfgpu error {rc}: {msg}") };
        }
        public void Dispose() { if (_h != IntPtr.Zero) { Native.mcf_destroy(_h); _h = IntPtr.Zero; } }
    }

    /// <summary>The README-style fluent facade (README.md:48-60) over the real surface.</summary>
    public static class NetworkSimplexFacade
    {
        public static CudaNetworkSimplex SupplyMap(this CudaNetworkSimplex s, Func<Node, long> supply, IGraph g)
        { for (int i = 0; i < g.NodeCount; i++) s.SetNodeSupply(new Node(i), supply(new Node(i))); return s; }
        public static SolverStatus Run(this CudaNetworkSimplex s) => s.Solve();
        public static long Flow(this CudaNetworkSimplex s, Arc a) => s.GetFlow(a);
        public static long Potential(this CudaNetworkSimplex s, Node n) => s.GetPotential(n);
        public static long TotalCost(this CudaNetworkSimplex s) => s.GetTotalCost();
    }
}
