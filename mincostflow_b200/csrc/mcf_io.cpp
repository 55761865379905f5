// mcf_io.cpp - bulk DIMACS I/O of libmcfgpu: the step before and after the solve.
//
//   mcf_dimacs_open / _dims / _copy / _close, mcf_create_from_dimacs
//        DimacsReader.ReadFromStream (src/MinCostFlow.Problems/Loaders/DimacsReader.cs:36-147) followed by the per-element
//        setter loop every caller of the reference runs (Benchmarks/NetworkSimplexBenchmarks.cs:166-189): here one parallel
//        pass over the text straight into the flat arrays the engine uploads (arc ids = order of the `a` lines,
//        DimacsReader.cs:118-121 -> GraphBuilder.AddArc order).
//   mcf_write_solution, mcf_read_solution
//        SolutionLoader.SaveToFile / LoadFromStream (Loaders/SolutionLoader.cs:186-214, :69-173).
//
// Host code only; nothing here touches the device.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mcfgpu.h"

namespace {

thread_local std::string g_io_error;

int io_fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_io_error = buf;
    return code;
}

bool read_file(const char* path, std::string* out)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = sz > 0 ? fread(&(*out)[0], 1, (size_t)sz, f) : 0;
    fclose(f);
    return got == out->size();
}

// long.Parse(token, InvariantCulture) on a token that holds no blanks: optional sign, decimal digits, overflow is an error
bool parse_i64(const char* b, const char* e, int64_t* out)
{
    if (b == e) return false;
    bool neg = false;
    if (*b == '-' || *b == '+') { neg = *b == '-'; ++b; if (b == e) return false; }
    uint64_t v = 0;
    const uint64_t lim = neg ? (uint64_t)1 << 63 : (uint64_t)std::numeric_limits<int64_t>::max();
    for (; b != e; ++b) {
        const unsigned d = (unsigned)(*b - '0');
        if (d > 9) return false;
        if (v > (lim - d) / 10) return false;
        v = v * 10 + d;
    }
    *out = neg ? (int64_t)(0 - v) : (int64_t)v;
    return true;
}
bool parse_i32(const char* b, const char* e, int64_t* out)
{
    return parse_i64(b, e, out) && *out >= std::numeric_limits<int32_t>::min() && *out <= std::numeric_limits<int32_t>::max();
}

struct ArcLine { int32_t from, to; int64_t lower, upper, cost; };

struct Part {                               // what one thread found in its slice of the text
    std::vector<ArcLine> arcs;
    std::vector<std::pair<int32_t, int64_t>> supplies;      // in file order: a later `n` line overrides (DimacsReader.cs:92)
    bool has_p = false; int64_t n = 0, m = 0;
    int p_lines = 0;
    std::string error;                      // first malformed line of the slice
};

// line.Split(' ', RemoveEmptyEntries); tabs and '\r' are treated as blanks too
int tokenize(const char* b, const char* e, const char* tb[], const char* te[], int max_tok)
{
    int nt = 0;
    while (b < e) {
        while (b < e && (*b == ' ' || *b == '\t' || *b == '\r')) ++b;
        if (b >= e) break;
        const char* s = b;
        while (b < e && *b != ' ' && *b != '\t' && *b != '\r') ++b;
        if (nt < max_tok) { tb[nt] = s; te[nt] = b; }
        ++nt;
    }
    return nt;
}

void parse_slice(const char* b, const char* e, Part* out)
{
    const char* tb[8]; const char* te[8];
    while (b < e && out->error.empty()) {
        const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
        const char* le = nl ? nl : e;
        const int nt = tokenize(b, le, tb, te, 8);
        if (nt > 0 && te[0] - tb[0] == 1) {
            auto line = [&] { return std::string(b, le); };
            switch (*tb[0]) {
                case 'p': {                                                                     // DimacsReader.cs:67-81
                    int64_t n, m;
                    if (nt != 4 || te[1] - tb[1] != 3 || memcmp(tb[1], "min", 3) != 0) { out->error = "Invalid problem line: " + line(); break; }
                    if (!parse_i32(tb[2], te[2], &n) || !parse_i32(tb[3], te[3], &m) || n < 0 || m < 0) { out->error = "Invalid problem line: " + line(); break; }
                    out->has_p = true; out->n = n; out->m = m; out->p_lines++;
                    break;
                }
                case 'n': {                                                                     // DimacsReader.cs:83-93
                    int64_t id, s;
                    if (nt != 3 || !parse_i32(tb[1], te[1], &id) || !parse_i64(tb[2], te[2], &s)) { out->error = "Invalid node line: " + line(); break; }
                    out->supplies.emplace_back((int32_t)(id - 1), s);
                    break;
                }
                case 'a': {                                                                     // DimacsReader.cs:95-109
                    int64_t f, t, lo, up, c;
                    if (nt != 6 || !parse_i32(tb[1], te[1], &f) || !parse_i32(tb[2], te[2], &t) || !parse_i64(tb[3], te[3], &lo) ||
                        !parse_i64(tb[4], te[4], &up) || !parse_i64(tb[5], te[5], &c)) { out->error = "Invalid arc line: " + line(); break; }
                    out->arcs.push_back(ArcLine{(int32_t)(f - 1), (int32_t)(t - 1), lo, up, c});
                    break;
                }
                default: break;                                                                 // `c` and unknown line types are skipped (:62-64, :111-113)
            }
        }
        b = nl ? nl + 1 : e;
    }
}

}  // namespace

struct mcf_dimacs {
    int32_t n = 0, m = 0;
    std::vector<int32_t> source, target;
    std::vector<int64_t> lower, upper, cost, supply;
};

extern "C" {

const char* mcf_io_last_error(void) { return g_io_error.c_str(); }

int mcf_dimacs_parse(const char* text, int64_t length, mcf_dimacs** out)
{
    if (!text || length < 0 || !out) return MCF_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    const char* b = text; const char* e = text + length;
    unsigned hw = std::thread::hardware_concurrency();
    int T = (int)std::min<unsigned>(hw ? hw : 1, 32);
    if (length < (1 << 20)) T = 1;
    std::vector<const char*> cut(T + 1);
    cut[0] = b; cut[T] = e;
    for (int i = 1; i < T; ++i) {
        const char* p = b + (size_t)((double)length * i / T);
        if (p < cut[i - 1]) p = cut[i - 1];
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        cut[i] = nl ? nl + 1 : e;
    }
    std::vector<Part> parts(T);
    if (T == 1) parse_slice(cut[0], cut[1], &parts[0]);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < T; ++i) th.emplace_back(parse_slice, cut[i], cut[i + 1], &parts[i]);
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < T; ++i) if (!parts[i].error.empty()) return io_fail(MCF_ERR_FORMAT, "%s", parts[i].error.c_str());      // FormatException
    auto* d = new mcf_dimacs();
    bool has_p = false; int64_t n = 0, m_decl = 0;
    for (int i = 0; i < T; ++i) if (parts[i].has_p) { has_p = true; n = parts[i].n; m_decl = parts[i].m; }      // the last `p` line wins (:75-76)
    size_t m = 0;
    for (int i = 0; i < T; ++i) m += parts[i].arcs.size();
    // The reference sizes its arc arrays by the declared count (:129-131) and fills them from the `a` lines (:140-146): more lines
    // than declared is an IndexOutOfRangeException there, fewer leaves a graph and arrays of different lengths.  Both are refused.
    if (!has_p && (m > 0 || std::any_of(parts.begin(), parts.end(), [](const Part& p) { return !p.supplies.empty(); }))) {
        delete d; return io_fail(MCF_ERR_FORMAT, "no problem line (p min NODES ARCS)");
    }
    if ((int64_t)m != m_decl) { delete d; return io_fail(MCF_ERR_FORMAT, "problem line declares %lld arcs, the file holds %zu", (long long)m_decl, m); }
    d->n = (int32_t)n; d->m = (int32_t)m;
    d->source.resize(m); d->target.resize(m); d->lower.resize(m); d->upper.resize(m); d->cost.resize(m);
    d->supply.assign((size_t)n, 0);
    std::vector<size_t> off(T + 1, 0);
    for (int i = 0; i < T; ++i) off[i + 1] = off[i] + parts[i].arcs.size();
    std::atomic<int> bad{-1};
    auto fill = [&](int i) {
        size_t k = off[i];
        for (const ArcLine& a : parts[i].arcs) {
            if (a.from < 0 || a.from >= n || a.to < 0 || a.to >= n) bad.store(i);
            d->source[k] = a.from; d->target[k] = a.to; d->lower[k] = a.lower; d->upper[k] = a.upper; d->cost[k] = a.cost; ++k;
        }
    };
    if (T == 1) fill(0);
    else { std::vector<std::thread> th; for (int i = 0; i < T; ++i) th.emplace_back(fill, i); for (auto& t : th) t.join(); }
    if (bad.load() >= 0) { delete d; return io_fail(MCF_ERR_FORMAT, "arc endpoint outside 1..%lld", (long long)n); }             // GraphBuilder.AddArc throws
    for (int i = 0; i < T; ++i)
        for (const auto& s : parts[i].supplies) {
            if (s.first < 0 || s.first >= n) { delete d; return io_fail(MCF_ERR_FORMAT, "node id %d outside 1..%lld", s.first + 1, (long long)n); }
            d->supply[s.first] = s.second;
        }
    *out = d;
    return MCF_OK;
}

int mcf_dimacs_open(const char* path, mcf_dimacs** out)
{
    if (!path || !out) return MCF_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    std::string text;
    if (!read_file(path, &text)) return io_fail(MCF_ERR_IO, "cannot read %s: %s", path, strerror(errno));
    return mcf_dimacs_parse(text.data(), (int64_t)text.size(), out);
}

int mcf_dimacs_dims(const mcf_dimacs* d, int32_t* n_out, int32_t* m_out)
{
    if (!d) return MCF_ERR_INVALID_ARGUMENT;
    if (n_out) *n_out = d->n;
    if (m_out) *m_out = d->m;
    return MCF_OK;
}

int mcf_dimacs_copy(const mcf_dimacs* d, int32_t* source, int32_t* target, int64_t* lower, int64_t* upper, int64_t* cost, int64_t* supply)
{
    if (!d) return MCF_ERR_INVALID_ARGUMENT;
    const size_t m = (size_t)d->m, n = (size_t)d->n;
    if (source) memcpy(source, d->source.data(), m * 4);
    if (target) memcpy(target, d->target.data(), m * 4);
    if (lower) memcpy(lower, d->lower.data(), m * 8);
    if (upper) memcpy(upper, d->upper.data(), m * 8);
    if (cost) memcpy(cost, d->cost.data(), m * 8);
    if (supply) memcpy(supply, d->supply.data(), n * 8);
    return MCF_OK;
}

void mcf_dimacs_close(mcf_dimacs* d) { delete d; }

int mcf_create_from_dimacs(const char* path, mcf_handle** out)
{
    if (!out) return MCF_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    mcf_dimacs* d = nullptr;
    int rc = mcf_dimacs_open(path, &d);
    if (rc != MCF_OK) return rc;
    mcf_handle* h = nullptr;
    rc = mcf_create(d->n, d->m, d->source.data(), d->target.data(), &h);
    if (rc == MCF_OK) rc = mcf_set_arcs(h, d->lower.data(), d->upper.data(), d->cost.data());
    if (rc == MCF_OK) rc = mcf_set_supply(h, d->supply.data());
    if (rc != MCF_OK) { if (h) mcf_destroy(h); h = nullptr; io_fail(rc, "mcf_create failed for %s (%d)", path, rc); }
    mcf_dimacs_close(d);
    *out = h;
    return rc;
}

int mcf_write_solution(mcf_handle* h, const char* path, int32_t format, int32_t with_potentials)
{
    if (!h || !path || format < 0 || format > 1) return MCF_ERR_INVALID_ARGUMENT;
    int64_t cost = 0;
    int rc = mcf_get_total_cost(h, &cost);
    if (rc != MCF_OK) return rc;                                            // MCF_ERR_NOT_OPTIMAL unless Optimal
    int32_t n = 0, m = 0;
    rc = mcf_get_dims(h, &n, &m);
    if (rc != MCF_OK) return rc;
    std::vector<int64_t> flow((size_t)m), pi((size_t)n);
    std::vector<int32_t> src, tgt;
    if (m > 0 && (rc = mcf_get_flows(h, flow.data())) != MCF_OK) return rc;
    if (with_potentials && n > 0 && (rc = mcf_get_potentials(h, pi.data())) != MCF_OK) return rc;
    if (format == 1) { src.resize(m); tgt.resize(m); if ((rc = mcf_get_endpoints(h, src.data(), tgt.data())) != MCF_OK) return rc; }
    std::string text;
    text.reserve((size_t)m * 12 + 64);
    char buf[96];
    text.append(buf, (size_t)snprintf(buf, sizeof(buf), "s %lld\n", (long long)cost));                                  // SolutionLoader.cs:191
    for (int32_t e = 0; e < m; ++e) {
        if (flow[e] == 0) continue;                                                                                         // :196
        if (format == 0) text.append(buf, (size_t)snprintf(buf, sizeof(buf), "f %d %lld\n", e, (long long)flow[e]));        // f ARC_ID FLOW (:198)
        else text.append(buf, (size_t)snprintf(buf, sizeof(buf), "f %d %d %lld\n", src[e] + 1, tgt[e] + 1, (long long)flow[e]));   // f SRC DST FLOW, 1-based (:124-129)
    }
    if (with_potentials) for (int32_t u = 0; u < n; ++u) text.append(buf, (size_t)snprintf(buf, sizeof(buf), "p %d %lld\n", u, (long long)pi[u]));   // :203-209
    FILE* f = fopen(path, "wb");
    if (!f) return io_fail(MCF_ERR_IO, "cannot write %s: %s", path, strerror(errno));
    const size_t put = fwrite(text.data(), 1, text.size(), f);
    if (fclose(f) != 0 || put != text.size()) return io_fail(MCF_ERR_IO, "short write to %s", path);
    return MCF_OK;
}

int mcf_read_solution(const char* path, int64_t* cost_out, int32_t capacity, int32_t* a_out, int32_t* b_out, int64_t* flow_out,
                      int32_t* flow_lines_out, int32_t* endpoint_form_out)
{
    if (!path) return MCF_ERR_INVALID_ARGUMENT;
    std::string text;
    if (!read_file(path, &text)) return io_fail(MCF_ERR_IO, "cannot read %s: %s", path, strerror(errno));
    const char* b = text.data(); const char* e = b + text.size();
    const char* tb[8]; const char* te[8];
    int64_t cost = 0; int32_t lines = 0; int32_t endpoint = 0;
    while (b < e) {
        const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
        const char* le = nl ? nl : e;
        const int nt = tokenize(b, le, tb, te, 8);
        if (nt >= 2 && te[0] - tb[0] == 1 && *tb[0] == 's') {                                                             // SolutionLoader.cs:107-113
            if (!parse_i64(tb[1], te[1], &cost)) return io_fail(MCF_ERR_FORMAT, "Invalid solution line: %s", std::string(b, le).c_str());
        } else if (nt >= 3 && te[0] - tb[0] == 1 && *tb[0] == 'f') {                                                      // :115-143
            int64_t x, y, fl;
            if (nt == 3) {
                if (!parse_i32(tb[1], te[1], &x) || !parse_i64(tb[2], te[2], &fl)) return io_fail(MCF_ERR_FORMAT, "Invalid flow line: %s", std::string(b, le).c_str());
                y = -1;
            } else {
                if (!parse_i32(tb[1], te[1], &x) || !parse_i32(tb[2], te[2], &y) || !parse_i64(tb[3], te[3], &fl)) return io_fail(MCF_ERR_FORMAT, "Invalid flow line: %s", std::string(b, le).c_str());
                x -= 1; y -= 1; endpoint = 1;
            }
            if (lines < capacity) { if (a_out) a_out[lines] = (int32_t)x; if (b_out) b_out[lines] = (int32_t)y; if (flow_out) flow_out[lines] = fl; }
            ++lines;
        }
        b = nl ? nl + 1 : e;
    }
    if (cost == 0 && endpoint) cost = std::numeric_limits<int64_t>::min();                                                // "cost not specified" marker (:165-170)
    if (cost_out) *cost_out = cost;
    if (flow_lines_out) *flow_lines_out = lines;
    if (endpoint_form_out) *endpoint_form_out = endpoint;
    return MCF_OK;
}

}  // extern "C"
