/*
 * netgen.c - NETGEN-style min-cost-flow instance generator (host side, plain C).
 *
 * The reference ships four instances produced by the public-domain DIMACS
 * "NETGEN flow network generator (C version)" (Klingman/Napier/Stutz 1974)
 * but no generator:
 *   /root/reference/src/MinCostFlow.Problems/Resources/netgen/netgen_8_08a.min:1-22
 * BASELINE.json's configs are larger members of the same "NETGEN-8" family, so
 * this file re-implements the published algorithm.  It is written for n = 2^20:
 * the original's O(n) flag-array index lists (rebuilt once per node, O(n^2)
 * overall) are replaced by a Fenwick order-statistics list for the chain
 * shuffle and by "interval minus a short sorted removal list" for the per-node
 * head lists.  Acceptance test: tests/test_netgen.py regenerates the four
 * fixtures byte for byte.
 *
 * Exported C ABI (ctypes / the C++ host layer):
 *   mcfgen_netgen(seed, parms[13], &n_arcs, tail, head, cost, cap, supply)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define P_NODES 0
#define P_SOURCES 1
#define P_SINKS 2
#define P_DENSITY 3
#define P_MINCOST 4
#define P_MAXCOST 5
#define P_SUPPLY 6
#define P_TSOURCES 7
#define P_TSINKS 8
#define P_HICOST 9
#define P_CAPACITATED 10
#define P_MINCAP 11
#define P_MAXCAP 12

/* ---- Lehmer generator x <- 16807 x mod (2^31-1); draw = a + x % (b-a+1) ---- */
typedef struct { int64_t seed; } rng_t;

static int64_t rnd(rng_t *r, int64_t a, int64_t b)
{
    r->seed = (r->seed * 16807) % 2147483647;
    if (b <= a) return b;
    return a + r->seed % (b - a + 1);
}

/* ---- index list A: Fenwick tree over [base, base+size) (k-th remaining) ---- */
typedef struct { int64_t base, size, remaining; int32_t *fw; int logn; } flist_t;

static flist_t *flist_make(int64_t from, int64_t to)
{
    flist_t *l = (flist_t *)calloc(1, sizeof(flist_t));
    if (from <= 0 || from > to) { l->size = 0; return l; }
    l->base = from; l->size = to - from + 1; l->remaining = l->size;
    l->fw = (int32_t *)malloc((size_t)(l->size + 1) * sizeof(int32_t));
    for (int64_t i = 1; i <= l->size; i++) l->fw[i] = (int32_t)(i & -i);
    l->logn = 0; while ((1LL << (l->logn + 1)) <= l->size) l->logn++;
    return l;
}
static void flist_free(flist_t *l) { free(l->fw); free(l); }
static int64_t flist_choose(flist_t *l, int64_t position)
{
    if (position < 1 || position > l->remaining) return 0;
    int64_t idx = 0, k = position;
    for (int b = l->logn; b >= 0; b--) {
        int64_t nx = idx + (1LL << b);
        if (nx <= l->size && l->fw[nx] < k) { idx = nx; k -= l->fw[nx]; }
    }
    idx += 1;                                   /* 1-based slot of the k-th remaining */
    for (int64_t i = idx; i <= l->size; i += i & -i) l->fw[i]--;
    l->remaining--;
    return l->base + idx - 1;
}

/* ---- index list B: [from,to] minus a short sorted removal list ---- */
typedef struct { int64_t from, to, index_size, pseudo_size; int64_t *rem; int nrem, cap; } slist_t;

static void slist_init(slist_t *l, int64_t from, int64_t to)
{
    l->from = from; l->to = to; l->nrem = 0;
    if (from <= 0 || from > to) { l->index_size = l->pseudo_size = 0; l->to = from - 1; return; }
    l->index_size = l->pseudo_size = to - from + 1;
}
static void slist_insert(slist_t *l, int64_t v)
{
    if (l->nrem == l->cap) { l->cap = l->cap ? 2 * l->cap : 64; l->rem = (int64_t *)realloc(l->rem, (size_t)l->cap * sizeof(int64_t)); }
    int i = l->nrem++;
    while (i > 0 && l->rem[i - 1] > v) { l->rem[i] = l->rem[i - 1]; i--; }
    l->rem[i] = v;
}
static int64_t slist_choose(slist_t *l, int64_t position)
{
    if (position < 1 || position > l->index_size) return 0;
    int64_t v = l->from + position - 1;
    for (int i = 0; i < l->nrem && l->rem[i] <= v; i++) v++;
    slist_insert(l, v);
    l->index_size--; l->pseudo_size--;
    return v;
}
static void slist_remove(slist_t *l, int64_t v)
{
    l->pseudo_size--;                       /* also for values that are not in the list */
    if (v < l->from || v > l->to) return;
    for (int i = 0; i < l->nrem; i++) if (l->rem[i] == v) return;
    slist_insert(l, v);
    l->index_size--;
}

/* ---- generator state ---- */
typedef struct {
    const int64_t *parms; rng_t rng;
    int64_t nodes_left, arc_count, max_arcs;
    int32_t *tail, *head; int64_t *cost, *cap;   /* output arcs */
    int64_t *B;                                   /* supplies, 0-based */
    int64_t *sk_tail, *sk_head;                   /* skeleton scratch */
    int overflow;
} gen_t;

static void save_arc(gen_t *g, int64_t t, int64_t h, int64_t c, int64_t u)
{
    if (g->arc_count >= g->max_arcs) { g->overflow = 1; return; }
    g->tail[g->arc_count] = (int32_t)t; g->head[g->arc_count] = (int32_t)h;
    g->cost[g->arc_count] = c; g->cap[g->arc_count] = u; g->arc_count++;
}

static void create_supply(gen_t *g, int64_t sources, int64_t supply)
{
    int64_t per = supply / sources;
    for (int64_t i = 0; i < sources; i++) {
        int64_t part = rnd(&g->rng, 1, per);
        g->B[i] += part;
        g->B[rnd(&g->rng, 0, sources - 1)] += per - part;
    }
    g->B[rnd(&g->rng, 0, sources - 1)] += supply % sources;
}

static void sort_skeleton(gen_t *g, int64_t count)      /* Shell sort by tail over [1..count], as published */
{
    int64_t m = count;
    while ((m /= 2) != 0) {
        int64_t k = count - m;
        for (int64_t j = 1; j <= k; j++) {
            int64_t i = j;
            while (i >= 1 && g->sk_tail[i] > g->sk_tail[i + m]) {
                int64_t t = g->sk_tail[i]; g->sk_tail[i] = g->sk_tail[i + m]; g->sk_tail[i + m] = t;
                t = g->sk_head[i]; g->sk_head[i] = g->sk_head[i + m]; g->sk_head[i + m] = t;
                i -= m;
            }
        }
    }
}

static void pick_head(gen_t *g, slist_t *handle, int64_t desired_tail)
{
    const int64_t *parms = g->parms;
    int64_t non_sources = parms[P_NODES] - parms[P_SOURCES] + parms[P_TSOURCES];
    int64_t remaining_arcs = parms[P_DENSITY] - g->arc_count;
    int64_t limit, upper_bound;

    g->nodes_left--;
    if (2 * g->nodes_left >= remaining_arcs) return;

    if ((remaining_arcs + non_sources - handle->pseudo_size - 1) / (g->nodes_left + 1) >= non_sources - 1) {
        limit = non_sources;
    } else {
        upper_bound = 2 * (remaining_arcs / (g->nodes_left + 1) - 1);
        do {
            limit = rnd(&g->rng, 1, upper_bound);
            if (g->nodes_left == 0) limit = remaining_arcs;
        } while ((double)g->nodes_left * (double)(non_sources - 1) < (double)remaining_arcs - (double)limit);
    }

    for (; limit > 0; limit--) {
        int64_t index = slist_choose(handle, rnd(&g->rng, 1, handle->pseudo_size));
        int64_t cap = parms[P_SUPPLY];
        if (rnd(&g->rng, 1, 100) <= parms[P_CAPACITATED])
            cap = rnd(&g->rng, parms[P_MINCAP], parms[P_MAXCAP]);
        if (1 <= index && index <= parms[P_NODES]) {
            int64_t c = rnd(&g->rng, parms[P_MINCOST], parms[P_MAXCOST]);
            save_arc(g, desired_tail, index, c, cap);
        }
    }
}

/* returns 0 on success, <0 on error: -1 bad seed, -2 bad parameters, -3 arc buffer overflow,
 * -4 assignment-problem special case (not part of the NETGEN-8 family; unsupported) */
int mcfgen_netgen(int64_t seed, const int64_t *parms, int64_t max_arcs, int64_t *n_arcs,
                  int32_t *tail, int32_t *head, int64_t *cost, int64_t *cap, int64_t *supply)
{
    int64_t NODES = parms[P_NODES], SOURCES = parms[P_SOURCES], SINKS = parms[P_SINKS];
    int64_t DENSITY = parms[P_DENSITY], SUPPLY = parms[P_SUPPLY];
    int64_t TSOURCES = parms[P_TSOURCES], TSINKS = parms[P_TSINKS];
    if (seed <= 0) return -1;
    if (NODES <= 0 || NODES > DENSITY || SOURCES <= 0 || SINKS <= 0 || SOURCES + SINKS > NODES ||
        parms[P_MINCOST] > parms[P_MAXCOST] || SUPPLY < SOURCES || TSOURCES > SOURCES || TSINKS > SINKS ||
        parms[P_HICOST] < 0 || parms[P_HICOST] > 100 || parms[P_CAPACITATED] < 0 || parms[P_CAPACITATED] > 100 ||
        parms[P_MINCAP] > parms[P_MAXCAP])
        return -2;
    if ((SOURCES - TSOURCES) + (SINKS - TSINKS) == NODES && (SOURCES - TSOURCES) == (SINKS - TSINKS) && SOURCES == SUPPLY)
        return -4;

    gen_t g; memset(&g, 0, sizeof(g));
    g.parms = parms; g.rng.seed = seed;
    g.max_arcs = max_arcs; g.tail = tail; g.head = head; g.cost = cost; g.cap = cap; g.B = supply;
    g.nodes_left = NODES - SINKS + TSINKS;
    memset(supply, 0, (size_t)NODES * sizeof(int64_t));
    int64_t *pred = (int64_t *)calloc((size_t)NODES + 2, sizeof(int64_t));
    g.sk_tail = (int64_t *)calloc((size_t)NODES + SINKS + 4, sizeof(int64_t));
    g.sk_head = (int64_t *)calloc((size_t)NODES + SINKS + 4, sizeof(int64_t));

    create_supply(&g, SOURCES, SUPPLY);

    /* distribute the transshipment nodes over SOURCES chains */
    for (int64_t i = 1; i <= SOURCES; i++) pred[i] = i;
    flist_t *fl = flist_make(SOURCES + 1, NODES - SINKS);
    int64_t source = 1, T = NODES - SOURCES - SINKS, i;
    for (i = T; i > (4 * T + 9) / 10; i--) {
        int64_t node = flist_choose(fl, rnd(&g.rng, 1, fl->remaining));
        pred[node] = pred[source]; pred[source] = node;
        if (++source > SOURCES) source = 1;
    }
    for (; i > 0; --i) {
        int64_t node = flist_choose(fl, rnd(&g.rng, 1, fl->remaining));
        source = rnd(&g.rng, 1, SOURCES);
        pred[node] = pred[source]; pred[source] = node;
    }
    flist_free(fl);

    slist_t sl; memset(&sl, 0, sizeof(sl));
    int64_t *sinks = (int64_t *)malloc((size_t)(2 * SINKS + 4) * sizeof(int64_t));
    for (source = 1; source <= SOURCES; source++) {
        int64_t sort_count = 0, node = pred[source];
        while (node != source) {                 /* chain arcs pred[node] -> node */
            sort_count++;
            g.sk_head[sort_count] = node;
            node = g.sk_tail[sort_count] = pred[node];
        }
        int64_t sinks_per_source;
        if (T == 0) sinks_per_source = SINKS / SOURCES + 1;
        else sinks_per_source = (int64_t)(((double)2 * (double)sort_count * (double)SINKS) / (double)T);
        if (sinks_per_source > SINKS) sinks_per_source = SINKS;
        if (sinks_per_source < 2) sinks_per_source = 2;
        slist_init(&sl, NODES - SINKS, NODES - 1);
        for (i = 0; i < sinks_per_source; i++)
            sinks[i] = slist_choose(&sl, rnd(&g.rng, 1, sl.index_size));
        if (source == SOURCES && sl.index_size > 0) {
            while (sl.index_size > 0) {
                int64_t j = slist_choose(&sl, 1);
                if (g.B[j] == 0) sinks[sinks_per_source++] = j;
            }
        }

        int64_t chain_length = sort_count;
        int64_t supply_per_sink = g.B[source - 1] / sinks_per_source;
        int64_t k = pred[source];
        for (i = 0; i < sinks_per_source; i++) {
            sort_count++;
            int64_t partial = rnd(&g.rng, 1, supply_per_sink);
            int64_t j = rnd(&g.rng, 0, sinks_per_source - 1);
            g.sk_tail[sort_count] = k;
            g.sk_head[sort_count] = sinks[i] + 1;
            g.B[sinks[i]] -= partial;
            g.B[sinks[j]] -= supply_per_sink - partial;
            k = source;
            for (j = rnd(&g.rng, 1, chain_length); j > 0; j--) k = pred[k];
        }
        g.B[sinks[0]] -= g.B[source - 1] % sinks_per_source;

        sort_skeleton(&g, sort_count);
        g.sk_tail[sort_count + 1] = 0;
        for (i = 1; i <= sort_count;) {
            slist_init(&sl, SOURCES - TSOURCES + 1, NODES);
            slist_remove(&sl, g.sk_tail[i]);
            int64_t it = g.sk_tail[i];
            while (it == g.sk_tail[i]) {
                slist_remove(&sl, g.sk_head[i]);
                int64_t cap_ = SUPPLY;
                if (rnd(&g.rng, 1, 100) <= parms[P_CAPACITATED]) {
                    cap_ = g.B[source - 1];
                    if (cap_ < parms[P_MINCAP]) cap_ = parms[P_MINCAP];
                }
                int64_t cost_ = parms[P_MAXCOST];
                if (rnd(&g.rng, 1, 100) > parms[P_HICOST])
                    cost_ = rnd(&g.rng, parms[P_MINCOST], parms[P_MAXCOST]);
                save_arc(&g, it, g.sk_head[i], cost_, cap_);
                i++;
            }
            pick_head(&g, &sl, it);
        }
    }

    for (i = NODES - SINKS + 1; i <= NODES - SINKS + TSINKS; i++) {
        slist_init(&sl, SOURCES - TSOURCES + 1, NODES);
        slist_remove(&sl, i);
        pick_head(&g, &sl, i);
    }

    free(sinks); free(sl.rem); free(pred); free(g.sk_tail); free(g.sk_head);
    *n_arcs = g.arc_count;
    return g.overflow ? -3 : 0;
}
