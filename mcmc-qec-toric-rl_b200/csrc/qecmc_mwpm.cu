// qecmc_mwpm.cu -- class-sorted minimum-weight perfect matching start states for planar chains (host code).
//
// Replaces the reference's src/mwpm.py for Planar_code: MWPM.generate_edges (:66-133), generate_edges_constrained
// (:136-229), eliminate_defect_pair (:232-288), eliminate_border_defect (:291-316), solve_layer (:319-373), solve
// (:408-415), generate_classes (:417-438), class_sorted_mwpm (:462-475) and regular_mwpm (:479-487).  The reference writes
// the defect graph to a text file and runs the external blossom5 binary on it (:376-405); here the matching is solved in
// process by a dense O(n^3) primal-dual blossom algorithm, one syndrome per host thread.  The optimisation problems are the
// reference's own (its graphs with the ancillas folded away, solve_layer_reduced below; mode | 2 solves the graphs as the
// reference writes them), so the matching WEIGHT is the reference's; which of several minimum-weight matchings comes out
// is the solver's choice there and here.
//
// This is an initialiser that runs once per syndrome before the chains start (decoders.py:272-279 take the list of
// per-class codes it returns); it is not on the Metropolis path and stays on the host like the reference's.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <deque>
#include <thread>
#include <vector>

#include "qecmc_internal.h"

namespace qecmc {
namespace {

// Maximum-weight matching in a general graph with integer weights, dense O(n^3) primal-dual blossom algorithm
// (Edmonds; Galil's exposition).  Vertices are 1..n, blossoms n+1..2n.  An edge exists iff its weight is > 0.
class Blossom {
  public:
    explicit Blossom(int n_) : n(n_), nx(n_), N(2 * n_ + 1)
    {
        g.assign((size_t)N * N, Edge{0, 0, 0});
        for (int u = 0; u < N; u++)
            for (int v = 0; v < N; v++) at(u, v) = Edge{u, v, 0};
        lab.assign(N, 0); match.assign(N, 0); slack.assign(N, 0); st.assign(N, 0); pa.assign(N, 0);
        S.assign(N, -1); vis.assign(N, 0); flower.assign(N, {});
        from.assign((size_t)N * (n + 1), 0);
    }
    void add_edge(int u, int v, int w) { at(u, v).w = w; at(v, u).w = w; }   // 1-based, w > 0
    // returns the matching's total weight; mate(u) = partner or 0
    long long solve()
    {
        std::fill(match.begin(), match.end(), 0);
        nx = n;
        int wmax = 0;
        for (int u = 0; u <= n; u++) { st[u] = u; flower[u].clear(); }
        for (int u = 1; u <= n; u++)
            for (int v = 1; v <= n; v++) { fr(u, v) = (u == v ? u : 0); wmax = std::max(wmax, at(u, v).w); }
        for (int u = 1; u <= n; u++) lab[u] = wmax;
        while (phase()) {}
        long long tot = 0;
        for (int u = 1; u <= n; u++)
            if (match[u] && match[u] < u) tot += at(u, match[u]).w;
        return tot;
    }
    int mate(int u) const { return match[u]; }

  private:
    struct Edge { int u, v, w; };
    int n, nx, N, tick = 0;
    std::vector<Edge> g;
    std::vector<long long> lab;
    std::vector<int> match, slack, st, pa, S, vis, from;
    std::vector<std::vector<int>> flower;
    std::deque<int> q;

    Edge &at(int u, int v) { return g[(size_t)u * N + v]; }
    int &fr(int b, int x) { return from[(size_t)b * (n + 1) + x]; }
    long long delta(const Edge &e) { return lab[e.u] + lab[e.v] - 2ll * at(e.u, e.v).w; }
    void update_slack(int u, int x) { if (!slack[x] || delta(at(u, x)) < delta(at(slack[x], x))) slack[x] = u; }
    void set_slack(int x)
    {
        slack[x] = 0;
        for (int u = 1; u <= n; u++)
            if (at(u, x).w > 0 && st[u] != x && S[st[u]] == 0) update_slack(u, x);
    }
    void q_push(int x)
    {
        if (x <= n) q.push_back(x);
        else for (int i : flower[x]) q_push(i);
    }
    void set_st(int x, int b)
    {
        st[x] = b;
        if (x > n) for (int i : flower[x]) set_st(i, b);
    }
    int get_pr(int b, int xr)
    {
        int pr = (int)(std::find(flower[b].begin(), flower[b].end(), xr) - flower[b].begin());
        if (pr % 2 == 1) { std::reverse(flower[b].begin() + 1, flower[b].end()); return (int)flower[b].size() - pr; }
        return pr;
    }
    void set_match(int u, int v)
    {
        match[u] = at(u, v).v;
        if (u <= n) return;
        Edge e = at(u, v);
        int xr = fr(u, e.u), pr = get_pr(u, xr);
        for (int i = 0; i < pr; i++) set_match(flower[u][i], flower[u][i ^ 1]);
        set_match(xr, v);
        std::rotate(flower[u].begin(), flower[u].begin() + pr, flower[u].end());
    }
    void augment(int u, int v)
    {
        for (;;) {
            int xnv = st[match[u]];
            set_match(u, v);
            if (!xnv) return;
            set_match(xnv, st[pa[xnv]]);
            u = st[pa[xnv]]; v = xnv;
        }
    }
    int get_lca(int u, int v)
    {
        for (++tick; u || v; std::swap(u, v)) {
            if (u == 0) continue;
            if (vis[u] == tick) return u;
            vis[u] = tick;
            u = st[match[u]];
            if (u) u = st[pa[u]];
        }
        return 0;
    }
    void add_blossom(int u, int lca, int v)
    {
        int b = n + 1;
        while (b <= nx && st[b]) ++b;
        if (b > nx) ++nx;
        lab[b] = 0; S[b] = 0;
        match[b] = match[lca];
        flower[b].clear();
        flower[b].push_back(lca);
        for (int x = u, y; x != lca; x = st[pa[y]]) { flower[b].push_back(x); flower[b].push_back(y = st[match[x]]); q_push(y); }
        std::reverse(flower[b].begin() + 1, flower[b].end());
        for (int x = v, y; x != lca; x = st[pa[y]]) { flower[b].push_back(x); flower[b].push_back(y = st[match[x]]); q_push(y); }
        set_st(b, b);
        for (int x = 1; x <= nx; x++) at(b, x).w = at(x, b).w = 0;
        for (int x = 1; x <= n; x++) fr(b, x) = 0;
        for (int xs : flower[b]) {
            for (int x = 1; x <= nx; x++)
                if (at(b, x).w == 0 || delta(at(xs, x)) < delta(at(b, x))) { at(b, x) = at(xs, x); at(x, b) = at(x, xs); }
            for (int x = 1; x <= n; x++)
                if (fr(xs, x)) fr(b, x) = xs;
        }
        set_slack(b);
    }
    void expand_blossom(int b)
    {
        for (int i : flower[b]) set_st(i, i);
        int xr = fr(b, at(b, pa[b]).u), pr = get_pr(b, xr);
        for (int i = 0; i < pr; i += 2) {
            int xs = flower[b][i], xns = flower[b][i + 1];
            pa[xs] = at(xns, xs).u;
            S[xs] = 1; S[xns] = 0;
            slack[xs] = 0; set_slack(xns);
            q_push(xns);
        }
        S[xr] = 1; pa[xr] = pa[b];
        for (size_t i = pr + 1; i < flower[b].size(); i++) { int xs = flower[b][i]; S[xs] = -1; set_slack(xs); }
        st[b] = 0;
    }
    bool on_found_edge(const Edge &e)
    {
        int u = st[e.u], v = st[e.v];
        if (S[v] == -1) {
            pa[v] = e.u; S[v] = 1;
            int nu = st[match[v]];
            slack[v] = slack[nu] = 0;
            S[nu] = 0; q_push(nu);
        } else if (S[v] == 0) {
            int lca = get_lca(u, v);
            if (!lca) { augment(u, v); augment(v, u); return true; }
            add_blossom(u, lca, v);
        }
        return false;
    }
    bool phase()
    {
        std::fill(S.begin(), S.begin() + nx + 1, -1);
        std::fill(slack.begin(), slack.begin() + nx + 1, 0);
        q.clear();
        for (int x = 1; x <= nx; x++)
            if (st[x] == x && !match[x]) { pa[x] = 0; S[x] = 0; q_push(x); }
        if (q.empty()) return false;
        for (;;) {
            while (!q.empty()) {
                int u = q.front(); q.pop_front();
                if (S[st[u]] == 1) continue;
                for (int v = 1; v <= n; v++)
                    if (at(u, v).w > 0 && st[u] != st[v]) {
                        if (delta(at(u, v)) == 0) { if (on_found_edge(at(u, v))) return true; }
                        else update_slack(u, st[v]);
                    }
            }
            long long d = INT64_MAX;
            for (int b = n + 1; b <= nx; b++)
                if (st[b] == b && S[b] == 1) d = std::min(d, lab[b] / 2);
            for (int x = 1; x <= nx; x++)
                if (st[x] == x && slack[x]) {
                    if (S[x] == -1) d = std::min(d, delta(at(slack[x], x)));
                    else if (S[x] == 0) d = std::min(d, delta(at(slack[x], x)) / 2);
                }
            for (int u = 1; u <= n; u++) {
                if (S[st[u]] == 0) { if (lab[u] <= d) return false; lab[u] -= d; }
                else if (S[st[u]] == 1) lab[u] += d;
            }
            for (int b = n + 1; b <= nx; b++)
                if (st[b] == b) {
                    if (S[b] == 0) lab[b] += d * 2;
                    else if (S[b] == 1) lab[b] -= d * 2;
                }
            q.clear();
            for (int x = 1; x <= nx; x++)
                if (st[x] == x && slack[x] && st[slack[x]] != x && delta(at(slack[x], x)) == 0)
                    if (on_found_edge(at(slack[x], x))) return true;
            for (int b = n + 1; b <= nx; b++)
                if (st[b] == b && S[b] == 1 && lab[b] == 0) expand_blossom(b);
        }
    }
};

struct Coord { int r, c; };
struct GEdge { int a, b, w; };      // 0-based nodes, distance
struct Graph {
    int nodes = 0, ndef = 0;
    std::vector<GEdge> edges;
    std::vector<int> ancilla_side;  // side of ancilla node (index - ndef)
};

// minimum-weight perfect matching of a graph given as distances: pairs (a < b) and the total distance; false if the
// graph has no perfect matching
static bool min_weight_perfect_matching(const Graph &gr, std::vector<std::pair<int, int>> &pairs, long long &weight)
{
    pairs.clear();
    weight = 0;
    if (gr.nodes == 0) return true;
    int wmax = 0;
    for (const GEdge &e : gr.edges) wmax = std::max(wmax, e.w);
    const int big = wmax * (gr.nodes / 2 + 1) + 1;   // every matched edge is worth more than any saving in distance
    Blossom bl(gr.nodes);
    for (const GEdge &e : gr.edges) bl.add_edge(e.a + 1, e.b + 1, big - e.w);
    bl.solve();
    std::vector<int> dist((size_t)gr.nodes * gr.nodes, -1);
    for (const GEdge &e : gr.edges) dist[(size_t)e.a * gr.nodes + e.b] = dist[(size_t)e.b * gr.nodes + e.a] = e.w;
    for (int u = 1; u <= gr.nodes; u++) {
        int m = bl.mate(u);
        if (!m) return false;
        if (u < m) { pairs.push_back({u - 1, m - 1}); weight += dist[(size_t)(u - 1) * gr.nodes + (m - 1)]; }
    }
    return true;
}

static inline void connect_all(int count, int offset, int w, std::vector<GEdge> &out)   // mwpm.py:448-458
{
    for (int i = 0; i < count; i++)
        for (int j = i + 1; j < count; j++) out.push_back({i + offset, j + offset, w});
}

static inline int manhattan(const Coord &a, const Coord &b) { return abs(a.r - b.r) + abs(a.c - b.c); }   // mwpm.py:442-444

// MWPM.generate_edges for Planar_code (mwpm.py:66-133): every defect gets an ancilla of its own on its nearest border;
// ancillas are connected to each other at distance 0
static Graph edges_free(const std::vector<Coord> &def, int layer, int L)
{
    Graph gr;
    const int nd = (int)def.size();
    gr.ndef = nd;
    gr.nodes = 2 * nd;
    gr.ancilla_side.assign(nd, 0);
    for (int i = 0; i < nd; i++)
        for (int j = i + 1; j < nd; j++) gr.edges.push_back({i, j, manhattan(def[i], def[j])});
    connect_all(nd, nd, 0, gr.edges);
    for (int s = 0; s < nd; s++) {
        int distance = (layer == 0 ? def[s].r : def[s].c) + 1;
        if (distance * 2 < L) gr.ancilla_side[s] = 0;
        else { gr.ancilla_side[s] = 1; distance = L - distance; }
        gr.edges.push_back({s, s + nd, distance});
    }
    return gr;
}

// MWPM.generate_edges_constrained (mwpm.py:136-229): ancillas of the two borders are connected among themselves only, so
// the parity of the number of chains ending on each border is fixed; parity 1 adds one ancilla per border and flips it
static Graph edges_constrained(const std::vector<Coord> &def, int layer, int L, int parity)
{
    Graph gr;
    const int nd = (int)def.size();
    gr.ndef = nd;
    gr.nodes = 2 * nd;
    for (int i = 0; i < nd; i++)
        for (int j = i + 1; j < nd; j++) gr.edges.push_back({i, j, manhattan(def[i], def[j])});
    std::vector<int> nearest(nd), bdist(nd);
    int n_anc[2] = {0, 0};
    for (int s = 0; s < nd; s++) {
        const int b0 = (layer == 0 ? def[s].r : def[s].c) + 1;
        nearest[s] = b0 * 2 > L;
        bdist[s] = nearest[s] ? L - b0 : b0;
        n_anc[nearest[s]]++;
    }
    if (parity == 1) {
        gr.ancilla_side.assign(nd + 2, 0);
        std::vector<GEdge> parity_edges;
        for (int b = 0; b < 2; b++) {
            if (n_anc[b] == 0) {   // a border no defect is nearest to: every defect may reach its node the long way
                parity_edges.clear();
                for (int s = 0; s < nd; s++) parity_edges.push_back({s, nd + (nd + 1) * b, L - bdist[s]});
                gr.ancilla_side[(nd + 1) * b] = b;
            }
            n_anc[b]++;
        }
        gr.nodes += 2;
        gr.edges.insert(gr.edges.end(), parity_edges.begin(), parity_edges.end());
    } else {
        gr.ancilla_side.assign(nd, 0);
    }
    for (int b = 0; b < 2; b++) connect_all(n_anc[b], nd + b * n_anc[0], 0, gr.edges);
    int counts[2] = {0, 0};
    for (int s = 0; s < nd; s++) {
        const int b = nearest[s];
        const int end = nd + b * n_anc[0] + counts[b];
        gr.ancilla_side[end - nd] = b;
        counts[b]++;
        gr.edges.push_back({s, end, bdist[s]});
    }
    return gr;
}

struct PlanarMwpm {
    int L;
    std::vector<Coord> def[2];

    inline size_t idx(int layer, int r, int c) const { return ((size_t)layer * L + r) * L + c; }

    // MWPM.eliminate_defect_pair, planar branch (mwpm.py:232-288): down the start column, then along the end row
    void eliminate_pair(const Coord &a, const Coord &b, int layer, uint8_t *corr) const
    {
        const uint8_t op = layer == 0 ? 3 : 1;
        const int top = std::min(a.r, b.r), bot = std::max(a.r, b.r), left = std::min(a.c, b.c), right = std::max(a.c, b.c);
        for (int i = top; i < bot; i++) corr[idx(layer, i + (layer == 0), a.c)] ^= op;
        for (int i = left; i < right; i++) corr[idx(!layer, b.r, i + layer)] ^= op;
    }
    // MWPM.eliminate_border_defect (mwpm.py:291-316)
    void eliminate_border(const Coord &a, int layer, int border, uint8_t *corr) const
    {
        const uint8_t op = layer == 0 ? 3 : 1;
        if (layer == 0) {
            if (border == 0) for (int i = 0; i <= a.r; i++) corr[idx(0, i, a.c)] ^= op;
            else for (int i = a.r + 1; i < L; i++) corr[idx(0, i, a.c)] ^= op;
        } else {
            if (border == 0) for (int i = 0; i <= a.c; i++) corr[idx(0, a.r, i)] ^= op;
            else for (int i = a.c + 1; i < L; i++) corr[idx(0, a.r, i)] ^= op;
        }
    }
    // The same optimisation problem on half the nodes.  In the reference's graphs every defect owns an ancilla it alone may
    // reach (at the cost of its border distance) and the ancillas of a border are joined to each other for free, so a perfect
    // matching is: a set M of defects sent to the border + a pairing of the rest, with |M| of fixed parity per border (free
    // graph: |M| = N mod 2 overall).  Two defects of M are as good as joined by an edge of weight b_i + b_j "through the
    // border"; an odd |M| needs one more node per border that any of its defects may take alone.  The matching then runs on
    // N + (0..2) nodes instead of 2 N + (0..2) -- an eighth of the O(n^3) work -- and has the same minimum weight.
    long long solve_layer_reduced(int layer, int parity, uint8_t *corr) const
    {
        const std::vector<Coord> &d = def[layer];
        const int nd = (int)d.size();
        std::vector<int> side(nd), bd(nd);
        int n_side[2] = {0, 0};
        for (int s = 0; s < nd; s++) {
            const int b0 = (layer == 0 ? d[s].r : d[s].c) + 1;
            if (parity < 0) side[s] = !(b0 * 2 < L);          // generate_edges, mwpm.py:106-110
            else side[s] = b0 * 2 > L;                        // generate_edges_constrained, mwpm.py:156
            bd[s] = side[s] ? L - b0 : b0;
            n_side[side[s]]++;
        }
        struct Border { int side; bool far; };
        std::vector<Border> borders;
        if (parity < 0) { if (nd & 1) borders.push_back({-1, false}); }
        else
            for (int b = 0; b < 2; b++)
                if ((n_side[b] + (parity == 1)) & 1) borders.push_back({b, n_side[b] == 0});   // far: the parity edges, mwpm.py:170-182
        Graph gr;
        gr.ndef = nd;
        gr.nodes = nd + (int)borders.size();
        std::vector<char> virt((size_t)nd * nd, 0);
        for (int i = 0; i < nd; i++)
            for (int j = i + 1; j < nd; j++) {
                int w = manhattan(d[i], d[j]);
                if ((parity < 0 || side[i] == side[j]) && bd[i] + bd[j] < w) { w = bd[i] + bd[j]; virt[(size_t)i * nd + j] = 1; }
                gr.edges.push_back({i, j, w});
            }
        for (int k = 0; k < (int)borders.size(); k++)
            for (int s = 0; s < nd; s++) {
                if (borders[k].far) gr.edges.push_back({s, nd + k, L - bd[s]});
                else if (borders[k].side < 0 || borders[k].side == side[s]) gr.edges.push_back({s, nd + k, bd[s]});
            }
        std::vector<std::pair<int, int>> pairs;
        long long w;
        if (!min_weight_perfect_matching(gr, pairs, w)) return -1;
        for (const auto &pr : pairs) {
            const int i = pr.first, j = pr.second;
            if (j >= nd) {
                const Border &b = borders[j - nd];
                eliminate_border(d[i], layer, b.far ? b.side : side[i], corr);
            } else if (virt[(size_t)i * nd + j]) {
                eliminate_border(d[i], layer, side[i], corr);
                eliminate_border(d[j], layer, side[j], corr);
            } else {
                eliminate_pair(d[i], d[j], layer, corr);
            }
        }
        return w;
    }

    bool reduced = true;   // false: the reference's own graphs, ancillas and all (tests: same weights)

    // MWPM.solve_layer (mwpm.py:319-373); parity < 0: unconstrained.  corr is XORed into; returns the matching weight
    // (-1: no perfect matching)
    long long solve_layer(int layer, int parity, uint8_t *corr) const
    {
        if (reduced) return solve_layer_reduced(layer, parity, corr);
        const std::vector<Coord> &d = def[layer];
        const Graph gr = parity < 0 ? edges_free(d, layer, L) : edges_constrained(d, layer, L, parity);
        std::vector<std::pair<int, int>> pairs;
        long long w;
        if (!min_weight_perfect_matching(gr, pairs, w)) return -1;
        const int nd = gr.ndef;
        for (const auto &pr : pairs) {
            if (pr.first < nd && pr.second >= nd) eliminate_border(d[pr.first], layer, gr.ancilla_side[pr.second - nd], corr);
            else if (pr.first < nd && pr.second < nd) eliminate_pair(d[pr.first], d[pr.second], layer, corr);
        }
        return w;
    }
};

// Planar_code.syndrom (planar_model.py:134-153): vertex defects [L-1][L] from Y/Z errors, plaquette defects [L][L-1] from
// X/Y errors; coordinates in np.nonzero order (row-major)
static void planar_defects(const uint8_t *qm, int L, std::vector<Coord> &vertex, std::vector<Coord> &plaquette)
{
    auto q = [&](int l, int r, int c) { return qm[((size_t)l * L + r) * L + c]; };
    auto yz = [&](int l, int r, int c) { const uint8_t v = q(l, r, c); return (int)(v == 2 || v == 3); };
    auto xy = [&](int l, int r, int c) { const uint8_t v = q(l, r, c); return (int)(v == 1 || v == 2); };
    vertex.clear();
    plaquette.clear();
    for (int r = 0; r < L - 1; r++)
        for (int c = 0; c < L; c++)
            if (yz(0, r + 1, c) ^ yz(0, r, c) ^ yz(1, r, c) ^ yz(1, r, (c + L - 1) % L)) vertex.push_back({r, c});
    for (int r = 0; r < L; r++)
        for (int c = 0; c < L - 1; c++)
            if (xy(1, r, c) ^ xy(1, (r + L - 1) % L, c) ^ xy(0, r, c + 1) ^ xy(0, r, c)) plaquette.push_back({r, c});
}

static int planar_class(const uint8_t *qm, int L)   // planar_model.py:379-390
{
    int x = 0, z = 0;
    for (int i = 0; i < L; i++) {
        const uint8_t a = qm[(size_t)i * L], b = qm[i];
        x += (a == 1 || a == 2);
        z += (b == 3 || b == 2);
    }
    return (x & 1) + 2 * (z & 1);
}

static int mwpm_one(int L, const uint8_t *qm, const uint8_t *vdef, const uint8_t *pdef, int mode, uint8_t *out, int32_t *weights)
{
    const size_t n = (size_t)2 * L * L;
    PlanarMwpm m;
    m.L = L;
    m.reduced = !(mode & 2);
    mode &= 1;
    if (qm) planar_defects(qm, L, m.def[0], m.def[1]);
    else {
        for (int r = 0; r < L - 1; r++)
            for (int c = 0; c < L; c++)
                if (vdef[(size_t)r * L + c]) m.def[0].push_back({r, c});
        for (int r = 0; r < L; r++)
            for (int c = 0; c < L - 1; c++)
                if (pdef[(size_t)r * (L - 1) + c]) m.def[1].push_back({r, c});
    }
    if (mode == 0) {   // MWPM.solve (mwpm.py:408-415)
        memset(out, 0, n);
        for (int layer = 0; layer < 2; layer++) {
            long long w = 0;
            if (!m.def[layer].empty()) {
                w = m.solve_layer(layer, -1, out);
                if (w < 0) return -1;
            }
            if (weights) weights[layer] = (int32_t)w;
        }
        return 0;
    }
    // MWPM.generate_classes + class_sorted_mwpm (mwpm.py:417-438, 462-475)
    std::vector<uint8_t> sol[2][2];
    for (int layer = 0; layer < 2; layer++)
        for (int parity = 0; parity < 2; parity++) {
            sol[layer][parity].assign(n, 0);
            long long w = 0;
            if (!m.def[layer].empty()) {
                w = m.solve_layer(layer, parity, sol[layer][parity].data());
                if (w < 0) return -1;
            } else if (parity == 1) {
                // Planar_code(size).apply_logical((not layer) * 2 + 1): operator 3 applies the X row AND the Z column
                // (planar_model.py:247-262), operator 1 the X row
                uint8_t *s = sol[layer][1].data();
                for (int i = 0; i < L; i++) {
                    s[i] ^= 1;                                       // X on [0, 0, i]
                    if (layer == 0) s[(size_t)i * L] ^= 3;           // Z on [0, i, 0]
                }
                for (size_t i = 0; i < n; i++) w += s[i] != 0;
            }
            if (weights) weights[layer * 2 + parity] = (int32_t)w;
        }
    std::vector<uint8_t> chain(n);
    uint8_t seen = 0;
    for (int p0 = 0; p0 < 2; p0++)
        for (int p1 = 0; p1 < 2; p1++) {
            for (size_t i = 0; i < n; i++) chain[i] = sol[0][p0][i] ^ sol[1][p1][i];
            const int cls = planar_class(chain.data(), L);
            seen |= 1u << cls;
            memcpy(out + (size_t)cls * n, chain.data(), n);
        }
    return seen == 0xF ? 0 : -2;
}

}  // namespace
}  // namespace qecmc

using namespace qecmc;

extern "C" int qecmc_mwpm_planar(int32_t L, int64_t S, const uint8_t *qm, const uint8_t *vertex_defects,
                                 const uint8_t *plaquette_defects, int32_t mode, uint8_t *out, int32_t *weights, int32_t threads)
{
    if (L < 2 || L > 32 || S < 0 || !out || mode < 0 || mode > 3) return set_err(QECMC_ERR_ARG, "bad arguments");
    if (!qm && (!vertex_defects || !plaquette_defects))
        return set_err(QECMC_ERR_ARG, "either qm or both defect arrays must be given");
    const size_t n = (size_t)2 * L * L, nout = (mode & 1) ? 4 * n : n, nw = (mode & 1) ? 4 : 2;
    const size_t nv = (size_t)(L - 1) * L;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((int64_t)nt > S) nt = (int)std::max<int64_t>(S, 1);
    std::atomic<int64_t> next(0);
    std::atomic<int> bad(0);
    auto work = [&] {
        for (;;) {
            const int64_t s = next.fetch_add(1);
            if (s >= S) break;
            const int rc = mwpm_one(L, qm ? qm + s * n : nullptr, qm ? nullptr : vertex_defects + s * nv,
                                    qm ? nullptr : plaquette_defects + s * nv, mode, out + s * nout, weights ? weights + s * nw : nullptr);
            if (rc) bad.store(rc);
        }
    };
    if (nt == 1) work();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; t++) pool.emplace_back(work);
        for (auto &t : pool) t.join();
    }
    if (bad.load() == -1) return set_err(QECMC_ERR_ARG, "a defect graph has no perfect matching (odd defects on a closed border?)");
    if (bad.load() == -2) return set_err(QECMC_ERR_ARG, "the four class-constrained matchings did not land in four classes");
    return 0;
}
