// partition.cu -- host restatement of the reference's algebraic partitioner.
//
// The north-star keeps aggregation on the host; this file is host-only C++ (OpenMP), no kernels.  It
// exists so the non-Rust harness can build hierarchies from *algebraic* aggregates the way
// `AggregationConfig::build` (src/interpolation/mod.rs:129-156) and `BlockSmootherConfig::build`
// (src/preconditioners/block_smoothers.rs:56-69) do, through `PartitionerConfig::build_partition`
// (src/partitioners/mod.rs:273-329):
//   AdjacencyList::new_ls_strength_graph   partitioners/mod.rs:337-393   (+ extract_local_subgraph :695-718)
//   AdjacencyList::pairwise_merge / map_indices / merge_pair              partitioners/mod.rs:439-463, 505-586
//   Partitioner::new                        partitioners/modularity.rs:28-137
//   initialize_partition / greedy_matching / generate_modularity_triplets modularity.rs:179-192, 305-383
//   improve_partition / delta_q / size_cost                              modularity.rs:385-510
//
// The reference is not deterministic here: it sorts floats with `sort_unstable_by` /
// `par_sort_unstable_by` where ties are everywhere on uniform grids, and takes `max_by` over a
// `HashSet` iteration (random SipHash keys).  Every such choice is made deterministic below and is
// marked TIE-BREAK; the Python oracle (partitioner.py in the oracle directory) uses the same rules, so aggregates
// are comparable bit for bit between the two restatements.  They are NOT comparable with a given
// run of the Rust crate beyond "same algorithm, some admissible tie order".
//   TIE-BREAK 1  strength neighbourhood sorted by distance: stable (ties keep ascending neighbour id)
//   TIE-BREAK 2  map_indices sort by mapped id: stable (duplicates keep their previous order)
//   TIE-BREAK 3  merge candidates sorted ascending by weight: stable over the (i ascending, list
//                order) generation sequence; popped from the back => among equal weights the
//                last generated triplet is matched first
//   TIE-BREAK 4  best destination aggregate of a node: candidates visited in ascending aggregate id,
//                `max_by` semantics (the last maximal element wins => the largest id among ties)
//   TIE-BREAK 5  swaps sorted descending by gain: stable over ascending node id
// Reference bug F10c (`pairwise_merge_rowsums` writes the unmatched row sums at `new_idx` instead of
// `new_idx + pairs.len()`, modularity.rs:299-301) is fixed: the evident intent is implemented.
//   TIE-BREAK 6  AdjacencyList::aggregate (block_size > 1, mod.rs:293-300, 465-496, 588-656): the k-way merge
//                pops a BinaryHeap keyed by neighbour id only; equal ids are summed here in ascending
//                (member of the aggregate, position in its list) order
#include <cmath>
#include <parallel/algorithm>  // __gnu_parallel::stable_sort: a stable sort has one result, so threads do not change it
#include <queue>

#include "common.cuh"

struct famg_graph {
    std::vector<std::vector<std::pair<int64_t, double>>> nodes;  // AdjacencyList.nodes
};

namespace famg {

using Edge = std::pair<int64_t, double>;
using Adj = std::vector<std::vector<Edge>>;

static bool by_id(const Edge &a, const Edge &b) { return a.first < b.first; }
static bool by_w(const Edge &a, const Edge &b) { return a.second < b.second; }

// partitioners/mod.rs:337-393
static famg_status ls_strength_graph(int64_t n, const uint64_t *row_ptr, const uint64_t *col, const double *nn, int64_t ldn, int64_t k,
                                     const double *weights, int64_t max_depth, Adj &nodes) {
    const double theta = 0.5, eps_norm = 1e-30;
    Adj upper((size_t)n);
    std::vector<double> norms((size_t)n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int64_t c = 0; c < k; ++c) s += nn[c * ldn + i] * weights[c] * nn[c * ldn + i];
        norms[(size_t)i] = std::max(s, eps_norm);
    }
#pragma omp parallel
    {
        // extract_local_subgraph (:695-718): BFS to max_depth over the pattern; `stamp` is the visited set
        std::vector<int64_t> stamp((size_t)n, -1), frontier, next, found;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            frontier.assign(1, i);
            found.clear();
            stamp[(size_t)i] = i;
            for (int64_t depth = 0; depth < max_depth && !frontier.empty(); ++depth) {
                next.clear();
                for (int64_t j : frontier)
                    for (uint64_t q = row_ptr[j]; q < row_ptr[j + 1]; ++q) {
                        const int64_t nb = (int64_t)col[q];
                        if (stamp[(size_t)nb] != i) {
                            stamp[(size_t)nb] = i;
                            next.push_back(nb);
                            if (nb > i) found.push_back(nb);
                        }
                    }
                frontier.swap(next);
            }
            std::sort(found.begin(), found.end());  // BTreeSet iteration order
            auto &out = upper[(size_t)i];
            out.reserve(found.size());
            for (int64_t j : found) {
                double dot = 0.0;
                for (int64_t c = 0; c < k; ++c) dot += nn[c * ldn + i] * weights[c] * nn[c * ldn + j];
                const double rho2 = (dot * dot) / (norms[(size_t)i] * norms[(size_t)j]);
                out.emplace_back(j, 2.0 * std::sqrt(std::max(1.0 - rho2, 0.0)));
            }
        }
    }
    // nodes[i] = { (i', d) : i' < i, i in N(i') } ++ { (j, d) : j > i, j in N(i) }: the push order of :357-358
    std::vector<int64_t> lower_count((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i)
        for (const Edge &e : upper[(size_t)i]) ++lower_count[(size_t)e.first];
    nodes.assign((size_t)n, {});
    for (int64_t i = 0; i < n; ++i) nodes[(size_t)i].reserve((size_t)lower_count[(size_t)i] + upper[(size_t)i].size());
    for (int64_t i = 0; i < n; ++i)
        for (const Edge &e : upper[(size_t)i]) nodes[(size_t)e.first].emplace_back(i, e.second);
    bool disconnected = false;
    const double eps = 1e-12, alpha = 4.0;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; ++i) {
        auto &nb = nodes[(size_t)i];
        nb.insert(nb.end(), upper[(size_t)i].begin(), upper[(size_t)i].end());
        std::vector<Edge>().swap(upper[(size_t)i]);
        if (nb.empty()) { disconnected = true; continue; }
        std::stable_sort(nb.begin(), nb.end(), by_w);  // TIE-BREAK 1
        const size_t keep = std::max<size_t>((size_t)std::floor((double)nb.size() * theta), 1);
        nb.resize(keep);
        nb.shrink_to_fit();
        const double d_min = nb.front().second, d_max = nb.back().second;
        if (std::fabs(d_max - d_min) < eps) {
            for (Edge &e : nb) e.second = 1.0;
        } else {
            for (Edge &e : nb) e.second = std::pow((d_max - e.second) / (d_max - d_min + eps), alpha);
        }
        std::sort(nb.begin(), nb.end(), by_id);  // ids are unique
    }
    if (disconnected) FAMG_FAIL(FAMG_ERR_INVALID, "strength graph: a node has no neighbours (graph is disconnected)");
    return FAMG_OK;
}

// partitioners/mod.rs:518-586, literal: two-way merge; equal heads are combined first, then the result is
// pushed or accumulated into the last merged entry (lists may hold duplicate ids after map_indices)
static void merge_pair(const std::vector<Edge> &a, const std::vector<Edge> &b, std::vector<Edge> &merged) {
    merged.clear();
    merged.reserve(a.size() + b.size());
    auto add = [&](Edge e) {
        if (!merged.empty() && merged.back().first == e.first) merged.back().second += e.second;
        else merged.push_back(e);
    };
    size_t ia = 0, ib = 0;
    while (ia < a.size() && ib < b.size()) {
        if (a[ia].first == b[ib].first) { add(Edge(a[ia].first, a[ia].second + b[ib].second)); ++ia; ++ib; }
        else if (a[ia].first < b[ib].first) add(a[ia++]);
        else add(b[ib++]);
    }
    while (ia < a.size()) add(a[ia++]);
    while (ib < b.size()) add(b[ib++]);
}

struct Partitioner {
    double cf = 8.0, agg_pen = 1.0;
    int64_t max_iters = 100;
    const Adj *base = nullptr;  // base_strength (fine graph, never modified without rebase)
    Adj strength;               // current (merged) graph
    std::vector<double> row_sums;
    double inverse_total = 0.0;
    std::vector<int64_t> node_to_agg;
    std::vector<std::vector<int64_t>> agg_to_node;  // ascending (BTreeSet)
    std::vector<int64_t> agg_sizes;                  // node weights are all 1 (modularity.rs:81)

    int64_t nnodes() const { return (int64_t)node_to_agg.size(); }
    int64_t naggs() const { return (int64_t)agg_to_node.size(); }
    double part_cf() const { return (double)nnodes() / (double)naggs(); }

    // modularity.rs:28-137 with starting_partition = None
    famg_status init(const Adj &g) {
        base = &g;
        strength = g;
        const int64_t n = (int64_t)g.size();
        row_sums.assign((size_t)n, 0.0);
        for (int64_t i = 0; i < n; ++i) {
            double s = 0.0;
            for (const Edge &e : g[(size_t)i]) {
                if (e.first == i) FAMG_FAIL(FAMG_ERR_INVALID, "strength graph has a self loop at node %lld", (long long)i);
                s += e.second;
            }
            row_sums[(size_t)i] = s < 0.0 ? 0.0 : s;
        }
        double total = 0.0;
        for (double s : row_sums) total += s;
        inverse_total = 1.0 / total;
        node_to_agg.resize((size_t)n);
        agg_to_node.assign((size_t)n, {});
        for (int64_t i = 0; i < n; ++i) { node_to_agg[(size_t)i] = i; agg_to_node[(size_t)i].assign(1, i); }
        agg_sizes.assign((size_t)n, 1);
        return FAMG_OK;
    }

    struct Triplet { int64_t i, j; double w; };

    // modularity.rs:305-337 (same order: i ascending, list order; generated by all cores)
    void modularity_triplets(std::vector<Triplet> &out) const {
        const int64_t nv = (int64_t)strength.size();
        std::vector<int64_t> start((size_t)nv + 1, 0);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nv; ++i) {
            int64_t c = 0;
            for (const Edge &e : strength[(size_t)i]) c += i > e.first;
            start[(size_t)i + 1] = c;
        }
        for (int64_t i = 0; i < nv; ++i) start[(size_t)i + 1] += start[(size_t)i];
        out.resize((size_t)start[(size_t)nv]);
#pragma omp parallel for schedule(dynamic, 1024)
        for (int64_t i = 0; i < nv; ++i) {
            int64_t t = start[(size_t)i];
            for (const Edge &e : strength[(size_t)i]) {
                if (!(i > e.first)) continue;
                const int64_t j = e.first;
                const double expected = inverse_total * row_sums[(size_t)i] * row_sums[(size_t)j];
                double w = e.second - expected;
                const double new_weight = (double)(agg_sizes[(size_t)i] + agg_sizes[(size_t)j]);
                const double square_diff = std::pow(new_weight - cf, 2.0);
                if (new_weight > cf) w -= agg_pen * square_diff;
                else w += agg_pen * square_diff;
                out[(size_t)t++] = {i, j, w};
            }
        }
    }

    // modularity.rs:339-383
    void greedy_matching(double step_cf, std::vector<std::pair<int64_t, int64_t>> &pairs, std::vector<int64_t> &unmatched) const {
        pairs.clear();
        unmatched.clear();
        const int64_t vertex_count = (int64_t)row_sums.size();
        const double t = std::ceil((double)vertex_count - (double)nnodes() / step_cf);
        const uint64_t target_matches = (t > 0.0 ? (uint64_t)t : 0) + 1;  // `as usize` saturates at 0
        std::vector<Triplet> wants;
        modularity_triplets(wants);
        if (wants.empty()) return;
        __gnu_parallel::stable_sort(wants.begin(), wants.end(), [](const Triplet &a, const Triplet &b) { return a.w < b.w; });  // TIE-BREAK 3
        std::vector<char> alive((size_t)vertex_count, 1);
        while (!wants.empty()) {
            const Triplet tr = wants.back();
            wants.pop_back();
            if (alive[(size_t)tr.i] && alive[(size_t)tr.j]) {
                alive[(size_t)tr.i] = alive[(size_t)tr.j] = 0;
                pairs.emplace_back(tr.i, tr.j);
            }
            if (pairs.size() > target_matches) break;
        }
        for (int64_t i = 0; i < vertex_count; ++i)
            if (alive[(size_t)i]) unmatched.push_back(i);
    }

    // partitioners/mod.rs:439-463 (graph), :110-129 (partition), modularity.rs:291-303 (row sums, F10c fixed)
    void pairwise_merge(const std::vector<std::pair<int64_t, int64_t>> &pairs, const std::vector<int64_t> &unmatched) {
        const int64_t old_n = (int64_t)strength.size(), pairs_n = (int64_t)pairs.size();
        const int64_t new_n = pairs_n + (int64_t)unmatched.size();
        std::vector<int64_t> agg_ids((size_t)old_n, 0);
        for (int64_t a = 0; a < pairs_n; ++a) { agg_ids[(size_t)pairs[(size_t)a].first] = a; agg_ids[(size_t)pairs[(size_t)a].second] = a; }
        for (int64_t a = 0; a < (int64_t)unmatched.size(); ++a) agg_ids[(size_t)unmatched[(size_t)a]] = a + pairs_n;
#pragma omp parallel for schedule(dynamic, 1024)
        for (int64_t i = 0; i < old_n; ++i) {  // map_indices (:505-512)
            auto &nb = strength[(size_t)i];
            for (Edge &e : nb) e.first = agg_ids[(size_t)e.first];
            std::stable_sort(nb.begin(), nb.end(), by_id);  // TIE-BREAK 2
        }
        Adj merged((size_t)new_n);
#pragma omp parallel for schedule(dynamic, 1024)
        for (int64_t a = 0; a < new_n; ++a) {
            if (a < pairs_n) merge_pair(strength[(size_t)pairs[(size_t)a].first], strength[(size_t)pairs[(size_t)a].second], merged[(size_t)a]);
            else merged[(size_t)a] = strength[(size_t)unmatched[(size_t)(a - pairs_n)]];
        }
        strength.swap(merged);

        std::vector<std::vector<int64_t>> new_aggs((size_t)new_n);
        std::vector<double> new_row_sums((size_t)new_n, 0.0);
        for (int64_t a = 0; a < new_n; ++a) {
            if (a < pairs_n) {
                const auto &x = agg_to_node[(size_t)pairs[(size_t)a].first], &y = agg_to_node[(size_t)pairs[(size_t)a].second];
                new_aggs[(size_t)a].resize(x.size() + y.size());
                std::merge(x.begin(), x.end(), y.begin(), y.end(), new_aggs[(size_t)a].begin());
                new_row_sums[(size_t)a] = row_sums[(size_t)pairs[(size_t)a].first] + row_sums[(size_t)pairs[(size_t)a].second];
            } else {
                const int64_t old = unmatched[(size_t)(a - pairs_n)];
                new_aggs[(size_t)a] = agg_to_node[(size_t)old];
                new_row_sums[(size_t)a] = row_sums[(size_t)old];  // F10c: the reference drops the `+ pairs.len()` offset
            }
        }
        agg_to_node.swap(new_aggs);
        row_sums.swap(new_row_sums);
        agg_sizes.resize((size_t)new_n);
        for (int64_t a = 0; a < new_n; ++a) {
            for (int64_t node : agg_to_node[(size_t)a]) node_to_agg[(size_t)node] = a;
            agg_sizes[(size_t)a] = (int64_t)agg_to_node[(size_t)a].size();  // update_agg_sizes (:194-207), unit weights
        }
    }

    // modularity.rs:179-192
    void initialize_partition() {
        std::vector<std::pair<int64_t, int64_t>> pairs;
        std::vector<int64_t> unmatched;
        while (part_cf() < cf) {
            greedy_matching(cf, pairs, unmatched);
            if (pairs.empty()) break;  // "no more matches are possible"
            pairwise_merge(pairs, unmatched);
        }
    }

    // modularity.rs:385-389
    double size_cost(int64_t size) const {
        const double relative_diff = std::fabs((double)size - cf) / cf;
        return std::pow(4.0 * relative_diff, 4.0) * agg_pen;
    }

    // modularity.rs:391-435 with source_agg = Some(..)
    double delta_q(int64_t node_i, int64_t source_agg, int64_t dest_agg) const {
        double in_degree = 0.0, out_degree = 0.0;
        for (const Edge &e : (*base)[(size_t)node_i]) {
            const int64_t agg_j = node_to_agg[(size_t)e.first];
            if (agg_j == source_agg) in_degree += e.second;
            else if (agg_j == dest_agg) out_degree += e.second;
        }
        const int64_t old_dst = agg_sizes[(size_t)dest_agg], new_dst = old_dst + 1;
        const int64_t old_src = agg_sizes[(size_t)source_agg], new_src = old_src - 1;
        const double old_size_cost = size_cost(old_dst) + size_cost(old_src);
        const double new_size_cost = size_cost(new_dst) + size_cost(new_src);
        const double delta_degree = out_degree - in_degree;
        const double delta_size = old_size_cost - new_size_cost;
        return delta_degree + agg_pen * delta_size;
    }

    // modularity.rs:437-510.  The reference recomputes every node's best move in every pass; the result of
    // that computation for node i depends only on the aggregates of i and of its neighbours and on the sizes
    // of those aggregates, so after a pass only the members of aggregates whose size changed, and the nodes
    // that list such a member as a neighbour, are recomputed -- the same swaps, pass for pass (the candidate
    // count falls from ~n/4 to a handful within ~50 passes on grid problems).
    void improve_partition() {
        struct Swap { int64_t node, agg; double gain; };
        const int64_t n = nnodes();
        // reverse adjacency of the (directed) base graph: who lists node j as a neighbour
        std::vector<int64_t> in_ptr((size_t)n + 1, 0), in_idx;
        for (int64_t i = 0; i < n; ++i)
            for (const Edge &e : (*base)[(size_t)i]) ++in_ptr[(size_t)e.first + 1];
        for (int64_t i = 0; i < n; ++i) in_ptr[(size_t)i + 1] += in_ptr[(size_t)i];
        in_idx.resize((size_t)in_ptr[(size_t)n]);
        {
            std::vector<int64_t> fill(in_ptr.begin(), in_ptr.end() - 1);
            for (int64_t i = 0; i < n; ++i)
                for (const Edge &e : (*base)[(size_t)i]) in_idx[(size_t)fill[(size_t)e.first]++] = i;
        }
        std::vector<Swap> best((size_t)n);
        std::vector<Swap> swaps;
        std::vector<int64_t> todo((size_t)n), changed_aggs;
        for (int64_t i = 0; i < n; ++i) todo[(size_t)i] = i;
        std::vector<char> dirty((size_t)n, 0);
        for (int64_t pass = 0; pass < max_iters; ++pass) {
            const int64_t ntodo = (int64_t)todo.size();
#pragma omp parallel
            {
                std::vector<int64_t> cand;
#pragma omp for schedule(dynamic, 256)
                for (int64_t t = 0; t < ntodo; ++t) {
                    const int64_t i = todo[(size_t)t];
                    best[(size_t)i] = {i, -1, 0.0};
                    const int64_t agg_i = node_to_agg[(size_t)i];
                    if (agg_sizes[(size_t)agg_i] == 1) continue;  // sole member cannot leave
                    cand.clear();
                    for (const Edge &e : (*base)[(size_t)i]) {
                        const int64_t agg_j = node_to_agg[(size_t)e.first];
                        if (agg_j != agg_i) cand.push_back(agg_j);
                    }
                    std::sort(cand.begin(), cand.end());
                    cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
                    for (int64_t agg_j : cand) {  // TIE-BREAK 4: ascending id, last maximal wins
                        const double dq = delta_q(i, agg_i, agg_j);
                        if (dq > 0.0 && (best[(size_t)i].agg < 0 || dq >= best[(size_t)i].gain)) best[(size_t)i] = {i, agg_j, dq};
                    }
                }
            }
            swaps.clear();
            for (int64_t i = 0; i < n; ++i)
                if (best[(size_t)i].agg >= 0) swaps.push_back(best[(size_t)i]);
            if (swaps.empty()) break;
            static const bool trace = getenv("FAMG_PARTITION_TRACE") != nullptr;
            if (trace) fprintf(stderr, "improve_partition pass %lld: %zu candidate swaps, %lld nodes re-evaluated\n", (long long)pass, swaps.size(), (long long)ntodo);
            __gnu_parallel::stable_sort(swaps.begin(), swaps.end(), [](const Swap &a, const Swap &b) { return a.gain > b.gain; });  // TIE-BREAK 5
            std::vector<char> alive_nodes((size_t)n, 1), alive_aggs((size_t)naggs(), 1);
            changed_aggs.clear();
            for (const Swap &s : swaps) {
                const int64_t old_agg = node_to_agg[(size_t)s.node];
                if (!(alive_nodes[(size_t)s.node] && alive_aggs[(size_t)s.agg] && alive_aggs[(size_t)old_agg])) continue;
                node_to_agg[(size_t)s.node] = s.agg;
                agg_sizes[(size_t)old_agg] -= 1;
                agg_sizes[(size_t)s.agg] += 1;
                auto &src = agg_to_node[(size_t)old_agg];
                src.erase(std::lower_bound(src.begin(), src.end(), s.node));
                auto &dst = agg_to_node[(size_t)s.agg];
                dst.insert(std::lower_bound(dst.begin(), dst.end(), s.node), s.node);
                changed_aggs.push_back(old_agg);
                changed_aggs.push_back(s.agg);
                alive_aggs[(size_t)s.agg] = alive_aggs[(size_t)old_agg] = 0;
                alive_nodes[(size_t)s.node] = 0;
                for (const Edge &e : (*base)[(size_t)s.node]) {
                    alive_nodes[(size_t)e.first] = 0;
                    alive_aggs[(size_t)node_to_agg[(size_t)e.first]] = 0;
                }
            }
            // nodes whose best move may have changed: members of the resized aggregates (the moved nodes are
            // among them) and everyone who lists such a member as a neighbour
            todo.clear();
            auto mark = [&](int64_t i) { if (!dirty[(size_t)i]) { dirty[(size_t)i] = 1; todo.push_back(i); } };
            for (int64_t a : changed_aggs)
                for (int64_t m : agg_to_node[(size_t)a]) {
                    mark(m);
                    for (int64_t q = in_ptr[(size_t)m]; q < in_ptr[(size_t)m + 1]; ++q) mark(in_idx[(size_t)q]);
                }
            std::sort(todo.begin(), todo.end());
            for (int64_t i : todo) dirty[(size_t)i] = 0;
        }
    }
};

}  // namespace famg

using namespace famg;

extern "C" {

famg_status famg_strength_graph_create(int64_t n, const uint64_t *row_ptr, const uint64_t *col_idx, const double *near_null, int64_t ldn,
                                       int64_t k, const double *weights, int64_t max_depth, famg_graph **out) {
    if (!row_ptr || (!col_idx && n > 0 && row_ptr[n] > 0) || !near_null || !weights || !out || n < 0 || k < 1 || ldn < n || max_depth < 0)
        FAMG_FAIL(FAMG_ERR_INVALID, "strength graph: bad argument");
    for (int64_t i = 0; i < n; ++i)
        for (uint64_t q = row_ptr[i]; q < row_ptr[i + 1]; ++q)
            if (col_idx[q] >= (uint64_t)n) FAMG_FAIL(FAMG_ERR_INVALID, "strength graph: column index out of range (square matrix expected)");
    famg_graph *g = new (std::nothrow) famg_graph;
    if (!g) FAMG_FAIL(FAMG_ERR_ALLOC, "out of host memory");
    famg_status st;
    try {
        st = ls_strength_graph(n, row_ptr, col_idx, near_null, ldn, k, weights, max_depth, g->nodes);
    } catch (const std::bad_alloc &) {
        set_error("strength graph: out of host memory");
        st = FAMG_ERR_ALLOC;
    }
    if (st != FAMG_OK) { delete g; return st; }
    *out = g;
    return FAMG_OK;
}

famg_status famg_graph_create(int64_t n, const uint64_t *row_ptr, const uint64_t *col_idx, const double *w, famg_graph **out) {
    if (!row_ptr || !out || n < 0) FAMG_FAIL(FAMG_ERR_INVALID, "graph: bad argument");
    famg_graph *g = new (std::nothrow) famg_graph;
    if (!g) FAMG_FAIL(FAMG_ERR_ALLOC, "out of host memory");
    g->nodes.assign((size_t)n, {});
    for (int64_t i = 0; i < n; ++i)
        for (uint64_t q = row_ptr[i]; q < row_ptr[i + 1]; ++q) {
            if (col_idx[q] >= (uint64_t)n) { delete g; FAMG_FAIL(FAMG_ERR_INVALID, "graph: neighbour index out of range"); }
            g->nodes[(size_t)i].emplace_back((int64_t)col_idx[q], w[q]);
        }
    *out = g;
    return FAMG_OK;
}

// strength.aggregate(&block_reduce); strength.filter_diag()  (partitioners/mod.rs:293-300): dofs -> nodes of
// `block_size` consecutive dofs.  aggregate (:465-496): neighbour ids mapped to node ids, the lists of a
// node's dofs merged with equal ids summed (merge_agg, :588-656), every weight divided by the largest
// merged weight of the whole graph (self loops included, as the reference does); then self loops dropped.
famg_status famg_graph_block_reduce(famg_graph *g, int64_t block_size) {
    if (!g || block_size < 1) FAMG_FAIL(FAMG_ERR_INVALID, "graph: bad argument");
    const int64_t n = (int64_t)g->nodes.size();
    if (n % block_size != 0) FAMG_FAIL(FAMG_ERR_INVALID, "graph: size is not a multiple of the block size");
    if (block_size == 1) return FAMG_OK;
    const int64_t nb = n / block_size;
    std::vector<std::vector<std::pair<int64_t, double>>> merged((size_t)nb);
    std::vector<double> local_max((size_t)nb, 0.0);
    bool empty = false;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t b = 0; b < nb; ++b) {
        std::vector<std::pair<int64_t, double>> all;
        for (int64_t o = 0; o < block_size; ++o) {
            const auto &nbrs = g->nodes[(size_t)(b * block_size + o)];
            if (nbrs.empty()) empty = true;
            for (const auto &e : nbrs) all.emplace_back(e.first / block_size, e.second);
        }
        std::stable_sort(all.begin(), all.end(), [](const std::pair<int64_t, double> &x, const std::pair<int64_t, double> &y) { return x.first < y.first; });  // TIE-BREAK 2 + 6
        auto &out = merged[(size_t)b];
        for (const auto &e : all) {
            if (!out.empty() && out.back().first == e.first) out.back().second += e.second;
            else out.push_back(e);
        }
        double mx = out.empty() ? 0.0 : out[0].second;
        for (const auto &e : out) mx = std::max(mx, e.second);
        local_max[(size_t)b] = mx;
    }
    if (empty) FAMG_FAIL(FAMG_ERR_INVALID, "empty neighborhood means graph is disconnected");
    double mx = local_max.empty() ? 1.0 : local_max[0];
    for (double v : local_max) mx = std::max(mx, v);
    for (int64_t b = 0; b < nb; ++b) {
        auto &row = merged[(size_t)b];
        for (auto &e : row) e.second /= mx;
        row.erase(std::remove_if(row.begin(), row.end(), [b](const std::pair<int64_t, double> &e) { return e.first == b; }), row.end());  // filter_diag (:498-502)
    }
    g->nodes.swap(merged);
    return FAMG_OK;
}

famg_status famg_graph_dims(const famg_graph *g, int64_t *n, int64_t *nnz) {
    if (!g || !n || !nnz) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *n = (int64_t)g->nodes.size();
    int64_t s = 0;
    for (const auto &nb : g->nodes) s += (int64_t)nb.size();
    *nnz = s;
    return FAMG_OK;
}

famg_status famg_graph_download(const famg_graph *g, uint64_t *row_ptr, uint64_t *col_idx, double *w) {
    if (!g || !row_ptr) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    uint64_t q = 0;
    row_ptr[0] = 0;
    for (size_t i = 0; i < g->nodes.size(); ++i) {
        for (const auto &e : g->nodes[i]) { col_idx[q] = (uint64_t)e.first; w[q] = e.second; ++q; }
        row_ptr[i + 1] = q;
    }
    return FAMG_OK;
}

famg_status famg_graph_destroy(famg_graph *g) {
    delete g;
    return FAMG_OK;
}

famg_status famg_partition_modularity(const famg_graph *g, double coarsening_factor, double agg_size_penalty, int64_t max_improvement_iters,
                                      uint64_t *node_to_agg, int64_t *naggs) {
    if (!g || !node_to_agg || !naggs || max_improvement_iters < 0) FAMG_FAIL(FAMG_ERR_INVALID, "partition: bad argument");
    if (g->nodes.empty()) { *naggs = 0; return FAMG_OK; }
    try {
        Partitioner p;
        p.cf = coarsening_factor;
        p.agg_pen = agg_size_penalty;
        p.max_iters = max_improvement_iters;
        FAMG_TRY(p.init(g->nodes));
        p.initialize_partition();
        p.improve_partition();
        for (size_t i = 0; i < p.node_to_agg.size(); ++i) node_to_agg[i] = (uint64_t)p.node_to_agg[i];
        *naggs = p.naggs();
    } catch (const std::bad_alloc &) {
        FAMG_FAIL(FAMG_ERR_ALLOC, "partition: out of host memory");
    }
    return FAMG_OK;
}

}  // extern "C"
