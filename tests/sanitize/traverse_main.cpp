// Driver of tests/test_sanitizers.py: random graphs of the shape the contig traversal sees (a few branching hubs joined by
// long chains of 1-in-1-out nodes with shuffled ids), walked by the serial sweep and by the piecewise routine on 6 threads;
// both must give the same contigs and the same left-over edge counts.  Built with -fsanitize=thread / address,undefined.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <algorithm>
extern "C" int ga_traverse_contigs(const int32_t*, const int32_t*, const int32_t*, const uint8_t*, const uint8_t*, int64_t,
                                   int64_t, int, uint8_t**, uint64_t**, uint64_t*, int32_t*);
extern "C" int ga_traverse_last_route(void);
struct Out { std::vector<uint8_t> text; std::vector<uint64_t> offs; std::vector<int32_t> left; int route; };
static Out run(const std::vector<int32_t>& rowptr, const std::vector<int32_t>& col, const std::vector<int32_t>& indeg,
               const std::vector<uint8_t>& br, const std::vector<uint8_t>& last, int64_t attr, int paired) {
    uint8_t* t; uint64_t* o; uint64_t n; Out r; r.left.assign(indeg.size(), -7);
    int rc = ga_traverse_contigs(rowptr.data(), col.data(), indeg.data(), br.data(), last.data(), (int64_t)indeg.size(), attr,
                                 paired, &t, &o, &n, r.left.data());
    if (rc) { printf("rc %d\n", rc); exit(1); }
    r.offs.assign(o, o + n + 1); r.text.assign(t, t + o[n]); r.route = ga_traverse_last_route(); free(t); free(o); return r;
}
int main() {
    std::mt19937_64 rng(5);
    for (int round = 0; round < 6; ++round) {
        const int n = 200000 + round * 50000, hubs = round == 5 ? 2 : 50;
        std::vector<int32_t> ids(n); for (int i = 0; i < n; ++i) ids[i] = i;
        std::shuffle(ids.begin(), ids.end(), rng);
        std::vector<std::vector<int32_t>> succ(n);
        std::vector<int32_t> hub(ids.end() - hubs, ids.end()); ids.resize(n - hubs);
        while (!ids.empty()) {
            size_t len = std::min<size_t>(ids.size(), rng() % 5000);
            int32_t cur = hub[rng() % hubs];
            for (size_t i = 0; i < len; ++i) { int32_t nx = ids.back(); ids.pop_back(); succ[cur].push_back(nx); cur = nx; }
            if (rng() % 5) succ[cur].push_back(hub[rng() % hubs]);
        }
        std::vector<int32_t> rowptr(n + 1, 0), col, indeg(n, 0);
        for (int i = 0; i < n; ++i) { rowptr[i + 1] = rowptr[i] + (int32_t)succ[i].size(); for (int32_t j : succ[i]) { col.push_back(j); ++indeg[j]; } }
        std::vector<uint8_t> br(n), last(n);
        for (int i = 0; i < n; ++i) { br[i] = succ[i].size() > 1 || indeg[i] > 1; last[i] = 'A' + (uint8_t)(rng() % 26); }
        for (int paired = 0; paired < 2; ++paired) {
            setenv("GA_TRAVERSE_THREADS", "0", 1);
            Out s = run(rowptr, col, indeg, br, last, (int64_t)col.size(), paired);
            setenv("GA_TRAVERSE_THREADS", "6", 1);
            Out p = run(rowptr, col, indeg, br, last, (int64_t)col.size(), paired);
            bool same = s.text == p.text && s.offs == p.offs && s.left == p.left;
            printf("round %d paired %d: n=%d contigs=%zu route=%d %s\n", round, paired, n, s.offs.size() - 1, p.route, same ? "same" : "DIFFERENT");
            if (!same || p.route != 1) return 1;
        }
    }
    return 0;
}
