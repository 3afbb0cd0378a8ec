// Driver of tests/test_sanitizers.py: read files in the reference's stdin format (assemble.py:40-71), unpaired and
// paired, parsed by ga_parse_reads with one thread (the serial scan) and with six (byte ranges cut at line breaks);
// both must give the same symbols, lengths, read count, pairing and distance.  Built with -fsanitize=thread /
// address,undefined.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "ga_b200.h"

struct Parsed {
    std::vector<uint8_t> symbols;
    std::vector<int32_t> lens;
    uint64_t n_reads = 0, n_symbols = 0;
    int paired = 0, rc = 0;
    int64_t distance = 0;
    bool operator==(const Parsed& o) const {
        if (rc != GA_OK || o.rc != GA_OK) return rc == o.rc;       // a rejected input: the same verdict, outputs unspecified
        return rc == o.rc && n_reads == o.n_reads && n_symbols == o.n_symbols && paired == o.paired &&
               distance == o.distance && lens == o.lens &&
               std::equal(symbols.begin(), symbols.begin() + (long)n_symbols, o.symbols.begin());
    }
};

static Parsed parse(const std::string& text, const char* threads) {
    setenv("GA_PARSE_THREADS", threads, 1);
    setenv("GA_PARSE_GRAIN", "65536", 1);
    Parsed p;
    const uint8_t* t = (const uint8_t*)text.data();
    p.rc = ga_parse_reads(t, text.size(), nullptr, nullptr, 0, &p.n_reads, &p.paired, &p.distance, &p.n_symbols);
    if (p.rc != GA_OK) return p;
    p.symbols.assign(text.size() + 1, 0);
    p.lens.assign(p.n_reads * (p.paired ? 2 : 1), 0);
    p.rc = ga_parse_reads(t, text.size(), p.symbols.data(), p.lens.data(), p.lens.size(), &p.n_reads, &p.paired,
                          &p.distance, &p.n_symbols);
    return p;
}

int main() {
    std::mt19937_64 rng(3);
    int bad = 0;
    for (int paired = 0; paired < 2; ++paired)
        for (int shape = 0; shape < 3; ++shape) {
            const int n = 60000 + 7000 * shape;
            std::string text = std::to_string(shape == 2 ? n + 500 : n) + "\n";     // shape 2: fewer lines than announced
            for (int r = 0; r < n; ++r) {
                const int len = shape == 1 ? (int)(rng() % 120) : 100;              // shape 1: ragged, some empty
                for (int m = 0; m < (paired ? 2 : 1); ++m) {
                    for (int i = 0; i < len; ++i) text.push_back("ACGT"[rng() & 3]);
                    if (paired) text.push_back('|');
                }
                if (paired) text += "125";
                if (shape == 1 && r % 97 == 0) text += "  ";                        // trailing blanks are stripped
                text.push_back('\n');
            }
            if (paired && shape == 2) text += "ACGT|AC\n";                          // a pair line with two fields: rejected
            const Parsed a = parse(text, "1"), b = parse(text, "6");
            const bool same = a == b;
            printf("paired %d shape %d: rc %d reads %llu symbols %llu %s\n", paired, shape, a.rc,
                   (unsigned long long)a.n_reads, (unsigned long long)a.n_symbols, same ? "same" : "DIFFERENT");
            bad += !same;
        }
    return bad ? 1 : 0;
}
