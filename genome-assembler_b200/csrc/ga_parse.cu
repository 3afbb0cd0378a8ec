// ga_parse.cu -- raw ingest on the host (no device code; up to 16 host threads).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ga_common.cuh"

// ------------------------------------------------------------------------------------------------
// Raw ingest: the reference's IOHandler.read_input (assemble.py:40-71) over the bytes of stdin, without
// a Python string per read.  Same rules: universal newlines ("\n", "\r\n", lone "\r");
// every line is stripped of leading / trailing white space; the first line is the number of reads n;
// the first read line decides paired ("|" present) vs unpaired; max(n, 1) read lines are consumed,
// missing ones are empty reads (unpaired) or an error (paired: a pair line must have exactly three
// "|" fields); the distance is the third field of the last pair line; trailing lines are ignored.
// Only plain ASCII is handled here (GA_ERR_ALPHABET otherwise: the caller then parses as text).
namespace {
inline bool ga_is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

struct LineScan {
    const uint8_t* p;
    const uint8_t* end;
    // next line, stripped: [lo, hi); false when the input is exhausted
    bool next(const uint8_t*& lo, const uint8_t*& hi) {
        if (p >= end) return false;
        // the line ends at the first "\n" or "\r" (memchr: the common case is one "\n" per ~100-200 bytes)
        const uint8_t* q = (const uint8_t*)memchr(p, '\n', (size_t)(end - p));
        if (!q) q = end;
        const uint8_t* cr = (const uint8_t*)memchr(p, '\r', (size_t)(q - p));
        if (cr) q = cr;
        lo = p;
        hi = q;
        if (q < end) p = (*q == '\r' && q + 1 < end && q[1] == '\n') ? q + 2 : q + 1;
        else p = end;
        while (lo < hi && ga_is_space(*lo)) ++lo;
        while (hi > lo && ga_is_space(hi[-1])) --hi;
        return true;
    }
};

bool ga_parse_int(const uint8_t* lo, const uint8_t* hi, int64_t* out) {
    while (lo < hi && ga_is_space(*lo)) ++lo;
    while (hi > lo && ga_is_space(hi[-1])) --hi;
    bool neg = false;
    if (lo < hi && (*lo == '+' || *lo == '-')) neg = *lo++ == '-';
    if (lo >= hi || hi - lo > 18) return false;
    int64_t v = 0;
    for (; lo < hi; ++lo) {
        if (*lo < '0' || *lo > '9') return false;
        v = v * 10 + (*lo - '0');
    }
    *out = neg ? -v : v;
    return true;
}
}  // namespace

namespace {

// One read line -> symbols + lengths.  WRITE = false only measures (same checks, no copies).
// Returns the symbols the line holds, or -1 for a pair line without exactly three "|" fields.
template <bool WRITE>
inline int64_t ga_take_line(const uint8_t* lo, const uint8_t* hi, bool paired, uint8_t* out, int32_t* lens, uint64_t r) {
    if (!paired) {
        const size_t n = (size_t)(hi - lo);
        if (n > 0x7FFFFFFFu) return -2;
        if (WRITE) {
            memcpy(out, lo, n);
            lens[r] = (int32_t)n;
        }
        return (int64_t)n;
    }
    const uint8_t* b1 = (const uint8_t*)memchr(lo, '|', (size_t)(hi - lo));
    const uint8_t* b2 = b1 ? (const uint8_t*)memchr(b1 + 1, '|', (size_t)(hi - b1 - 1)) : nullptr;
    if (!b1 || !b2 || memchr(b2 + 1, '|', (size_t)(hi - b2 - 1))) return -1;
    const size_t n1 = (size_t)(b1 - lo), n2 = (size_t)(b2 - b1 - 1);
    if (n1 > 0x7FFFFFFFu || n2 > 0x7FFFFFFFu) return -2;
    if (WRITE) {
        memcpy(out, lo, n1);
        memcpy(out + n1, b1 + 1, n2);
        lens[2 * r] = (int32_t)n1;
        lens[2 * r + 1] = (int32_t)n2;
    }
    return (int64_t)(n1 + n2);
}

// Parallel body of ga_parse_reads for big inputs whose only line break is "\n" (anything else takes the serial
// scan): the read lines are cut into one byte range per thread at line breaks; the lines of every range are
// counted (memchr), a prefix sum gives every range its first read index; pass 1 checks and measures the lines,
// a second prefix sum gives every range its output position; pass 2 copies.  Same result as the serial loop, line
// for line (tests/test_ingest.py runs every case through both).
struct ParseRange {
    const uint8_t *lo = nullptr, *hi = nullptr;
    uint64_t lines = 0, symbols = 0, first_line = 0, first_symbol = 0;
    int64_t bad_line = -1;
    int bad_kind = 0;
};

// pass 2 of the parallel parse: copy the range's lines (pass 1 has checked them)
void ga_write_range(const ParseRange& pr, bool paired, uint64_t count, uint8_t* symbols_out, int32_t* lens_out) {
    LineScan scan{pr.lo, pr.hi};
    const uint8_t *lo, *hi;
    uint64_t r = pr.first_line;
    uint8_t* out = symbols_out + pr.first_symbol;
    while (r < count && scan.next(lo, hi)) {
        out += ga_take_line<true>(lo, hi, paired, out, lens_out, r);
        ++r;
    }
}

}  // namespace

extern "C" int ga_parse_reads(const uint8_t* text, uint64_t n_bytes, uint8_t* symbols_out, int32_t* lens_out,
                              uint64_t lens_capacity, uint64_t* n_reads_out, int* paired_out, int64_t* distance_out,
                              uint64_t* n_symbols_out) {
    if (!text || !n_reads_out || !paired_out || !distance_out || !n_symbols_out) {
        ga_set_error("ga_parse_reads: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    // threads for the byte-wide passes: GA_PARSE_THREADS, else the cores (at most 16), one per 4 MB at least
    unsigned threads = std::thread::hardware_concurrency();
    if (threads == 0 || threads > 16) threads = threads == 0 ? 1 : 16;
    if (const char* e = getenv("GA_PARSE_THREADS")) threads = (unsigned)(atoi(e) > 0 ? atoi(e) : 1);
    const uint64_t grain = getenv("GA_PARSE_GRAIN") ? (uint64_t)atoll(getenv("GA_PARSE_GRAIN")) : (4ull << 20);
    if (n_bytes / (grain ? grain : 1) + 1 < threads) threads = (unsigned)(n_bytes / (grain ? grain : 1) + 1);
    LineScan scan{text, text + n_bytes};
    const uint8_t *lo = text, *hi = text;
    int64_t wanted = 0;
    if (!scan.next(lo, hi)) lo = hi = text;
    if (!ga_parse_int(lo, hi, &wanted)) {
        ga_set_error("ga_parse_reads: the first line is not an integer");
        return GA_ERR_BAD_ARG;
    }
    const uint64_t count = wanted < 1 ? 1ull : (uint64_t)wanted;
    // the first read line decides the input kind
    LineScan peek = scan;
    const uint8_t *flo = text, *fhi = text;
    if (!peek.next(flo, fhi)) flo = fhi = text;
    const bool paired = memchr(flo, '|', (size_t)(fhi - flo)) != nullptr;
    *paired_out = paired ? 1 : 0;
    *n_reads_out = count;
    *distance_out = 0;
    if (!symbols_out || !lens_out) {          // sizing call (the bytes themselves are looked at by the parsing call)
        *n_symbols_out = n_bytes;
        return GA_OK;
    }
    bool has_cr = false;
    {   // plain ASCII?  eight bytes at a time; and is "\n" the only line break?
        std::vector<uint64_t> acc(threads, 0);
        std::vector<char> cr(threads, 0);
        auto body = [&](unsigned t) {
            const uint64_t lo = n_bytes * t / threads, hi = n_bytes * (t + 1) / threads;
            uint64_t a = 0, i = lo;
            for (; i + 8 <= hi; i += 8) {
                uint64_t w;
                memcpy(&w, text + i, 8);
                a |= w;
            }
            for (; i < hi; ++i) a |= text[i];
            acc[t] = a;
            cr[t] = memchr(text + lo, '\r', (size_t)(hi - lo)) != nullptr;
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; ++t) pool.emplace_back(body, t);
        body(0);
        for (auto& th : pool) th.join();
        uint64_t all = 0;
        for (unsigned t = 0; t < threads; ++t) {
            all |= acc[t];
            has_cr = has_cr || cr[t];
        }
        if (all & 0x8080808080808080ull) {
            ga_set_error("ga_parse_reads: non-ASCII input");
            return GA_ERR_ALPHABET;
        }
    }
    if (lens_capacity < count * (paired ? 2u : 1u)) {
        ga_set_error("ga_parse_reads: length array too small");
        return GA_ERR_CAPACITY;
    }
    const uint8_t* body_lo = scan.p;
    const uint8_t* body_hi = text + n_bytes;
    if (threads > 1 && !has_cr && body_hi > body_lo) {
        // ranges end right after a "\n" (a range never starts inside a line)
        std::vector<ParseRange> ranges;
        const uint64_t span = (uint64_t)(body_hi - body_lo);
        const uint8_t* at = body_lo;
        for (unsigned t = 0; t < threads && at < body_hi; ++t) {
            const uint8_t* want = t + 1 == threads ? body_hi : body_lo + span * (t + 1) / threads;
            if (want < at) want = at;
            const uint8_t* stop = want >= body_hi ? body_hi : (const uint8_t*)memchr(want, '\n', (size_t)(body_hi - want));
            stop = stop ? (stop < body_hi ? stop + 1 : body_hi) : body_hi;
            ParseRange pr;
            pr.lo = at;
            pr.hi = stop;
            ranges.push_back(pr);
            at = stop;
        }
        {   // lines per range (memchr only), so that every range knows its first read index
            std::vector<std::thread> pool;
            auto count_lines = [&](size_t t) {
                uint64_t n = 0;
                const uint8_t* p = ranges[t].lo;
                while (p < ranges[t].hi) {
                    const uint8_t* q = (const uint8_t*)memchr(p, '\n', (size_t)(ranges[t].hi - p));
                    ++n;
                    if (!q) break;
                    p = q + 1;
                }
                ranges[t].lines = n;
            };
            for (size_t t = 1; t < ranges.size(); ++t) pool.emplace_back(count_lines, t);
            count_lines(0);
            for (auto& th : pool) th.join();
        }
        uint64_t line = 0;
        for (auto& pr : ranges) {
            pr.first_line = line;
            line += pr.lines;
        }
        const uint64_t available = line;
        // a range that starts at or beyond `count` has nothing to do; one that straddles it stops there
        std::vector<ParseRange> live;
        for (auto& pr : ranges)
            if (pr.first_line < count) live.push_back(pr);
        ranges.swap(live);
        {   // pass 1: symbols per range (stops at `count`)
            std::vector<std::thread> pool;
            auto measure = [&](size_t t) {
                ParseRange& pr = ranges[t];
                LineScan sc{pr.lo, pr.hi};
                const uint8_t *l, *h;
                uint64_t r = pr.first_line, sym = 0;
                while (r < count && sc.next(l, h)) {
                    const int64_t n = ga_take_line<false>(l, h, paired, nullptr, nullptr, r);
                    if (n < 0) {
                        pr.bad_line = (int64_t)r;
                        pr.bad_kind = (int)n;
                        break;
                    }
                    sym += (uint64_t)n;
                    ++r;
                }
                pr.symbols = sym;
            };
            for (size_t t = 1; t < ranges.size(); ++t) pool.emplace_back(measure, t);
            if (!ranges.empty()) measure(0);
            for (auto& th : pool) th.join();
        }
        for (auto& pr : ranges)
            if (pr.bad_line >= 0) {
                if (pr.bad_kind == -2) return GA_ERR_CAPACITY;
                ga_set_error("ga_parse_reads: read-pair line %llu does not have three '|' fields",
                             (unsigned long long)(pr.bad_line + 1));
                return GA_ERR_BAD_ARG;
            }
        uint64_t sym = 0;
        for (auto& pr : ranges) {
            pr.first_symbol = sym;
            sym += pr.symbols;
        }
        {   // pass 2: copy
            std::vector<std::thread> pool;
            for (size_t t = 1; t < ranges.size(); ++t)
                pool.emplace_back([&, t] { ga_write_range(ranges[t], paired, count, symbols_out, lens_out); });
            if (!ranges.empty()) ga_write_range(ranges[0], paired, count, symbols_out, lens_out);
            for (auto& th : pool) th.join();
        }
        const uint64_t got = available < count ? available : count;
        if (got < count) {                      // missing lines: empty reads, or an error for pairs
            if (paired) {
                ga_set_error("ga_parse_reads: read-pair line %llu does not have three '|' fields", (unsigned long long)(got + 1));
                return GA_ERR_BAD_ARG;
            }
            for (uint64_t r = got; r < count; ++r) lens_out[r] = 0;
        }
        if (paired) {                           // the distance is the third field of the LAST pair line
            const ParseRange* last = nullptr;
            for (auto& pr : ranges)
                if (pr.first_line <= count - 1) last = &pr;
            LineScan sc{last->lo, last->hi};
            const uint8_t *l = last->lo, *h = last->lo;
            for (uint64_t r = last->first_line; r <= count - 1; ++r) sc.next(l, h);
            const uint8_t* b1 = (const uint8_t*)memchr(l, '|', (size_t)(h - l));
            const uint8_t* b2 = (const uint8_t*)memchr(b1 + 1, '|', (size_t)(h - b1 - 1));
            if (!ga_parse_int(b2 + 1, h, distance_out)) {
                // not a plain decimal number: Python's int() decides (it takes "1_0", refuses "x"), as upstream
                ga_set_error("ga_parse_reads: the distance field of the last pair is not a plain integer");
                return GA_ERR_ALPHABET;
            }
        }
        *n_symbols_out = sym;
        return GA_OK;
    }
    uint8_t* out = symbols_out;
    for (uint64_t r = 0; r < count; ++r) {
        if (!scan.next(lo, hi)) lo = hi = text;
        const int64_t n = ga_take_line<true>(lo, hi, paired, out, lens_out, r);
        if (n == -2) return GA_ERR_CAPACITY;
        if (n < 0) {
            ga_set_error("ga_parse_reads: read-pair line %llu does not have three '|' fields",
                         (unsigned long long)(r + 1));
            return GA_ERR_BAD_ARG;
        }
        out += n;
        if (paired && r + 1 == count) {
            const uint8_t* b1 = (const uint8_t*)memchr(lo, '|', (size_t)(hi - lo));
            const uint8_t* b2 = (const uint8_t*)memchr(b1 + 1, '|', (size_t)(hi - b1 - 1));
            if (!ga_parse_int(b2 + 1, hi, distance_out)) {
                ga_set_error("ga_parse_reads: the distance field of the last pair is not a plain integer");
                return GA_ERR_ALPHABET;
            }
        }
    }
    *n_symbols_out = (uint64_t)(out - symbols_out);
    return GA_OK;
}
