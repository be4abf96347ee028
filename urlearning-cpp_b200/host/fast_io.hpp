// fast_io.hpp — the two host-side Amdahl terms of `score`: reading the CSV and formatting the `.pss`.
//
//   parse_csv      datastructures::RecordFile::read + BayesianNetwork::initialize (base/record_file.h:39-54,
//                  base/record.h:35-39, base/bayesian_network.cpp:25-42, base/variable.h:43-48) in one parallel pass:
//                  the file is cut at line boundaries, every thread tokenises its lines (trim, split on the delimiter
//                  with token compression) and codes each field against a per-column dictionary of the strings it has
//                  seen; the chunk dictionaries are merged in file order, which reproduces the reference's
//                  first-appearance value order exactly.  Value identity is the STRING ("1" != "1.0").
//   format_score   printf("%f ", (double)score) (score_main.cpp:191) for a float, exactly: a float is m * 2^e with a 24-bit
//                  m, so score * 10^6 is an exact integer ratio and round-half-even on it is what glibc prints.
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace urlhost {

struct ParsedCsv {
    int p = 0;
    int64_t n = 0;
    std::vector<std::string> header;                 // empty without -s
    std::vector<std::vector<std::string>> values;    // [p]: distinct strings of the column in first-appearance order
    std::vector<std::vector<int32_t>> codes;         // [p][n]: index into values[column]
};

namespace detail {
struct ColumnDict { // strings seen in one column by one thread, in first-appearance order
    std::vector<std::string> vals;
    std::unordered_map<std::string, int32_t> map;    // only used once the column has more than kLinear values
    static constexpr size_t kLinear = 12;
    int last = 0;
    int32_t code(const char *s, size_t len) {
        if (!vals.empty()) {
            const std::string &l = vals[last];
            if (l.size() == len && memcmp(l.data(), s, len) == 0) return last;
        }
        if (vals.size() <= kLinear) {
            for (size_t i = 0; i < vals.size(); i++)
                if (vals[i].size() == len && memcmp(vals[i].data(), s, len) == 0) { last = (int)i; return (int32_t)i; }
        } else {
            auto it = map.find(std::string(s, len));
            if (it != map.end()) { last = it->second; return it->second; }
        }
        vals.emplace_back(s, len);
        const int32_t c = (int32_t)vals.size() - 1;
        if (vals.size() == kLinear + 1) for (size_t i = 0; i < vals.size(); i++) map.emplace(vals[i], (int32_t)i);
        else if (vals.size() > kLinear + 1) map.emplace(vals.back(), c);
        last = c;
        return c;
    }
};
// fields of one line: trim, then split on `delim` with runs counting once (boost token_compress_on); an empty trimmed line is one empty field
template <typename F> inline int tokenize_line(const char *b, const char *e, char delim, F &&field) {
    while (b < e && std::isspace((unsigned char)*b)) b++;
    while (e > b && std::isspace((unsigned char)e[-1])) e--;
    int k = 0;
    const char *t = b;
    for (const char *q = b;; q++) {
        if (q == e || *q == delim) {
            field(k++, t, (size_t)(q - t));
            if (q == e) break;
            while (q + 1 < e && q[1] == delim) q++;
            t = q + 1;
        }
    }
    return k;
}
} // namespace detail

inline ParsedCsv parse_csv(const std::string &filename, char delimiter, bool hasHeader, int threads = 0) {
    std::ifstream file(filename, std::ios::binary | std::ios::ate);
    if (!file.good()) throw std::runtime_error("Could not open the input file: '" + filename + "'");
    const std::streamsize size = file.tellg();
    file.seekg(0);
    std::string buf((size_t)size, '\0');
    if (size && !file.read(&buf[0], size)) throw std::runtime_error("Could not read the input file: '" + filename + "'");
    const char *base = buf.data(), *end = base + buf.size();
    ParsedCsv out;
    const char *cur = base;
    auto next_line = [&](const char *s) { const char *nl = (const char *)memchr(s, '\n', (size_t)(end - s)); return nl ? nl : end; };
    if (hasHeader && cur < end) {
        const char *nl = next_line(cur);
        detail::tokenize_line(cur, nl, delimiter, [&](int, const char *s, size_t len) { out.header.emplace_back(s, len); });
        cur = nl < end ? nl + 1 : end;
    }
    if (cur >= end) throw std::runtime_error("The input file has no records: '" + filename + "'");
    { // width = fields of the first record
        int w = 0;
        detail::tokenize_line(cur, next_line(cur), delimiter, [&](int, const char *, size_t) { w++; });
        out.p = w;
    }
    const int p = out.p;
    int T = threads > 0 ? threads : (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    const size_t body = (size_t)(end - cur);
    T = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, body / (1 << 20) + 1));
    std::vector<const char *> cut(T + 1);
    cut[0] = cur;
    cut[T] = end;
    for (int t = 1; t < T; t++) {
        const char *q = cur + body * (size_t)t / (size_t)T;
        q = next_line(std::max(q, cut[t - 1]));
        cut[t] = q < end ? q + 1 : end;
    }
    struct Chunk { std::vector<detail::ColumnDict> dict; std::vector<std::vector<int32_t>> codes; int64_t n = 0; std::string err; };
    std::vector<Chunk> chunks(T);
    auto work = [&](int t) {
        Chunk &c = chunks[t];
        c.dict.resize(p);
        c.codes.resize(p);
        const size_t guess = (size_t)(cut[t + 1] - cut[t]) / (size_t)std::max(2 * p, 2) + 16;
        for (auto &v : c.codes) v.reserve(guess);
        for (const char *s = cut[t]; s < cut[t + 1];) {
            const char *nl = next_line(s);
            if (nl > cut[t + 1]) nl = cut[t + 1];
            const int k = detail::tokenize_line(s, nl, delimiter, [&](int j, const char *f, size_t len) { if (j < p) c.codes[j].push_back(c.dict[j].code(f, len)); });
            if (k < p) { c.err = "Record " + std::to_string(c.n + 1) + " of chunk " + std::to_string(t) + " has fewer fields than the first record"; return; }
            c.n++;
            s = nl < end ? nl + 1 : end;
        }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (auto &c : chunks) if (!c.err.empty()) throw std::runtime_error(c.err);
    // merge the dictionaries in file order: global order = first appearance
    out.values.resize(p);
    out.codes.resize(p);
    for (auto &c : chunks) out.n += c.n;
    std::vector<std::vector<std::vector<int32_t>>> remap(T, std::vector<std::vector<int32_t>>(p));
    for (int j = 0; j < p; j++) {
        std::unordered_map<std::string, int32_t> global;
        for (int t = 0; t < T; t++) {
            auto &vals = chunks[t].dict[j].vals;
            remap[t][j].resize(vals.size());
            for (size_t i = 0; i < vals.size(); i++) {
                auto it = global.find(vals[i]);
                if (it == global.end()) { it = global.emplace(vals[i], (int32_t)out.values[j].size()).first; out.values[j].push_back(vals[i]); }
                remap[t][j][i] = it->second;
            }
        }
    }
    std::vector<int64_t> first(T + 1, 0);
    for (int t = 0; t < T; t++) first[t + 1] = first[t] + chunks[t].n;
    for (int j = 0; j < p; j++) out.codes[j].resize((size_t)out.n);
    auto fill = [&](int t) {
        for (int j = 0; j < p; j++) {
            const auto &src = chunks[t].codes[j];
            const auto &rm = remap[t][j];
            int32_t *dst = out.codes[j].data() + first[t];
            for (size_t i = 0; i < src.size(); i++) dst[i] = rm[src[i]];
        }
    };
    if (T == 1) fill(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(fill, t);
        for (auto &x : th) x.join();
    }
    return out;
}

// appends printf("%f", (double)x) to out (exact: see the file header); returns the number of characters
inline int format_score(float x, char *out) {
    uint32_t bits;
    memcpy(&bits, &x, 4);
    const bool neg = bits >> 31;
    const int ex = (int)((bits >> 23) & 0xff);
    uint32_t man = bits & 0x7fffffu;
    if (ex == 0xff || ex - 150 > 20) return snprintf(out, 64, "%f", (double)x); // inf / nan / |x| >= 2^44: rare, let libc do it
    int e;
    if (ex == 0) e = -149; else { man |= 0x800000u; e = ex - 150; }                // x = man * 2^e
    unsigned long long q;                                                           // round(|x| * 10^6)
    if (e >= 0) q = ((unsigned long long)man << e) * 1000000ull;
    else {
        const unsigned long long num = (unsigned long long)man * 1000000ull;       // < 2^44
        const int k = -e;
        if (k >= 64) q = 0;
        else {
            q = num >> k;
            const unsigned long long rem = num & ((1ull << k) - 1), half = 1ull << (k - 1);
            if (rem > half || (rem == half && (q & 1))) q++;
        }
    }
    char tmp[32];
    int len = 0;
    unsigned long long ip = q / 1000000ull;
    unsigned fr = (unsigned)(q % 1000000ull);
    do { tmp[len++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
    int o = 0;
    if (neg) out[o++] = '-';
    while (len) out[o++] = tmp[--len];
    out[o++] = '.';
    for (int d = 100000; d >= 1; d /= 10) { out[o++] = (char)('0' + fr / d); fr %= d; }
    out[o] = 0;
    return o;
}

} // namespace urlhost
