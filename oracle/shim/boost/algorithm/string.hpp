// Shim for the subset of Boost.StringAlgo the reference's hot-path sources use (boost is not installed here).
// Written from the documented behaviour of boost::trim / boost::split / boost::is_any_of / token_compress_on.
#pragma once
#include <algorithm>
#include <cctype>
#include <string>
#include <vector>
namespace boost {
enum token_compress_mode_type { token_compress_on, token_compress_off };
namespace algorithm {
using boost::token_compress_mode_type;
using boost::token_compress_on;
using boost::token_compress_off;
struct is_any_of_pred {
    std::string set;
    bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline is_any_of_pred is_any_of(const std::string &s) { return is_any_of_pred{s}; }
inline void trim(std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    s = s.substr(a, b - a);
}
inline void to_lower(std::string &s) { for (auto &c : s) c = (char)std::tolower((unsigned char)c); }
inline void to_upper(std::string &s) { for (auto &c : s) c = (char)std::toupper((unsigned char)c); }
// split: every separator ends a token; with token_compress_on a run of separators counts as one.
template <class Seq, class Pred>
inline Seq &split(Seq &out, const std::string &in, Pred pred, token_compress_mode_type mode = token_compress_off) {
    out.clear();
    std::string cur;
    size_t i = 0;
    while (true) {
        if (i == in.size()) { out.push_back(cur); break; }
        if (pred(in[i])) {
            out.push_back(cur);
            cur.clear();
            i++;
            if (mode == token_compress_on) while (i < in.size() && pred(in[i])) i++;
            continue;
        }
        cur.push_back(in[i++]);
    }
    return out;
}
inline bool contains(const std::string &a, const std::string &b) { return a.find(b) != std::string::npos; }
} // namespace algorithm
using algorithm::is_any_of;
using algorithm::split;
using algorithm::trim;
using algorithm::to_lower;
using algorithm::to_upper;
using algorithm::contains;
} // namespace boost
