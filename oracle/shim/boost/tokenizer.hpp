// Shim: boost::tokenizer<boost::char_separator<char>> with dropped delimiters and empty tokens dropped.
#pragma once
#include <cstring>
#include <string>
#include <vector>
namespace boost {
template <class Char = char>
class char_separator {
public:
    explicit char_separator(const Char *dropped = " ") : d(dropped) {}
    std::string d;
};
template <class Sep = char_separator<char>>
class tokenizer {
public:
    typedef std::vector<std::string>::const_iterator iterator;
    tokenizer(const std::string &s, const Sep &sep) {
        std::string cur;
        for (char c : s) {
            if (sep.d.find(c) != std::string::npos) { if (!cur.empty()) { toks.push_back(cur); cur.clear(); } }
            else cur.push_back(c);
        }
        if (!cur.empty()) toks.push_back(cur);
    }
    iterator begin() const { return toks.begin(); }
    iterator end() const { return toks.end(); }
private:
    std::vector<std::string> toks;
};
} // namespace boost
