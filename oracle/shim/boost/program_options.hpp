// shim: the part of boost::program_options the reference's main() functions use, working: "long,s" option names, typed values
// bound to variables with default_value / required, flags, positional arguments, --long=value / --long value / -s value.
// (TEST INFRASTRUCTURE: lets the reference's own `score` main() run here; Boost is absent.)
#pragma once
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace boost { namespace program_options {
namespace detail {
template <class T> inline void assign(T &dst, const std::string &s) { std::istringstream in(s); in >> dst; if (in.fail()) throw std::runtime_error("the argument ('" + s + "') is invalid"); }
template <> inline void assign<std::string>(std::string &dst, const std::string &s) { dst = s; }
template <> inline void assign<char>(char &dst, const std::string &s) { dst = s.empty() ? '\0' : s[0]; }
struct option {
    std::string longName, shortName;
    bool takesValue = false;
    std::function<void(const std::string &)> set;
};
}
template <class T> struct typed_value {
    T *target;
    typed_value *required() { return this; }
    template <class V> typed_value *default_value(const V &v) { if (target) *target = (T)v; return this; }
};
template <class T> typed_value<T> *value(T *target = nullptr) { return new typed_value<T>{target}; }
struct options_description;
struct options_description_easy_init {
    options_description *owner;
    template <class T> options_description_easy_init &operator()(const char *name, typed_value<T> *v, const char *);
    options_description_easy_init &operator()(const char *name, const char *);
};
struct options_description {
    std::string caption;
    std::vector<detail::option> options;
    options_description(const std::string &c = "") : caption(c) {}
    options_description_easy_init add_options() { return options_description_easy_init{this}; }
    void add(const char *name, bool takesValue, std::function<void(const std::string &)> set) {
        detail::option o;
        const std::string n = name;
        const size_t comma = n.find(',');
        o.longName = n.substr(0, comma);
        if (comma != std::string::npos) o.shortName = n.substr(comma + 1);
        o.takesValue = takesValue;
        o.set = set;
        options.push_back(o);
    }
};
template <class T> options_description_easy_init &options_description_easy_init::operator()(const char *name, typed_value<T> *v, const char *) {
    T *target = v->target;
    owner->add(name, true, [target](const std::string &s) { if (target) detail::assign(*target, s); });
    return *this;
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *) {
    owner->add(name, false, [](const std::string &) {});
    return *this;
}
inline std::ostream &operator<<(std::ostream &o, const options_description &d) { return o << d.caption << "\n"; }
struct positional_options_description {
    std::vector<std::string> names;
    positional_options_description &add(const char *name, int) { names.push_back(name); return *this; }
};
struct variables_map {
    std::map<std::string, int> seen;
    int count(const char *name) const { auto it = seen.find(name); return it == seen.end() ? 0 : it->second; }
};
struct parsed_options { std::vector<std::pair<const detail::option *, std::string>> hits; };
struct command_line_parser {
    int argc; char **argv;
    const options_description *desc = nullptr;
    const positional_options_description *pos = nullptr;
    command_line_parser(int c, char **v) : argc(c), argv(v) {}
    command_line_parser &options(const options_description &d) { desc = &d; return *this; }
    command_line_parser &positional(const positional_options_description &p) { pos = &p; return *this; }
    const detail::option *find(const std::string &n, bool isShort) const {
        for (auto &o : desc->options) if ((isShort ? o.shortName : o.longName) == n) return &o;
        throw std::runtime_error("unrecognised option '" + std::string(isShort ? "-" : "--") + n + "'");
    }
    parsed_options run() {
        parsed_options out;
        size_t nextPos = 0;
        for (int i = 1; i < argc; i++) {
            const std::string a = argv[i];
            if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
                const size_t eq = a.find('=');
                const detail::option *o = find(a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2), false);
                if (!o->takesValue) out.hits.push_back({o, ""});
                else if (eq != std::string::npos) out.hits.push_back({o, a.substr(eq + 1)});
                else { if (i + 1 >= argc) throw std::runtime_error("the required argument for option '" + a + "' is missing"); out.hits.push_back({o, argv[++i]}); }
            } else if (a.size() > 1 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
                const detail::option *o = find(a.substr(1, 1), true);
                if (!o->takesValue) out.hits.push_back({o, ""});
                else if (a.size() > 2) out.hits.push_back({o, a.substr(2)});
                else { if (i + 1 >= argc) throw std::runtime_error("the required argument for option '" + a + "' is missing"); out.hits.push_back({o, argv[++i]}); }
            } else {
                if (!pos || nextPos >= pos->names.size()) throw std::runtime_error("too many positional options have been specified on the command line");
                out.hits.push_back({find(pos->names[nextPos++], false), a});
            }
        }
        return out;
    }
};
inline void store(const parsed_options &p, variables_map &vm) {
    for (auto &h : p.hits) { h.first->set(h.second); vm.seen[h.first->longName]++; }
}
inline void notify(variables_map &) {}
} }
