// shim: just enough of boost::program_options for the reference's main() functions to COMPILE; the pin drivers never call them
// (they set the option globals themselves), so nothing here parses anything
#pragma once
#include <iostream>
#include <string>
namespace boost { namespace program_options {
template <class T> struct typed_value {
    T *target;
    typed_value *required() { return this; }
    template <class V> typed_value *default_value(const V &v) { if (target) *target = v; return this; }
};
template <class T> typed_value<T> *value(T *target = nullptr) { return new typed_value<T>{target}; }
struct options_description_easy_init {
    template <class V> options_description_easy_init &operator()(const char *, V *, const char *) { return *this; }
    options_description_easy_init &operator()(const char *, const char *) { return *this; }
};
struct options_description {
    options_description(const std::string & = "") {}
    options_description_easy_init add_options() { return {}; }
};
inline std::ostream &operator<<(std::ostream &o, const options_description &) { return o; }
struct positional_options_description { positional_options_description &add(const char *, int) { return *this; } };
struct variables_map { int count(const char *) const { return 0; } };
struct parsed_options {};
struct command_line_parser {
    command_line_parser(int, char **) {}
    command_line_parser &options(const options_description &) { return *this; }
    command_line_parser &positional(const positional_options_description &) { return *this; }
    parsed_options run() { return {}; }
};
inline void store(const parsed_options &, variables_map &) {}
inline void notify(variables_map &) {}
} }
