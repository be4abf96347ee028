// Shim: boost::lexical_cast via iostreams with Boost's precision rule (max_digits10 for floating types).
#pragma once
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
namespace boost {
template <class Target, class Source>
inline Target lexical_cast(const Source &src) {
    std::stringstream ss;
    if (std::is_floating_point<Source>::value) ss.precision(std::numeric_limits<Source>::max_digits10);
    if (std::is_floating_point<Target>::value) ss.precision(std::numeric_limits<Target>::max_digits10);
    ss << src;
    Target out;
    if (std::is_same<Target, std::string>::value) { return *reinterpret_cast<Target *>(new std::string(ss.str())); }
    if (!(ss >> out)) throw std::runtime_error("bad lexical cast");
    return out;
}
} // namespace boost
