// Shim: boost/algorithm/string/trim.hpp -> the string-algorithm shim next door (boost is not installed here).
#pragma once
#include "../string.hpp"
