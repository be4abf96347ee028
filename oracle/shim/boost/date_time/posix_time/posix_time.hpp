#pragma once
#include "../../asio.hpp"
