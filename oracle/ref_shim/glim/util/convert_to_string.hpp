// stand-in for glim/util/convert_to_string.hpp (fmt-based pretty printing, only used in an error message of the path)
#pragma once
#include <string>
namespace glim {
template <typename T> std::string convert_to_string(const T&) { return std::string("<value>"); }
}  // namespace glim
