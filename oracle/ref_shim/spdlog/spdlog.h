// stand-in for spdlog (not installed here): the two reference files only report a fatal inconsistency through it
#pragma once
#include <cstdio>
namespace spdlog {
template <typename... A> inline void critical(const char* msg, A...) { std::fprintf(stderr, "[critical] %s\n", msg); }
template <typename... A> inline void warn(const char* msg, A...) { std::fprintf(stderr, "[warn] %s\n", msg); }
}  // namespace spdlog
