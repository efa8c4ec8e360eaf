// stand-in for spdlog and the fmt it bundles (neither is installed here): messages are dropped, fmt::format returns the pattern
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
namespace fmt {
template <typename... A> inline std::string format(const char* pattern, const A&...) { return std::string(pattern); }
}  // namespace fmt
namespace spdlog {
template <typename... A> inline void trace(const char*, const A&...) {}
template <typename... A> inline void debug(const char*, const A&...) {}
template <typename... A> inline void info(const char*, const A&...) {}
template <typename... A> inline void warn(const char*, const A&...) {}
template <typename... A> inline void error(const char* msg, const A&...) { std::fprintf(stderr, "[error] %s\n", msg); }
template <typename... A> inline void critical(const char* msg, const A&...) { std::fprintf(stderr, "[critical] %s\n", msg); }
}  // namespace spdlog
