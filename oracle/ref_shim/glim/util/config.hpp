// stand-in for glim/util/config.hpp (the JSON config system: nlohmann + spdlog, OUT of the hot-path scope).  Shadows the
// reference header because ref_shim/ comes first on the include path: every param<T>() returns the CODE default the caller
// passes, which is what CloudPreprocessorParams() would see with an empty config file; the glue then sets the fields it tests.
#pragma once
#include <string>
namespace glim {
class Config {
public:
  explicit Config(const std::string&) {}
  template <typename T> T param(const std::string&, const std::string&, const T& default_value) const { return default_value; }
};
class GlobalConfig {
public:
  static std::string get_config_path(const std::string&) { return std::string(); }
};
}  // namespace glim
