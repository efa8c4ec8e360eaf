#include <spdlog/spdlog.h>
