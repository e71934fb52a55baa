// stdafx.hpp -- common includes and the KF_* logging / assertion macros of the host facade.
// Behaviour follows reference include/stdafx.hpp:49-67: KF_CRITICAL logs and throws
// std::runtime_error, KF_ASSERT throws when its condition fails.
#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "kfglm.hpp"

#if defined(_WIN32)
#define KUAFU_API __declspec(dllexport)
#else
#define KUAFU_API __attribute__((visibility("default")))
#endif

namespace kuafu {

// A very small logger with the part of the spdlog interface the reference uses
// (global::logger->warn(...), set_level); spdlog itself is not available here.
namespace log_level {
enum Level { trace = 0, debug = 1, info = 2, warn = 3, err = 4, critical = 5, off = 6 };
}

class KUAFU_API Logger {
 public:
  void set_level(int level) { mLevel = level; }
  int level() const { return mLevel; }
  template <typename... Args> void debug(Args&&... a) { emit(log_level::debug, "debug", a...); }
  template <typename... Args> void info(Args&&... a) { emit(log_level::info, "info", a...); }
  template <typename... Args> void warn(Args&&... a) { emit(log_level::warn, "warning", a...); }
  template <typename... Args> void error(Args&&... a) { emit(log_level::err, "error", a...); }
  template <typename... Args> void critical(Args&&... a) { emit(log_level::critical, "critical", a...); }

 private:
  template <typename... Args>
  void emit(int lvl, const char* tag, Args&&... a) {
    if (lvl < mLevel) return;
    std::ostringstream os;
    (os << ... << a);
    std::fprintf(stderr, "[kuafu] [%s] %s\n", tag, os.str().c_str());
  }
  int mLevel = log_level::warn;
};

}  // namespace kuafu

#define KF_DEBUG(...) ::kuafu::global::logger->debug(__VA_ARGS__)
#define KF_INFO(...) ::kuafu::global::logger->info(__VA_ARGS__)
#define KF_WARN(...) ::kuafu::global::logger->warn(__VA_ARGS__)
#define KF_ERROR(...) ::kuafu::global::logger->error(__VA_ARGS__)
#define KF_CRITICAL(...)                                 \
  do {                                                   \
    ::kuafu::global::logger->critical(__VA_ARGS__);      \
    std::ostringstream _kf_os;                           \
    ::kuafu::detail::streamAll(_kf_os, __VA_ARGS__);     \
    throw std::runtime_error(_kf_os.str());              \
  } while (0)
#define KF_ASSERT(cond, ...)                  \
  do {                                        \
    if (!(cond)) { KF_CRITICAL(__VA_ARGS__); } \
  } while (0)

namespace kuafu::detail {
template <typename... Args>
void streamAll(std::ostringstream& os, Args&&... a) {
  (os << ... << a);
}
}  // namespace kuafu::detail
