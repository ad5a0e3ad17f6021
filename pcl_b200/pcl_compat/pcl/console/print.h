// pcl/console/print.h — the logging front of PCL that user programs reach through the PCL_ERROR / PCL_WARN / PCL_INFO /
// PCL_DEBUG macros (common/include/pcl/console/print.h:60-120, common/src/print.cpp): a verbosity level (default L_INFO;
// the PCL_VERBOSITY_LEVEL environment variable is not read — this library reads none) and printf-style printing, errors
// and warnings to stderr, the rest to stdout.  No colours.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <iostream>
#include <sstream>

namespace pcl {
namespace console {
enum VERBOSITY_LEVEL { L_ALWAYS, L_ERROR, L_WARN, L_INFO, L_DEBUG, L_VERBOSE };
namespace detail {
inline VERBOSITY_LEVEL& level()
{
  static VERBOSITY_LEVEL l = L_INFO;
  return l;
}
}  // namespace detail
inline void setVerbosityLevel(VERBOSITY_LEVEL level) { detail::level() = level; }
inline VERBOSITY_LEVEL getVerbosityLevel() { return detail::level(); }
inline bool isVerbosityLevelEnabled(VERBOSITY_LEVEL level) { return level <= detail::level(); }
inline void print(VERBOSITY_LEVEL level, const char* format, ...)
{
  if (!isVerbosityLevelEnabled(level)) return;
  FILE* stream = (level == L_WARN || level == L_ERROR) ? stderr : stdout;
  va_list ap;
  va_start(ap, format);
  std::vfprintf(stream, format, ap);
  va_end(ap);
}
inline void print(VERBOSITY_LEVEL level, FILE* stream, const char* format, ...)
{
  if (!isVerbosityLevelEnabled(level)) return;
  va_list ap;
  va_start(ap, format);
  std::vfprintf(stream, format, ap);
  va_end(ap);
}
// print.h:121-290: the named forms a command-line tool prints with (no colours here); each also exists with a FILE* first
#define PCLB_PRINT_FN(name, lvl, dflt)                                  \
  inline void name(const char* format, ...)                             \
  {                                                                     \
    if (!isVerbosityLevelEnabled(lvl)) return;                          \
    va_list ap;                                                         \
    va_start(ap, format);                                               \
    std::vfprintf(dflt, format, ap);                                    \
    va_end(ap);                                                         \
  }                                                                     \
  inline void name(FILE* stream, const char* format, ...)               \
  {                                                                     \
    if (!isVerbosityLevelEnabled(lvl)) return;                          \
    va_list ap;                                                         \
    va_start(ap, format);                                               \
    std::vfprintf(stream, format, ap);                                  \
    va_end(ap);                                                         \
  }
PCLB_PRINT_FN(print_info, L_INFO, stdout)
PCLB_PRINT_FN(print_value, L_INFO, stdout)
PCLB_PRINT_FN(print_error, L_ERROR, stderr)
PCLB_PRINT_FN(print_warn, L_WARN, stderr)
PCLB_PRINT_FN(print_debug, L_DEBUG, stdout)
#undef PCLB_PRINT_FN
// print_highlight prefixes "> " (print.cpp:150-172)
inline void print_highlight(const char* format, ...)
{
  if (!isVerbosityLevelEnabled(L_ALWAYS)) return;
  std::fputs("> ", stdout);
  va_list ap;
  va_start(ap, format);
  std::vfprintf(stdout, format, ap);
  va_end(ap);
}
inline void print_highlight(FILE* stream, const char* format, ...)
{
  if (!isVerbosityLevelEnabled(L_ALWAYS)) return;
  std::fputs("> ", stream);
  va_list ap;
  va_start(ap, format);
  std::vfprintf(stream, format, ap);
  va_end(ap);
}
}  // namespace console
}  // namespace pcl

#define PCL_LOG_STREAM(LEVEL, STREAM, CSTR, ATTR, FG, ARGS)                         \
  if (pcl::console::isVerbosityLevelEnabled(pcl::console::LEVEL)) {                 \
    std::ostringstream pcl_log_stream_;                                             \
    pcl_log_stream_ << ARGS;                                                        \
    std::fputs(pcl_log_stream_.str().c_str(), CSTR);                                \
  }
#define PCL_ALWAYS_STREAM(ARGS) PCL_LOG_STREAM(L_ALWAYS, std::cout, stdout, 0, 0, ARGS)
#define PCL_ERROR_STREAM(ARGS) PCL_LOG_STREAM(L_ERROR, std::cerr, stderr, 0, 0, ARGS)
#define PCL_WARN_STREAM(ARGS) PCL_LOG_STREAM(L_WARN, std::cerr, stderr, 0, 0, ARGS)
#define PCL_INFO_STREAM(ARGS) PCL_LOG_STREAM(L_INFO, std::cout, stdout, 0, 0, ARGS)
#define PCL_DEBUG_STREAM(ARGS) PCL_LOG_STREAM(L_DEBUG, std::cout, stdout, 0, 0, ARGS)
#define PCL_VERBOSE_STREAM(ARGS) PCL_LOG_STREAM(L_VERBOSE, std::cout, stdout, 0, 0, ARGS)

#define PCL_ALWAYS(...) pcl::console::print(pcl::console::L_ALWAYS, __VA_ARGS__)
#define PCL_ERROR(...) pcl::console::print(pcl::console::L_ERROR, __VA_ARGS__)
#define PCL_WARN(...) pcl::console::print(pcl::console::L_WARN, __VA_ARGS__)
#define PCL_INFO(...) pcl::console::print(pcl::console::L_INFO, __VA_ARGS__)
#define PCL_DEBUG(...) pcl::console::print(pcl::console::L_DEBUG, __VA_ARGS__)
#define PCL_VERBOSE(...) pcl::console::print(pcl::console::L_VERBOSE, __VA_ARGS__)
