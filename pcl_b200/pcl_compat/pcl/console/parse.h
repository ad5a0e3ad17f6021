// pcl/console/parse.h — command-line helpers of the PCL tools (common/include/pcl/console/parse.h, common/src/parse.cpp:48-470):
// "-name value" look-ups with strict numeric conversion (-1 on junk or overflow, the value left alone), comma-separated
// tuples, and the indices of the arguments that end in a file extension.
#pragma once
#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "print.h"

namespace pcl {
namespace console {
inline int find_argument(int argc, const char* const* argv, const char* argument_name)
{
  for (int i = 1; i < argc; ++i)
    if (std::strcmp(argv[i], argument_name) == 0) return i;
  return -1;
}
inline bool find_switch(int argc, const char* const* argv, const char* argument_name) { return find_argument(argc, argv, argument_name) != -1; }

// every parse_argument returns the index of the option, or -1 when it is absent or its value does not convert
inline int parse_argument(int argc, const char* const* argv, const char* str, std::string& val)
{
  const int index = find_argument(argc, argv, str) + 1;
  if (index > 0 && index < argc) val = argv[index];
  return index - 1;
}
namespace detail {
template <typename T, typename Convert>
inline int parse_generic(Convert convert, int argc, const char* const* argv, const char* str, T& val)
{
  char* endptr = nullptr;
  const int index = find_argument(argc, argv, str) + 1;
  errno = 0;
  if (index > 0 && index < argc) {
    const T v = convert(argv[index], &endptr);
    if (errno == ERANGE || *endptr != '\0' || argv[index] == endptr) return -1;  // out of range, junk at the end, nothing converted
    val = v;
  }
  return index - 1;
}
template <typename T>
inline int parse_integral(int argc, const char* const* argv, const char* str, T& val)
{
  long long dummy = 0;
  const int ret = parse_generic([](const char* s, char** e) { return std::strtoll(s, e, 10); }, argc, argv, str, dummy);
  if (ret == -1) return -1;
  const int index = find_argument(argc, argv, str) + 1;
  if (!(index > 0 && index < argc)) return ret;   // option given as the last word: nothing to convert
  if (dummy < static_cast<long long>(std::numeric_limits<T>::min()) || static_cast<unsigned long long>(dummy < 0 ? 0 : dummy) > static_cast<unsigned long long>(std::numeric_limits<T>::max()))
    return -1;
  val = static_cast<T>(dummy);
  return ret;
}
}  // namespace detail
inline int parse_argument(int argc, const char* const* argv, const char* str, long int& val) { return detail::parse_integral(argc, argv, str, val); }
inline int parse_argument(int argc, const char* const* argv, const char* str, long long int& val) { return detail::parse_integral(argc, argv, str, val); }
inline int parse_argument(int argc, const char* const* argv, const char* str, int& val) { return detail::parse_integral(argc, argv, str, val); }
inline int parse_argument(int argc, const char* const* argv, const char* str, unsigned int& val) { return detail::parse_integral(argc, argv, str, val); }
inline int parse_argument(int argc, const char* const* argv, const char* str, unsigned long long int& val)
{
  long long dummy = -1;
  const int ret = parse_argument(argc, argv, str, dummy);
  if (ret == -1 || dummy < 0) return -1;
  val = static_cast<unsigned long long>(dummy);
  return ret;
}
inline int parse_argument(int argc, const char* const* argv, const char* str, double& val)
{
  return detail::parse_generic([](const char* s, char** e) { return std::strtod(s, e); }, argc, argv, str, val);
}
inline int parse_argument(int argc, const char* const* argv, const char* str, float& val)
{
  return detail::parse_generic([](const char* s, char** e) { return std::strtof(s, e); }, argc, argv, str, val);
}
inline int parse_argument(int argc, const char* const* argv, const char* str, bool& val)
{
  long int dummy = 0;
  const int index = find_argument(argc, argv, str) + 1;
  const int ret = parse_argument(argc, argv, str, dummy);
  if (ret != -1 && index > 0 && index < argc) val = dummy != 0;
  return ret;
}
inline int parse_argument(int argc, const char* const* argv, const char* str, char& val)
{
  const int index = find_argument(argc, argv, str) + 1;
  if (index > 0 && index < argc) val = argv[index][0];
  return index - 1;
}
inline int parse(int argc, const char* const* argv, const char* str, std::string& val) { return parse_argument(argc, argv, str, val); }

// indices of the arguments that end in one of the extensions (case-insensitive; a name must be longer than ".ext")
inline std::vector<int> parse_file_extension_argument(int argc, const char* const* argv, const std::vector<std::string>& extensions)
{
  std::vector<int> indices;
  for (int i = 1; i < argc; ++i) {
    std::string fname(argv[i]);
    for (std::string ext : extensions) {
      if (fname.size() <= 4) continue;
      std::transform(fname.begin(), fname.end(), fname.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
      std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
      const std::string::size_type it = fname.rfind(ext);
      if (it != std::string::npos && ext.size() == fname.size() - it) {  // ".p" must not match ".png"
        indices.push_back(i);
        break;
      }
    }
  }
  return indices;
}
inline std::vector<int> parse_file_extension_argument(int argc, const char* const* argv, const std::string& ext)
{
  return parse_file_extension_argument(argc, argv, std::vector<std::string>{ext});
}

namespace detail {
inline std::vector<std::string> split_commas(const char* text)
{
  std::vector<std::string> values;
  std::string cur;
  for (const char* c = text;; ++c) {
    if (*c == ',' || *c == '\0') {
      if (!cur.empty()) values.push_back(cur);   // boost::token_compress_on: empty tokens between commas vanish
      cur.clear();
      if (*c == '\0') break;
    }
    else cur.push_back(*c);
  }
  return values;
}
template <typename T, typename Conv>
inline int parse_tuple(int argc, const char* const* argv, const char* str, std::size_t want, T* out, Conv conv, bool debug, const char* who)
{
  for (int i = 1; i < argc; ++i)
    if (std::strcmp(argv[i], str) == 0 && ++i < argc) {
      const std::vector<std::string> values = split_commas(argv[i]);
      if (values.size() != want) {
        if (debug) print_error("[%s] Number of values for %s (%lu) different than %lu!\n", who, str, static_cast<unsigned long>(values.size()), static_cast<unsigned long>(want));
        return -2;
      }
      for (std::size_t k = 0; k < want; ++k) out[k] = conv(values[k].c_str());
      return i - 1;
    }
  return -1;
}
}  // namespace detail
inline int parse_2x_arguments(int argc, const char* const* argv, const char* str, float& f, float& s, bool debug = true)
{
  float v[2] = {f, s};
  const int r = detail::parse_tuple(argc, argv, str, 2, v, [](const char* t) { return static_cast<float>(std::atof(t)); }, debug, "parse_2x_arguments");
  if (r >= 0) { f = v[0]; s = v[1]; }
  return r;
}
inline int parse_2x_arguments(int argc, const char* const* argv, const char* str, double& f, double& s, bool debug = true)
{
  double v[2] = {f, s};
  const int r = detail::parse_tuple(argc, argv, str, 2, v, [](const char* t) { return std::atof(t); }, debug, "parse_2x_arguments");
  if (r >= 0) { f = v[0]; s = v[1]; }
  return r;
}
inline int parse_2x_arguments(int argc, const char* const* argv, const char* str, int& f, int& s, bool debug = true)
{
  int v[2] = {f, s};
  const int r = detail::parse_tuple(argc, argv, str, 2, v, [](const char* t) { return std::atoi(t); }, debug, "parse_2x_arguments");
  if (r >= 0) { f = v[0]; s = v[1]; }
  return r;
}
inline int parse_3x_arguments(int argc, const char* const* argv, const char* str, float& f, float& s, float& t, bool debug = true)
{
  float v[3] = {f, s, t};
  const int r = detail::parse_tuple(argc, argv, str, 3, v, [](const char* x) { return static_cast<float>(std::atof(x)); }, debug, "parse_3x_arguments");
  if (r >= 0) { f = v[0]; s = v[1]; t = v[2]; }
  return r;
}
inline int parse_3x_arguments(int argc, const char* const* argv, const char* str, double& f, double& s, double& t, bool debug = true)
{
  double v[3] = {f, s, t};
  const int r = detail::parse_tuple(argc, argv, str, 3, v, [](const char* x) { return std::atof(x); }, debug, "parse_3x_arguments");
  if (r >= 0) { f = v[0]; s = v[1]; t = v[2]; }
  return r;
}
inline int parse_3x_arguments(int argc, const char* const* argv, const char* str, int& f, int& s, int& t, bool debug = true)
{
  int v[3] = {f, s, t};
  const int r = detail::parse_tuple(argc, argv, str, 3, v, [](const char* x) { return std::atoi(x); }, debug, "parse_3x_arguments");
  if (r >= 0) { f = v[0]; s = v[1]; t = v[2]; }
  return r;
}
inline int parse_x_arguments(int argc, const char* const* argv, const char* str, std::vector<double>& v)
{
  for (int i = 1; i < argc; ++i)
    if (std::strcmp(argv[i], str) == 0 && ++i < argc) {
      v.clear();
      for (const std::string& t : detail::split_commas(argv[i])) v.push_back(std::atof(t.c_str()));
      return i - 1;
    }
  return -1;
}
inline int parse_x_arguments(int argc, const char* const* argv, const char* str, std::vector<float>& v)
{
  std::vector<double> d;
  const int r = parse_x_arguments(argc, argv, str, d);
  if (r >= 0) v.assign(d.begin(), d.end());
  return r;
}
inline int parse_x_arguments(int argc, const char* const* argv, const char* str, std::vector<int>& v)
{
  for (int i = 1; i < argc; ++i)
    if (std::strcmp(argv[i], str) == 0 && ++i < argc) {
      v.clear();
      for (const std::string& t : detail::split_commas(argv[i])) v.push_back(std::atoi(t.c_str()));
      return i - 1;
    }
  return -1;
}
// every occurrence of an option that may be given several times
template <typename T>
inline bool parse_multiple_arguments(int argc, const char* const* argv, const char* str, std::vector<T>& values)
{
  for (int i = 1; i < argc; ++i)
    if (std::strcmp(argv[i], str) == 0 && ++i < argc) {
      const char* two[3] = {argv[0], str, argv[i]};
      T v{};
      if (parse_argument(3, two, str, v) != -1) values.push_back(v);
    }
  return !values.empty();
}
}  // namespace console
}  // namespace pcl
