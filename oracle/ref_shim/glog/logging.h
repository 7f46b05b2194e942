// TEST INFRASTRUCTURE ONLY.  Minimal stand-in for <glog/logging.h> (glog is not installed here) so that the few
// reference translation units that need nothing else -- segmentation/histograms.cpp, base/base.cpp -- compile
// UNMODIFIED from /root/reference into oracle/_ref (see oracle/Makefile, target _ref).  CHECKs abort with a
// message, LOG / DLOG / VLOG swallow their stream.
#ifndef VSO_REF_SHIM_GLOG_LOGGING_H_
#define VSO_REF_SHIM_GLOG_LOGGING_H_
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace vso_shim {
struct NullStream {
  template <class T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
// Formats with snprintf, not iostreams: the image's g++ links libstdc++ statically into every shared object, and the
// locale facets of a second copy are not reliably initialised in a process that already loaded another one (numpy
// loads the shared libstdc++ first), so `ostream << int` can crash exactly when a CHECK wants to report.
struct FatalStream {
  std::string s;
  FatalStream(const char* file, int line, const char* what) {
    char b[64];
    snprintf(b, sizeof(b), ":%d check failed: ", line);
    s = std::string(file) + b + what + " ";
  }
  FatalStream& operator<<(const char* v) { s += v ? v : "(null)"; return *this; }
  FatalStream& operator<<(const std::string& v) { s += v; return *this; }
  FatalStream& operator<<(char v) { s += v; return *this; }
  FatalStream& operator<<(bool v) { s += v ? "true" : "false"; return *this; }
  FatalStream& operator<<(double v) { char b[64]; snprintf(b, sizeof(b), "%g", v); s += b; return *this; }
  FatalStream& operator<<(float v) { return *this << (double)v; }
  FatalStream& operator<<(long long v) { char b[32]; snprintf(b, sizeof(b), "%lld", v); s += b; return *this; }
  FatalStream& operator<<(unsigned long long v) { char b[32]; snprintf(b, sizeof(b), "%llu", v); s += b; return *this; }
  FatalStream& operator<<(int v) { return *this << (long long)v; }
  FatalStream& operator<<(long v) { return *this << (long long)v; }
  FatalStream& operator<<(unsigned v) { return *this << (unsigned long long)v; }
  FatalStream& operator<<(unsigned long v) { return *this << (unsigned long long)v; }
  FatalStream& operator<<(const void* v) { char b[32]; snprintf(b, sizeof(b), "%p", v); s += b; return *this; }
  // anything else (cv::Size ...): through a local stream
  template <class T> FatalStream& operator<<(const T& v) { std::ostringstream o; o << v; s += o.str(); return *this; }
  ~FatalStream() { fputs(s.c_str(), stderr); fputc('\n', stderr); fflush(stderr); std::abort(); }
};
struct Voidify { void operator&(const NullStream&) {} void operator&(const FatalStream&) {} };
}  // namespace vso_shim

#define VSO_SHIM_CHECK(cond, text) (cond) ? (void)0 : ::vso_shim::Voidify() & ::vso_shim::FatalStream(__FILE__, __LINE__, text)
#define CHECK(c) VSO_SHIM_CHECK((c), #c)
#define CHECK_EQ(a, b) VSO_SHIM_CHECK((a) == (b), #a " == " #b)
#define CHECK_NE(a, b) VSO_SHIM_CHECK((a) != (b), #a " != " #b)
#define CHECK_LT(a, b) VSO_SHIM_CHECK((a) < (b), #a " < " #b)
#define CHECK_LE(a, b) VSO_SHIM_CHECK((a) <= (b), #a " <= " #b)
#define CHECK_GT(a, b) VSO_SHIM_CHECK((a) > (b), #a " > " #b)
#define CHECK_GE(a, b) VSO_SHIM_CHECK((a) >= (b), #a " >= " #b)
#define CHECK_NOTNULL(p) (p)
// debug checks are compiled out, as in the reference's Release build (NDEBUG)
#define DCHECK(c) while (false) CHECK(c)
#define DCHECK_EQ(a, b) while (false) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) while (false) CHECK_NE(a, b)
#define DCHECK_LT(a, b) while (false) CHECK_LT(a, b)
#define DCHECK_LE(a, b) while (false) CHECK_LE(a, b)
#define DCHECK_GT(a, b) while (false) CHECK_GT(a, b)
#define DCHECK_GE(a, b) while (false) CHECK_GE(a, b)
#define LOG(level) ::vso_shim::NullStream()
#define DLOG(level) ::vso_shim::NullStream()
#define VLOG(level) ::vso_shim::NullStream()
#define LOG_IF(level, c) ::vso_shim::NullStream()
#endif
