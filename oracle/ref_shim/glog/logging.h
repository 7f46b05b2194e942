// TEST INFRASTRUCTURE ONLY.  Minimal stand-in for <glog/logging.h> (glog is not installed here) so that the few
// reference translation units that need nothing else -- segmentation/histograms.cpp, base/base.cpp -- compile
// UNMODIFIED from /root/reference into oracle/_ref (see oracle/Makefile, target _ref).  CHECKs abort with a
// message, LOG / DLOG / VLOG swallow their stream.
#ifndef VSO_REF_SHIM_GLOG_LOGGING_H_
#define VSO_REF_SHIM_GLOG_LOGGING_H_
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace vso_shim {
struct NullStream {
  template <class T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  std::ostringstream s;
  FatalStream(const char* file, int line, const char* what) { s << file << ":" << line << " check failed: " << what << " "; }
  template <class T> FatalStream& operator<<(const T& v) { s << v; return *this; }
  ~FatalStream() { std::cerr << s.str() << std::endl; std::abort(); }
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
