// TEST INFRASTRUCTURE ONLY.  Stand-in for <gflags/gflags.h>: flags are plain globals holding their defaults.
#ifndef VSO_REF_SHIM_GFLAGS_H_
#define VSO_REF_SHIM_GFLAGS_H_
#include <string>
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DEFINE_bool(name, def, doc) bool FLAGS_##name = (def)
#define DECLARE_int32(name) extern int FLAGS_##name
#define DEFINE_int32(name, def, doc) int FLAGS_##name = (def)
#define DECLARE_double(name) extern double FLAGS_##name
#define DEFINE_double(name, def, doc) double FLAGS_##name = (def)
#define DECLARE_string(name) extern std::string FLAGS_##name
#define DEFINE_string(name, def, doc) std::string FLAGS_##name = (def)
#endif
