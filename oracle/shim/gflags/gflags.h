// Minimal stand-in for gflags: flags are plain globals FLAGS_x.
#pragma once
#define GFLAGS_GFLAGS_H_
#include <cstdint>
#include <string>
namespace gflags {}
#define DEFINE_int32(n, v, h) int32_t FLAGS_##n = (v)
#define DEFINE_bool(n, v, h) bool FLAGS_##n = (v)
#define DEFINE_double(n, v, h) double FLAGS_##n = (v)
#define DEFINE_string(n, v, h) std::string FLAGS_##n = (v)
#define DECLARE_int32(n) extern int32_t FLAGS_##n
#define DECLARE_bool(n) extern bool FLAGS_##n
#define DECLARE_double(n) extern double FLAGS_##n
#define DECLARE_string(n) extern std::string FLAGS_##n
