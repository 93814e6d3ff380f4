// Minimal stand-in for glog so the reference's translation units compile unmodified (oracle/_ref).
// LOG(FATAL)/failed CHECK abort like glog; everything else is swallowed.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <gflags/gflags.h>

namespace xw_shim {
struct LogSink {
    bool fatal;
    std::ostringstream os;
    explicit LogSink(bool f) : fatal(f) {}
    ~LogSink() { if (fatal) { std::cerr << "FATAL: " << os.str() << std::endl; std::abort(); } }
    template <typename T> LogSink& operator<<(const T&) { return *this; }  // message text is dropped
    LogSink& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct Voidify { void operator&(LogSink&) {} };
}  // namespace xw_shim
#define XW_LOG_INFO xw_shim::LogSink(false)
#define XW_LOG_WARNING xw_shim::LogSink(false)
#define XW_LOG_ERROR xw_shim::LogSink(false)
#define XW_LOG_FATAL xw_shim::LogSink(true)
#define LOG(sev) XW_LOG_##sev
#define VLOG(n) xw_shim::LogSink(false)
#define CHECK(c) (c) ? (void)0 : xw_shim::Voidify() & xw_shim::LogSink(true) << "Check failed: " #c " "
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
