// TEST INFRASTRUCTURE ONLY: minimal stand-in for glog so that the reference's own C++ sources
// (runtime/core/speaker/speaker_engine.cc, frontend/feature_pipeline.cc, frontend/fft.cc, frontend/fbank.h) compile from
// where they lie under /root/reference without the FetchContent dependencies of its CMake build.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace ws_glog_shim {
struct Sink {
    bool fatal;
    std::ostringstream os;
    explicit Sink(bool f) : fatal(f) {}
    ~Sink() {
        if (fatal) { std::cerr << os.str() << std::endl; std::abort(); }
    }
    template <class T>
    Sink& operator<<(const T& v) { os << v; return *this; }
};
struct Voidify { void operator&(Sink&) {} };
}  // namespace ws_glog_shim

#define LOG(sev) ws_glog_shim::Sink(false)
#define VLOG(n) ws_glog_shim::Sink(false)
#define CHECK(c) (c) ? (void)0 : ws_glog_shim::Voidify() & ws_glog_shim::Sink(true) << "CHECK failed: " #c " "
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
