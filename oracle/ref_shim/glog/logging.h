// Shim for <glog/logging.h> so that the reference's own src/helpers.cpp compiles from where it lies under
// /root/reference without building its vendored glog (a cmake project with generated headers).  Logging goes
// nowhere; CHECK* abort on failure like glog's.  Test infrastructure (oracle/_ref), not product code.
#pragma once
#include <cstdlib>
#include <iostream>
namespace ref_shim {
struct NullStream {
  template <typename T>
  NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
  explicit operator bool() const { return true; }  // `LOG(INFO) << a << b ? x : y;` appears in the reference (std::ostream converts too)
};
struct FatalStream {  // LOG(FATAL) / failed CHECK: print the message, then abort like glog
  const char* file;
  int line;
  FatalStream(const char* f = "", int l = 0) : file(f), line(l) { std::cerr << "[reference FATAL] " << file << ":" << line << " "; }
  template <typename T>
  FatalStream& operator<<(const T& v) { std::cerr << v; return *this; }
  FatalStream& operator<<(std::ostream& (*f)(std::ostream&)) { std::cerr << f; return *this; }
  explicit operator bool() const { return true; }
  ~FatalStream() { std::cerr << std::endl; std::abort(); }
};
struct LogStream_INFO : NullStream { LogStream_INFO(const char*, int) {} };
struct LogStream_WARNING : NullStream { LogStream_WARNING(const char*, int) {} };
struct LogStream_ERROR : NullStream { LogStream_ERROR(const char*, int) {} };
using LogStream_FATAL = FatalStream;
}  // namespace ref_shim
#define LOG(severity) ref_shim::LogStream_##severity(__FILE__, __LINE__)
#define VLOG(n) ref_shim::NullStream()
#define LOG_IF(severity, cond) ref_shim::NullStream()
#define CHECK(cond) if (cond) {} else ref_shim::FatalStream(__FILE__, __LINE__) << "CHECK failed: " #cond " "
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_NOTNULL(p) (p)
#define VLOG_IF(n, cond) ref_shim::NullStream()
#define DLOG(severity) ref_shim::NullStream()
#define DVLOG(n) ref_shim::NullStream()
#define LOG_EVERY_N(severity, n) ref_shim::NullStream()
#define LOG_FIRST_N(severity, n) ref_shim::NullStream()
#define VLOG_IS_ON(n) false
#define DCHECK(cond) CHECK(cond)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
namespace google {
inline void InitGoogleLogging(const char*) {}
}
