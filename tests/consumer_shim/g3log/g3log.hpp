// stand-in for <g3log/g3log.hpp>: LOG(level) << ... and CHECK(cond) << ... as stream sinks
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace g3shim {
struct Sink {
  bool fatal;
  std::ostringstream os;
  explicit Sink(bool f) : fatal(f) {}
  ~Sink() { if (fatal) { std::cerr << os.str() << std::endl; std::abort(); } }
  template <typename T> Sink &operator<<(const T &v) { os << v; return *this; }
  Sink &operator<<(std::ostream &(*f)(std::ostream &)) { os << f; return *this; }
  Sink &operator<<(std::ios_base &(*f)(std::ios_base &)) { os << f; return *this; }
};
}  // namespace g3shim
#define LOG(level) g3shim::Sink(false)
#define CHECK(cond) if (cond) {} else g3shim::Sink(true)
