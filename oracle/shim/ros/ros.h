// oracle/shim/ros/ros.h — TEST INFRASTRUCTURE ONLY.
//
// A self-written, in-process stand-in for the handful of roscpp facilities the reference touches (ROS is absent from this
// image), so that the reference's own sources compile unmodified into oracle/_ref/ (oracle/Makefile.ref):
//   * NodeHandle::getParam reads a process-wide parameter table the harness fills (shc_shim::params());
//   * Publisher::publish keeps the last message of every topic (shc_shim::bus()) so that the harness can read what the
//     reference's publishers produced; subscribe() only records that it happened — the harness invokes the reference's
//     callbacks directly, as ros::spinOnce() would;
//   * ros::Time::now() is a harness-controlled clock; console macros are silent; ROS_ASSERT counts violations.
#ifndef SHC_SHIM_ROS_H
#define SHC_SHIM_ROS_H

#include <climits>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace shc_shim {

struct ParamTable {
  std::map<std::string, double> d;
  std::map<std::string, bool> b;
  std::map<std::string, int> i;
  std::map<std::string, std::string> s;
  std::map<std::string, std::vector<std::string>> vs;
  std::map<std::string, std::vector<int>> vi;
  std::map<std::string, std::vector<double>> vd;
  std::map<std::string, std::map<std::string, double>> md;
  std::map<std::string, std::map<std::string, int>> mi;
  void clear() { *this = ParamTable(); }
};
inline ParamTable& params() { static ParamTable t; return t; }

struct Bus {
  std::map<std::string, std::shared_ptr<void>> last;  // topic -> last published message
  std::map<std::string, long> count;
};
inline Bus& bus() { static Bus b; return b; }

struct Runtime {
  double now = 0.0;
  long assert_failures = 0;
  std::string first_assert;
  bool shutdown_requested = false;
};
inline Runtime& runtime() { static Runtime r; return r; }

inline void assertion(bool ok, const char* expr, const char* file, int line) {
  if (ok) return;
  Runtime& r = runtime();
  if (r.assert_failures++ == 0) r.first_assert = std::string(file) + ":" + std::to_string(line) + ": " + expr;
}

template <class T>
inline bool lookup(const std::map<std::string, T>& m, const std::string& k, T& out) {
  auto it = m.find(k);
  if (it == m.end()) return false;
  out = it->second;
  return true;
}
}  // namespace shc_shim

namespace ros {

struct Duration {
  double sec_;
  Duration(double s = 0.0) : sec_(s) {}
  double toSec() const { return sec_; }
  bool sleep() const { return true; }
};
struct Time {
  double sec_;
  Time() : sec_(0.0) {}
  Time(double s) : sec_(s) {}
  static Time now() { return Time(shc_shim::runtime().now); }
  double toSec() const { return sec_; }
  bool isZero() const { return sec_ == 0.0; }
  Time operator-(const Duration& d) const { return Time(sec_ - d.sec_); }
  Time operator+(const Duration& d) const { return Time(sec_ + d.sec_); }
  Duration operator-(const Time& o) const { return Duration(sec_ - o.sec_); }
  bool operator<(const Time& o) const { return sec_ < o.sec_; }
  bool operator>(const Time& o) const { return sec_ > o.sec_; }
  bool operator==(const Time& o) const { return sec_ == o.sec_; }
  bool operator!=(const Time& o) const { return sec_ != o.sec_; }
};
struct Rate {
  Rate(double) {}
  bool sleep() { return true; }
};

inline void init(int&, char**, const std::string&) {}
inline void spinOnce() {}
inline bool ok() { return !shc_shim::runtime().shutdown_requested; }
inline void shutdown() { shc_shim::runtime().shutdown_requested = true; }

class Publisher {
 public:
  Publisher() {}
  explicit Publisher(const std::string& topic) : topic_(topic) {}
  template <class M>
  void publish(const M& m) const {
    shc_shim::Bus& b = shc_shim::bus();
    b.last[topic_] = std::make_shared<M>(m);
    ++b.count[topic_];
  }
  std::string getTopic() const { return topic_; }
 private:
  std::string topic_;
};
class Subscriber {};

class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <class M, class T>
  Subscriber subscribe(const std::string&, uint32_t, void (T::*)(const M&), T*) { return Subscriber(); }
  template <class M>
  Publisher advertise(const std::string& topic, uint32_t, bool = false) { return Publisher(topic); }

  bool getParam(const std::string& k, double& v) const {
    if (shc_shim::lookup(shc_shim::params().d, k, v)) return true;
    int iv;
    if (shc_shim::lookup(shc_shim::params().i, k, iv)) { v = iv; return true; }
    return false;
  }
  bool getParam(const std::string& k, bool& v) const { return shc_shim::lookup(shc_shim::params().b, k, v); }
  bool getParam(const std::string& k, int& v) const { return shc_shim::lookup(shc_shim::params().i, k, v); }
  bool getParam(const std::string& k, std::string& v) const { return shc_shim::lookup(shc_shim::params().s, k, v); }
  bool getParam(const std::string& k, std::vector<std::string>& v) const { return shc_shim::lookup(shc_shim::params().vs, k, v); }
  bool getParam(const std::string& k, std::vector<int>& v) const { return shc_shim::lookup(shc_shim::params().vi, k, v); }
  bool getParam(const std::string& k, std::vector<double>& v) const { return shc_shim::lookup(shc_shim::params().vd, k, v); }
  bool getParam(const std::string& k, std::map<std::string, double>& v) const { return shc_shim::lookup(shc_shim::params().md, k, v); }
  bool getParam(const std::string& k, std::map<std::string, int>& v) const { return shc_shim::lookup(shc_shim::params().mi, k, v); }
  template <class T>
  bool param(const std::string& k, T& v, const T& dflt) const {
    if (getParam(k, v)) return true;
    v = dflt;
    return false;
  }
  bool hasParam(const std::string& k) const {
    const shc_shim::ParamTable& p = shc_shim::params();
    return p.d.count(k) || p.b.count(k) || p.i.count(k) || p.s.count(k) || p.vs.count(k) || p.vi.count(k) || p.vd.count(k) ||
           p.md.count(k) || p.mi.count(k);
  }
  template <class T>
  void setParam(const std::string&, const T&) const {}
};

}  // namespace ros

#include "ros/console.h"
#include "ros/assert.h"
#endif
