// oracle/shim/dynamic_reconfigure/server.h — TEST INFRASTRUCTURE ONLY: a dynamic_reconfigure::Server that stores its
// callback and configs; plus the little of Boost the reference names next to it (bind, recursive_mutex).
#ifndef SHC_SHIM_DYNREC_SERVER_H
#define SHC_SHIM_DYNREC_SERVER_H
#include <functional>
#include <cstdint>
namespace boost {
struct recursive_mutex {};
template <class... A>
auto bind(A&&... a) -> decltype(std::bind(std::forward<A>(a)...)) { return std::bind(std::forward<A>(a)...); }
}
using std::placeholders::_1;
using std::placeholders::_2;
namespace dynamic_reconfigure {
template <class Config>
class Server {
 public:
  typedef std::function<void(Config&, uint32_t)> CallbackType;
  explicit Server(boost::recursive_mutex&) {}
  void setCallback(const CallbackType& cb) { cb_ = cb; }
  void setConfigMax(const Config& c) { max_ = c; }
  void setConfigMin(const Config& c) { min_ = c; }
  void setConfigDefault(const Config& c) { default_ = c; }
  void updateConfig(const Config& c) { config_ = c; }
  CallbackType cb_;
  Config max_, min_, default_, config_;
};
}  // namespace dynamic_reconfigure
#endif
