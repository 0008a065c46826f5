// oracle/shim/tf2_ros/tf2_common.h — TEST INFRASTRUCTURE ONLY: tf2 stand-in.  Broadcasters keep the last transform per
// child frame; Buffer::lookupTransform answers from a table the harness fills (shc_shim::tf_table()), keyed
// "target<-source" (the time-travel form first tries "target@<target time><-source", so that requests stamped at
// different times can see different robot movements), and throws tf2::TransformException when the harness has not
// provided one — the reference catches it.
#ifndef SHC_SHIM_TF2_COMMON_H
#define SHC_SHIM_TF2_COMMON_H
#include <map>
#include <stdexcept>
#include "msg_common.h"
namespace shc_shim {
inline std::map<std::string, geometry_msgs::TransformStamped>& tf_table() { static std::map<std::string, geometry_msgs::TransformStamped> t; return t; }
inline std::map<std::string, geometry_msgs::TransformStamped>& tf_sent() { static std::map<std::string, geometry_msgs::TransformStamped> t; return t; }
}
namespace tf2 {
class TransformException : public std::runtime_error { public: TransformException(const std::string& w) : std::runtime_error(w) {} };
}
namespace tf2_ros {
class Buffer {
 public:
  geometry_msgs::TransformStamped lookupTransform(const std::string& target, const std::string& source, const ros::Time&,
                                                  const ros::Duration = ros::Duration(0.0)) const {
    auto it = shc_shim::tf_table().find(target + "<-" + source);
    if (it == shc_shim::tf_table().end()) throw tf2::TransformException("no transform " + target + "<-" + source);
    return it->second;
  }
  geometry_msgs::TransformStamped lookupTransform(const std::string& target, const ros::Time& target_time, const std::string& source,
                                                  const ros::Time&, const std::string&,
                                                  const ros::Duration = ros::Duration(0.0)) const {
    auto it = shc_shim::tf_table().find(target + "@" + std::to_string(target_time.toSec()) + "<-" + source);
    if (it != shc_shim::tf_table().end()) return it->second;
    return lookupTransform(target, source, ros::Time(0));
  }
};
class TransformListener { public: explicit TransformListener(Buffer&) {} };
class TransformBroadcaster {
 public:
  void sendTransform(const geometry_msgs::TransformStamped& t) { shc_shim::tf_sent()[t.child_frame_id] = t; }
  void sendTransform(const std::vector<geometry_msgs::TransformStamped>& v) { for (const auto& t : v) sendTransform(t); }
};
class StaticTransformBroadcaster : public TransformBroadcaster {};
}  // namespace tf2_ros
#endif
