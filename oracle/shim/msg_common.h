// oracle/shim/msg_common.h — TEST INFRASTRUCTURE ONLY: plain structs with the field layout of the ROS message types the
// reference fills (std_msgs, geometry_msgs, sensor_msgs, visualization_msgs: public message definitions).
#ifndef SHC_SHIM_MSG_COMMON_H
#define SHC_SHIM_MSG_COMMON_H
#include <cstdint>
#include <string>
#include <vector>
#include "ros/ros.h"

namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
struct Bool { bool data = false; };
struct Int8 { int8_t data = 0; };
struct UInt16 { uint16_t data = 0; };
struct Float64 { double data = 0.0; };
struct MultiArrayDimension { std::string label; uint32_t size = 0; uint32_t stride = 0; };
struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; uint32_t data_offset = 0; };
struct Float32MultiArray { MultiArrayLayout layout; std::vector<float> data; };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}  // namespace std_msgs

namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct Twist { Vector3 linear; Vector3 angular; };
struct TwistStamped { std_msgs::Header header; Twist twist; };
struct Wrench { Vector3 force; Vector3 torque; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::Header header; std::string child_frame_id; Transform transform; };
}  // namespace geometry_msgs

namespace sensor_msgs {
struct JointState {
  std_msgs::Header header;
  std::vector<std::string> name;
  std::vector<double> position, velocity, effort;
};
struct Imu {
  std_msgs::Header header;
  geometry_msgs::Quaternion orientation;
  double orientation_covariance[9] = {0};
  geometry_msgs::Vector3 angular_velocity;
  double angular_velocity_covariance[9] = {0};
  geometry_msgs::Vector3 linear_acceleration;
  double linear_acceleration_covariance[9] = {0};
};
struct Joy { std_msgs::Header header; std::vector<float> axes; std::vector<int32_t> buttons; };
}  // namespace sensor_msgs

namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7,
         POINTS = 8, TEXT_VIEW_FACING = 9, MESH_RESOURCE = 10, TRIANGLE_LIST = 11 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
  std_msgs::Header header;
  std::string ns;
  int32_t id = 0;
  int32_t type = 0;
  int32_t action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  ros::Duration lifetime;
  bool frame_locked = false;
  std::vector<geometry_msgs::Point> points;
  std::vector<std_msgs::ColorRGBA> colors;
  std::string text;
  std::string mesh_resource;
  bool mesh_use_embedded_materials = false;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs
#endif
