// oracle/shim/ros/exceptions.h — TEST INFRASTRUCTURE ONLY (see ros.h).
#ifndef SHC_SHIM_ROS_EXCEPTIONS_H
#define SHC_SHIM_ROS_EXCEPTIONS_H
#include <stdexcept>
namespace ros { class Exception : public std::runtime_error { public: Exception(const std::string& w) : std::runtime_error(w) {} }; }
#endif
