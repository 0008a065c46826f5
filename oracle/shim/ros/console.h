// oracle/shim/ros/console.h — TEST INFRASTRUCTURE ONLY (see ros.h): rosconsole macros, silent.
#ifndef SHC_SHIM_ROS_CONSOLE_H
#define SHC_SHIM_ROS_CONSOLE_H
#include <string>
#define ROSCONSOLE_DEFAULT_NAME "ros.shc"
namespace ros { namespace console {
namespace levels { enum Level { Debug, Info, Warn, Error, Fatal }; }
inline bool set_logger_level(const std::string&, levels::Level) { return true; }
inline void notifyLoggerLevelsChanged() {}
}}
#define SHC_SHIM_NOLOG(...) do { } while (0)
#define ROS_DEBUG(...) SHC_SHIM_NOLOG()
#define ROS_INFO(...) SHC_SHIM_NOLOG()
#define ROS_WARN(...) SHC_SHIM_NOLOG()
#define ROS_ERROR(...) SHC_SHIM_NOLOG()
#define ROS_FATAL(...) SHC_SHIM_NOLOG()
#define ROS_DEBUG_ONCE(...) SHC_SHIM_NOLOG()
#define ROS_INFO_ONCE(...) SHC_SHIM_NOLOG()
#define ROS_WARN_ONCE(...) SHC_SHIM_NOLOG()
#define ROS_ERROR_ONCE(...) SHC_SHIM_NOLOG()
#define ROS_DEBUG_COND(c, ...) do { (void)(c); } while (0)
#define ROS_INFO_COND(c, ...) do { (void)(c); } while (0)
#define ROS_WARN_COND(c, ...) do { (void)(c); } while (0)
#define ROS_ERROR_COND(c, ...) do { (void)(c); } while (0)
#define ROS_FATAL_COND(c, ...) do { (void)(c); } while (0)
#define ROS_DEBUG_THROTTLE(p, ...) do { (void)(p); } while (0)
#define ROS_INFO_THROTTLE(p, ...) do { (void)(p); } while (0)
#define ROS_WARN_THROTTLE(p, ...) do { (void)(p); } while (0)
#define ROS_ERROR_THROTTLE(p, ...) do { (void)(p); } while (0)
#define ROS_FATAL_THROTTLE(p, ...) do { (void)(p); } while (0)
#endif
