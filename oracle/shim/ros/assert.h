// oracle/shim/ros/assert.h — TEST INFRASTRUCTURE ONLY (see ros.h): ROS_ASSERT records a violation and carries on, so that
// the harness can report how many of the reference's own assertions a run tripped.
#ifndef SHC_SHIM_ROS_ASSERT_H
#define SHC_SHIM_ROS_ASSERT_H
#define ROS_ASSERT(cond) ::shc_shim::assertion(!!(cond), #cond, __FILE__, __LINE__)
#define ROS_ASSERT_MSG(cond, ...) ::shc_shim::assertion(!!(cond), #cond, __FILE__, __LINE__)
#define ROS_BREAK() ::shc_shim::assertion(false, "ROS_BREAK", __FILE__, __LINE__)
#endif
