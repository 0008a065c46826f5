// oracle/shim/tf2_ros/static_transform_broadcaster.h — TEST INFRASTRUCTURE ONLY.
#include "tf2_ros/tf2_common.h"
