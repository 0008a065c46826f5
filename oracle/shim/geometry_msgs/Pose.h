// oracle/shim/geometry_msgs/Pose.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
