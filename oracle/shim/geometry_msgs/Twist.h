// oracle/shim/geometry_msgs/Twist.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
