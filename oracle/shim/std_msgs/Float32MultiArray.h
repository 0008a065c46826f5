// oracle/shim/std_msgs/Float32MultiArray.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
