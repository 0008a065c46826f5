// oracle/shim/sensor_msgs/Imu.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
