// oracle/shim/std_msgs/Float64.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
