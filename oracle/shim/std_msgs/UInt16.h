// oracle/shim/std_msgs/UInt16.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
