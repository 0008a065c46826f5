// oracle/shim/std_msgs/Int8.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
