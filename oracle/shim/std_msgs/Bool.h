// oracle/shim/std_msgs/Bool.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
