// oracle/shim/visualization_msgs/MarkerArray.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
