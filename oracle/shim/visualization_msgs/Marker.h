// oracle/shim/visualization_msgs/Marker.h — TEST INFRASTRUCTURE ONLY (see msg_common.h).
#include "msg_common.h"
