// oracle/shim/syropod_highlevel_controller/TargetTipPose.h — TEST INFRASTRUCTURE ONLY.
#include "syropod_highlevel_controller/msgs_generated.h"
