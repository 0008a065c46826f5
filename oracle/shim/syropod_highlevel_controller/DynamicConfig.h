// oracle/shim/syropod_highlevel_controller/DynamicConfig.h — TEST INFRASTRUCTURE ONLY.
#include "syropod_highlevel_controller/msgs_generated.h"
