// oracle/shim/syropod_highlevel_controller/TipState.h — TEST INFRASTRUCTURE ONLY.
#include "syropod_highlevel_controller/msgs_generated.h"
