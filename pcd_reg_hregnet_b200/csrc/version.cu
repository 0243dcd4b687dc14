#include "common.cuh"
HRN_API const char* hrn_version(void) { return "hregnet_b200 0.1.0 sm_100a"; }
