// TEST INFRASTRUCTURE — agenttorch_b200/csrc/at_interventions.cu on the SIMT shim (one OS thread per CUDA thread, running
// concurrently), for the ThreadSanitizer run of tests/test_racecheck.py.
#define AT_HOST_EMULATION 1
#include "simt_shim.h"
#include "../../agenttorch_b200/csrc/at_interventions.cu"
