// TEST INFRASTRUCTURE — agenttorch_b200/csrc/at_predprey.cu on the SIMT shim (one OS thread per CUDA thread, running
// concurrently): used under ThreadSanitizer by tests/test_racecheck.py to confirm that the threads of a predator-prey
// launch only interact through order-independent atomics — the property the sequential shim (cuda_shim.h) relies on.
#define AT_HOST_EMULATION 1
#include "simt_shim.h"
struct longlong2 { long long x, y; };
static inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }
#include "../../agenttorch_b200/csrc/at_predprey.cu"
