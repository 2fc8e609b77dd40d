// TEST INFRASTRUCTURE — agenttorch_b200/csrc/at_predprey.cu (kernels AND its C-ABI wrappers) compiled for the host: the
// resulting library exports the same at_pp_* entry points as libat_predprey.so and works on host pointers, so the Python
// engine can be driven against it on a machine without a GPU (tests/test_predprey_emulated.py).
#define AT_HOST_EMULATION 1
#include "cuda_shim.h"
#include "../../agenttorch_b200/csrc/at_predprey.cu"
