// TEST INFRASTRUCTURE — agenttorch_b200/csrc/at_seirm.cu (kernels AND C-ABI wrappers) compiled for the host on the SIMT
// shim: the library exports the at_seirm_* entry points of libat_seirm.so and works on host pointers, so the Python engine
// can be driven against it on a machine without a GPU (tests/test_seirm_emulated.py).  Never linked into the product.
#define AT_HOST_EMULATION 1
#include "simt_shim.h"
#include "../../agenttorch_b200/csrc/at_seirm.cu"

extern "C" void emu_set_sm_count(int n) { g_emu_sm_count = n > 0 ? n : 1; g_sm_count = 0; }
