// TEST INFRASTRUCTURE — agenttorch_b200/csrc/at_interventions.cu (kernels AND C-ABI wrappers) compiled for the host.  The
// library exports the at_iv_* entry points of libat_interventions.so (host pointers; used to drive the torch ops of
// agenttorch_b200/interventions.py on a machine without a GPU) plus emu_* entry points that run ONE kernel with a chosen
// grid and vector/scalar path (tests/test_interventions_emulated.py).  Never linked into the product.
#define AT_HOST_EMULATION 1
#include "cuda_shim.h"
#include "../../agenttorch_b200/csrc/at_interventions.cu"

extern "C" {
void emu_quarantine(const at_iv_quarantine_args* a, int vec, int blocks) { AT_LAUNCH(iv_quarantine_kernel, blocks, 1, nullptr, *a, vec != 0); }
void emu_testing(const at_iv_testing_args* a, int vec, int blocks) { AT_LAUNCH(iv_testing_kernel, blocks, 1, nullptr, *a, vec != 0); }
void emu_seirms(const at_iv_seirms_args* a, int vec, int blocks)
{
    AT_LAUNCH(iv_seirms_kernel, blocks, 1, nullptr, *a, vec != 0);
    AT_LAUNCH(iv_seirms_epilogue_kernel, 1, 1, nullptr, *a);
}
}
