// Circuit-specialised lane kernels: the lane program of a plan (vm_ops.h) written out as straight-line CUDA, compiled once per
// plan by NVRTC for sm_100a and launched instead of the interpreter (vm_kernel.cuh).  Scratch cells become registers, the op
// fetch / decode / dispatch disappears; the arithmetic is the interpreter's, operation by operation (vm_jit_prelude.inc).
#pragma once
#include <cstddef>
#include <string>

#include "plan.hpp"

namespace krb {

struct JitKernel;   // a loaded module + its entry point on one device

// programs above this many operations stay on the interpreter (compile time, instruction-cache footprint); KRB_VM_JIT_MAX_OPS overrides
int64_t vm_jit_max_ops();
// KRB_VM_JIT=0 switches the specialised kernels off (A/B measurements, tests of the interpreter)
bool vm_jit_enabled();
bool vm_jit_eligible(const Plan &plan);
// NVRTC needs about half a second per thousand operations: programs above KRB_VM_JIT_SMALL_OPS are specialised only when the plan's
// first solve has enough points (lanes x rows) to pay for it
bool vm_jit_worthwhile(const Plan &plan, long long points);
// CUDA source of the plan's kernel; host_mode: the same text behind KRB_JIT_HOST (tests/jit_host_shim.h) for the CPU tests
std::string vm_jit_source(const Plan &plan, size_t vm_params_size, bool host_mode);
// NVRTC only (no device needed): cubin for sm_100a; false + err when NVRTC is missing or rejects the source
bool vm_jit_compile(const std::string &src, std::string &cubin, std::string &err);
// compile + load into the current context; nullptr + err on failure
JitKernel *vm_jit_build(const Plan &plan, size_t vm_params_size, std::string &err);
// params: the VmParams object (passed by value to the kernel)
bool vm_jit_launch(JitKernel *k, const void *params, unsigned grid, unsigned block, void *stream, std::string &err);
void vm_jit_free(JitKernel *k);

}  // namespace krb
