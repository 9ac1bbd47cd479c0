// Internal interface between the C ABI (api.cpp) and the CUDA runtime (runtime.cu).
#pragma once
#include <nvtx3/nvToolsExt.h>

#include <string>

#include "plan.hpp"

namespace krb {

// NVTX range per phase (analyze / inputs / stamp+factor+solve / deliver ...): header-only NVTX v3, a no-op unless a profiler
// (nsys, ncu --nvtx) injects itself; ranges bracket the host-side enqueue of the phase's kernels and copies.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct RunArgs {
    krb_config cfg;
    long long lanes = 0;
    int rows = 1;                        // result rows per lane (transient: n_steps + 1)
    const double *lane_param = nullptr;  // [lanes] DC sweep values or AC frequencies
    const double *varied_values = nullptr;   // [n_varied][lanes]
    const double *op_solution = nullptr; // AC linearisation point
    long long lanes_op = 0;
    void *x_out = nullptr;
    int *lane_status = nullptr;
    int *iterations = nullptr;
    int memory = KRB_MEM_HOST, layout = KRB_LAYOUT_POINT_MAJOR;
    void *stream = nullptr;
    long long chunk_lanes = 0;
    const int32_t *select = nullptr;     // extension: deliver only these unknowns (host pointer), n_select of them per point
    int n_select = 0;
    // sequential DC sweep (dc.rs:47-113 as written): row r of every lane solves with the sweep element at sweep_start + r * sweep_step,
    // Newton-Raphson warm-started from row r - 1 (dc.rs:58,112); lane_param is unused
    bool sweep_rows = false;
    double sweep_start = 0.0, sweep_step = 0.0;
};

void set_error(const std::string &s);
const char *get_error();
int device_count();
size_t vm_params_size();   // sizeof(krb::VmParams): the specialised lane kernels (vm_jit.cpp) take the same object by value
long long launch_count();
double last_kernel_ms();
void device_state_free(DeviceState *d);
bool device_state_has_jit(const DeviceState *d);
void front_device_free(FrontDevice *d);
void add_launches(long long k);
// krb_last_kernel_ms(): the most recent call registers how its device time is read back (events are resolved lazily so that
// device-memory calls never synchronise); an object that is destroyed first retires its timer and the value is kept
void set_last_timer(double (*fn)(void *), void *obj);
void retire_last_timer(void *obj);
int run_front(Plan *plan_first, Plan *plan_steps, const RunArgs &a);
// plan_steps != nullptr: transient (plan_first = DC plan for row 0, plan_steps = TRAN plan for rows 1..)
int run_lanes(Plan *plan_first, Plan *plan_steps, const RunArgs &a);

}  // namespace krb
