// Device runtime behind the C ABI: device buffers owned by plans, lane chunking, kernel launches, result delivery.
// No CPU fallback: every entry point that computes fails with KRB_ERR_NO_DEVICE when there is no CUDA device.
//
// Stream discipline
//   * device-memory calls enqueue everything on the caller's stream and return without synchronising (lane chunks reuse
//     the staging buffers in stream order);
//   * host-memory calls pipeline lane chunks over the caller's stream (H2D + kernels + pack) and an internal copy stream
//     (D2H of the previous chunk), double-buffered, and return when the last byte has landed in the caller's buffer.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "chain_kernel.cuh"
#include "chain_block_kernel.cuh"
#include "dense_tran_kernel.cuh"
#include "plan.hpp"
#include "runtime.hpp"
#include "vm_jit.hpp"
#include "vm_kernel.cuh"

namespace krb {

static thread_local std::string g_err;
void set_error(const std::string &s) { g_err = s; }
const char *get_error() { return g_err.c_str(); }

static std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }
void add_launches(long long k) { g_launches += k; }

#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);                    \
            return (_e == cudaErrorMemoryAllocation) ? KRB_ERR_OUT_OF_MEMORY : KRB_ERR_CUDA;                   \
        }                                                                                                      \
    } while (0)

int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
// per-device caches of function attributes / occupancy (a process may drive several GPUs; the values are idempotent, so
// concurrent first calls at worst repeat the query)
constexpr int KRB_MAX_DEVICES = 64;
static int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev >= 0 && dev < KRB_MAX_DEVICES ? dev : 0;
}

struct Buf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return KRB_OK;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        CUDA_TRY(cudaMalloc(&p, need));
        bytes = need;
        return KRB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct DeviceState {
    bool uploaded = false;
    int device = -1;                     // the device the plan's buffers live on (set at first use; a plan stays on one device)
    uint2 *sec[3] = {nullptr, nullptr, nullptr};
    uint32_t *chain_rowinfo = nullptr;
    uint4 *chain_stream = nullptr;
    int *chain_unknown = nullptr;
    int *chain_u2_rows = nullptr;
    uint32_t *chain_runs = nullptr;
    uint2 *block_rows = nullptr;         // block-partitioned chain kernel tables
    uint4 *block_dict = nullptr;
    int *block_gen = nullptr;
    DenseElem *dense_elems = nullptr;    // dense transient executor tables
    DenseTerm *dense_mat = nullptr, *dense_rhs = nullptr;
    int *dense_state = nullptr, *dense_col_order = nullptr;
    double *B = nullptr;                 // static constant pool (uploaded once: the plan owns a fixed copy of the circuit)
    Buf Bstep, IN, G, XOP, status, iters;
    Buf X[2], packed[2];                 // double-buffered for host delivery
    Buf selbuf[2], select;               // result selection (krb_io.select): compacted rows, the device copy of the index list
    std::vector<double> bstep_host;
    int bstep_rows = -1;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t compute_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    size_t events_used = 0;
    long long chunk_lanes = 0;           // lane capacity the buffers were sized for
    JitKernel *jit = nullptr;            // circuit-specialised lane kernel (vm_jit.cpp); null: interpreter
};

DeviceState *device_state_new() { return new DeviceState(); }
size_t vm_params_size() { return sizeof(VmParams); }
bool device_state_has_jit(const DeviceState *d) { return d && d->jit; }

static double (*g_last_fn)(void *) = nullptr;
static void *g_last_obj = nullptr;
static double g_last_ms = -1.0;   // kept when the plan that owned the events is destroyed
void set_last_timer(double (*fn)(void *), void *obj) { g_last_fn = fn; g_last_obj = obj; }
void retire_last_timer(void *obj) {
    if (g_last_obj == obj && g_last_fn) { g_last_ms = g_last_fn(obj); g_last_fn = nullptr; g_last_obj = nullptr; }
}
static double sum_event_ms(DeviceState *d) {
    if (!d || d->events_used == 0) return -1.0;
    double total = 0.0;
    for (size_t i = 0; i < d->events_used; i++) {
        if (cudaEventSynchronize(d->events[i].second) != cudaSuccess) return -1.0;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, d->events[i].first, d->events[i].second) != cudaSuccess) return -1.0;
        total += ms;
    }
    return total;
}
static double sum_event_ms_thunk(void *d) { return sum_event_ms((DeviceState *)d); }
double last_kernel_ms() { return g_last_fn ? g_last_fn(g_last_obj) : g_last_ms; }

void device_state_free(DeviceState *d) {
    if (!d) return;
    retire_last_timer(d);
    vm_jit_free(d->jit);
    for (int i = 0; i < 3; i++) if (d->sec[i]) cudaFree(d->sec[i]);
    if (d->chain_rowinfo) cudaFree(d->chain_rowinfo);
    if (d->chain_stream) cudaFree(d->chain_stream);
    if (d->chain_unknown) cudaFree(d->chain_unknown);
    if (d->chain_u2_rows) cudaFree(d->chain_u2_rows);
    if (d->chain_runs) cudaFree(d->chain_runs);
    if (d->block_rows) cudaFree(d->block_rows);
    if (d->block_dict) cudaFree(d->block_dict);
    if (d->block_gen) cudaFree(d->block_gen);
    if (d->B) cudaFree(d->B);
    if (d->dense_elems) cudaFree(d->dense_elems);
    if (d->dense_mat) cudaFree(d->dense_mat);
    if (d->dense_rhs) cudaFree(d->dense_rhs);
    if (d->dense_state) cudaFree(d->dense_state);
    if (d->dense_col_order) cudaFree(d->dense_col_order);
    d->Bstep.release(); d->IN.release(); d->G.release(); d->XOP.release(); d->status.release(); d->iters.release(); d->select.release();
    for (int i = 0; i < 2; i++) {
        d->X[i].release(); d->packed[i].release(); d->selbuf[i].release();
        if (d->compute_done[i]) cudaEventDestroy(d->compute_done[i]);
        if (d->copy_done[i]) cudaEventDestroy(d->copy_done[i]);
    }
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    for (auto &e : d->events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    delete d;
}

template <typename V>
static int upload_vec(const std::vector<V> &v, V **out) {
    CUDA_TRY(cudaMalloc((void **)out, std::max<size_t>(v.size(), 1) * sizeof(V)));
    if (!v.empty()) CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
    return KRB_OK;
}

static int upload_plan(Plan &plan, DeviceState &d, long long points) {
    if (d.uploaded) return KRB_OK;
    int rc;
    for (int s = 0; s < 3; s++) {
        const auto &ops = plan.sec[s].ops;
        if (ops.empty()) continue;
        // pad with END ops: the interpreter prefetches one 32-op chunk ahead
        std::vector<uint32_t> padded(ops);
        const size_t n_ops = padded.size() / 2;
        padded.resize((((n_ops + 31) / 32) * 32 + 64) * 2, 0u);
        if ((rc = upload_vec(padded, (uint32_t **)&d.sec[s]))) return rc;
    }
    if (plan.chain.valid) {
        if ((rc = upload_vec(plan.chain.rowinfo, &d.chain_rowinfo))) return rc;
        if ((rc = upload_vec(plan.chain.stream, (uint64_t **)&d.chain_stream))) return rc;
        if ((rc = upload_vec(plan.chain.unknown_of_row, &d.chain_unknown))) return rc;
        if ((rc = upload_vec(plan.chain.u2_rows, &d.chain_u2_rows))) return rc;
        if ((rc = upload_vec(plan.chain.runs, &d.chain_runs))) return rc;
        if (plan.chain.block.valid) {
            if ((rc = upload_vec(plan.chain.block.rows, (uint32_t **)&d.block_rows))) return rc;
            if ((rc = upload_vec(plan.chain.block.dict, (uint64_t **)&d.block_dict))) return rc;
            if ((rc = upload_vec(plan.chain.block.gen_rows, &d.block_gen))) return rc;
        }
    }
    if (plan.dense.valid) {
        static_assert(sizeof(DenseElem) == sizeof(DensePlan::Elem) && sizeof(DenseTerm) == sizeof(DensePlan::Term), "dense table layouts");
        if ((rc = upload_vec(plan.dense.elems, (DensePlan::Elem **)&d.dense_elems))) return rc;
        if ((rc = upload_vec(plan.dense.mat_terms, (DensePlan::Term **)&d.dense_mat))) return rc;
        if ((rc = upload_vec(plan.dense.rhs_terms, (DensePlan::Term **)&d.dense_rhs))) return rc;
        if ((rc = upload_vec(plan.state_elems, &d.dense_state))) return rc;
        if ((rc = upload_vec(plan.sym.col_order, &d.dense_col_order))) return rc;
    }
    std::vector<double> pool;
    evaluate_pool(plan, plan.circuit, plan.h_comp, pool);
    if ((rc = upload_vec(pool, &d.B))) return rc;
    CUDA_TRY(cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(cudaEventCreateWithFlags(&d.compute_done[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&d.copy_done[i], cudaEventDisableTiming));
    }
    if (!plan.use_chain && !plan.use_front && vm_jit_enabled() && vm_jit_eligible(plan) && vm_jit_worthwhile(plan, points)) {
        // lane programs run as a kernel specialised to the circuit (scratch cells in registers, no op decode); decided at the plan's first solve
        NvtxRange nvtx_jit("krb plan: specialise the lane kernel (NVRTC)");
        std::string why;
        d.jit = vm_jit_build(plan, sizeof(VmParams), why);
        static bool warned = false;
        if (!d.jit && !warned) {
            warned = true;
            fprintf(stderr, "krets_b200: specialised lane kernel unavailable, lane programs stay on the interpreter: %s\n", why.c_str());
        }
    }
    d.uploaded = true;
    return KRB_OK;
}

template <bool CPLX, int TPB>
static int launch_vm_t(const VmParams &p, long long Lpad, int n_slots, cudaStream_t st) {
    const size_t smem = (size_t)std::max(n_slots, 1) * TPB * (CPLX ? 16 : 8);
    static size_t configured_dev[KRB_MAX_DEVICES] = {};
    size_t &configured = configured_dev[current_device_slot()];
    if (smem > configured) {
        CUDA_TRY(cudaFuncSetAttribute(krb_lane_vm<CPLX, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        configured = std::max<size_t>(smem, 48 * 1024);
    }
    krb_lane_vm<CPLX, TPB><<<(unsigned)(Lpad / TPB), TPB, smem, st>>>(p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return KRB_OK;
}
static int launch_vm(bool cplx, int tpb, const VmParams &p, long long Lpad, int n_slots, cudaStream_t st, JitKernel *jit) {
    if (jit) {
        std::string why;
        if (!vm_jit_launch(jit, &p, (unsigned)(Lpad / tpb), (unsigned)tpb, (void *)st, why)) { set_error(why); return KRB_ERR_CUDA; }
        g_launches++;
        return KRB_OK;
    }
    if (cplx) {
        if (tpb == 32) return launch_vm_t<true, 32>(p, Lpad, n_slots, st);
        if (tpb == 128) return launch_vm_t<true, 128>(p, Lpad, n_slots, st);
        return launch_vm_t<true, 64>(p, Lpad, n_slots, st);
    }
    if (tpb == 32) return launch_vm_t<false, 32>(p, Lpad, n_slots, st);
    if (tpb == 128) return launch_vm_t<false, 128>(p, Lpad, n_slots, st);
    return launch_vm_t<false, 64>(p, Lpad, n_slots, st);
}

template <bool CPLX, int TPB, bool DIRECT>
static int launch_chain_t(const ChainParams &p, long long Lpad, cudaStream_t st) {
    krb_chain_kernel<CPLX, TPB, DIRECT><<<(unsigned)(Lpad / TPB), TPB, 0, st>>>(p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return KRB_OK;
}
template <bool DIRECT>
static int launch_chain_d(bool cplx, int tpb, const ChainParams &p, long long Lpad, cudaStream_t st) {
    if (cplx) {
        if (tpb == 32) return launch_chain_t<true, 32, DIRECT>(p, Lpad, st);
        if (tpb == 128) return launch_chain_t<true, 128, DIRECT>(p, Lpad, st);
        return launch_chain_t<true, 64, DIRECT>(p, Lpad, st);
    }
    if (tpb == 32) return launch_chain_t<false, 32, DIRECT>(p, Lpad, st);
    if (tpb == 128) return launch_chain_t<false, 128, DIRECT>(p, Lpad, st);
    return launch_chain_t<false, 64, DIRECT>(p, Lpad, st);
}
static int launch_chain(bool cplx, int tpb, bool direct, const ChainParams &p, long long Lpad, cudaStream_t st) {
    return direct ? launch_chain_d<true>(cplx, tpb, p, Lpad, st) : launch_chain_d<false>(cplx, tpb, p, Lpad, st);
}

// block-partitioned chain kernel: persistent CTAs (one lane at a time each), grid = resident CTAs of the device
template <bool CPLX, int T, int M, bool SD>
static int launch_chain_block_t(const ChainBlockParams &p, cudaStream_t st) {
    typedef typename ChainSc<CPLX>::T V;
    const size_t smem = block_smem_elems<CPLX, T, M, SD>() * sizeof(V);
    static int resident_dev[KRB_MAX_DEVICES] = {};
    int &resident = resident_dev[current_device_slot()];
    if (!resident) {
        CUDA_TRY(cudaFuncSetAttribute(krb_chain_block_kernel<CPLX, T, M, SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0, sms = 0, per_sm = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, krb_chain_block_kernel<CPLX, T, M, SD>, T, smem));
        if (per_sm < 1) { set_error("block chain kernel does not fit on an SM"); return KRB_ERR_INTERNAL; }
        resident = sms * per_sm;
    }
    static const char *grid_env = getenv("KRB_BLOCK_GRID");             // profiling: CTAs to launch instead of all resident ones
    const unsigned grid = (unsigned)std::min<long long>(p.lanes, grid_env && atoi(grid_env) > 0 ? atoi(grid_env) : resident);
    krb_chain_block_kernel<CPLX, T, M, SD><<<grid, T, smem, st>>>(p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return KRB_OK;
}
template <bool SD>
static int launch_chain_block_s(bool cplx, int T, int M, const ChainBlockParams &p, cudaStream_t st) {
    if (cplx && T == 256) {            // profiling shape (KRB_CHAIN_BLOCK_CT256 at analysis time)
        switch (M) {
        case 2: return launch_chain_block_t<true, 256, 2, SD>(p, st);
        case 4: return launch_chain_block_t<true, 256, 4, SD>(p, st);
        case 6: return launch_chain_block_t<true, 256, 6, SD>(p, st);
        case 8: return launch_chain_block_t<true, 256, 8, SD>(p, st);
        case 10: return launch_chain_block_t<true, 256, 10, SD>(p, st);
        }
    } else if (cplx) {
        switch (M) {
        case 2: return launch_chain_block_t<true, 512, 2, SD>(p, st);
        case 4: return launch_chain_block_t<true, 512, 4, SD>(p, st);
        case 6: return launch_chain_block_t<true, 512, 6, SD>(p, st);
        case 8: return launch_chain_block_t<true, 512, 8, SD>(p, st);
        case 10: return launch_chain_block_t<true, 512, 10, SD>(p, st);
        }
    } else {
        switch (M) {   // odd: a thread's column and the lane's contiguous result vector are both conflict-free
        case 5: return launch_chain_block_t<false, 256, 5, SD>(p, st);
        case 9: return launch_chain_block_t<false, 256, 9, SD>(p, st);
        case 13: return launch_chain_block_t<false, 256, 13, SD>(p, st);
        case 17: return launch_chain_block_t<false, 256, 17, SD>(p, st);
        case 21: return launch_chain_block_t<false, 256, 21, SD>(p, st);
        }
    }
    set_error("block chain kernel: no compiled shape for this plan");
    return KRB_ERR_INTERNAL;
}
static int launch_chain_block(bool cplx, int T, int M, bool shared_dict, const ChainBlockParams &p, cudaStream_t st) {
    return shared_dict ? launch_chain_block_s<true>(cplx, T, M, p, st) : launch_chain_block_s<false>(cplx, T, M, p, st);
}

template <int N, int NSP, int NWP>
static int launch_dense_t(const DenseTranParams &p, long long Lpad, cudaStream_t st) {
    constexpr int TPB = 32;
    const size_t per_thread = (size_t)2 * N * N + N + (size_t)N * NSP + (size_t)N * NWP + (size_t)NSP * N;
    const size_t smem = per_thread * TPB * 8 + (size_t)2 * 32 * NWP * 8;      // coefficient columns + the staged source table
    static size_t configured_dev[KRB_MAX_DEVICES] = {};
    size_t &configured = configured_dev[current_device_slot()];
    if (smem > configured) {
        CUDA_TRY(cudaFuncSetAttribute(krb_dense_tran<N, NSP, NWP, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        configured = std::max<size_t>(smem, 48 * 1024);
    }
    krb_dense_tran<N, NSP, NWP, TPB><<<(unsigned)(Lpad / TPB), TPB, smem, st>>>(p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return KRB_OK;
}
template <int N>
static int launch_dense_n(const DenseTranParams &p, long long Lpad, cudaStream_t st) {
    // padded state / source counts: {0, 2, 4, 8} x {1, 4}
    const int nsp = p.ns == 0 ? 0 : p.ns <= 2 ? 2 : p.ns <= 4 ? 4 : 8, nwp = p.nw <= 1 ? 1 : 4;
    if (nsp == 0) return nwp == 1 ? launch_dense_t<N, 0, 1>(p, Lpad, st) : launch_dense_t<N, 0, 4>(p, Lpad, st);
    if (nsp == 2) return nwp == 1 ? launch_dense_t<N, 2, 1>(p, Lpad, st) : launch_dense_t<N, 2, 4>(p, Lpad, st);
    if (nsp == 4) return nwp == 1 ? launch_dense_t<N, 4, 1>(p, Lpad, st) : launch_dense_t<N, 4, 4>(p, Lpad, st);
    return nwp == 1 ? launch_dense_t<N, 8, 1>(p, Lpad, st) : launch_dense_t<N, 8, 4>(p, Lpad, st);
}
static int launch_dense(const DenseTranParams &p, long long Lpad, cudaStream_t st) {
    switch (p.n) {
    case 1: return launch_dense_n<1>(p, Lpad, st);
    case 2: return launch_dense_n<2>(p, Lpad, st);
    case 3: return launch_dense_n<3>(p, Lpad, st);
    case 4: return launch_dense_n<4>(p, Lpad, st);
    case 5: return launch_dense_n<5>(p, Lpad, st);
    case 6: return launch_dense_n<6>(p, Lpad, st);
    case 7: return launch_dense_n<7>(p, Lpad, st);
    case 8: return launch_dense_n<8>(p, Lpad, st);
    default: set_error("dense transient executor: n out of range"); return KRB_ERR_INTERNAL;
    }
}

static int pick_tpb(const Plan &plan) {
    const int t = plan.opts.threads_per_block;
    if (t == 32 || t == 64 || t == 128) return t;
    return 64;
}

int run_lanes(Plan *plan_first, Plan *plan_steps, const RunArgs &a) {
    if (device_count() <= 0) { set_error("no CUDA device: krets_b200 has no CPU fallback"); return KRB_ERR_NO_DEVICE; }
    Plan &P0 = *plan_first;
    Plan *PS = plan_steps;
    if (a.sweep_rows && (P0.use_front || P0.use_chain || PS || P0.input_sweep < 0)) {
        // only Newton-Raphson makes a sweep point depend on its predecessor, and nonlinear circuits always compile to a lane program
        set_error("sequential DC sweep: the plan does not run on the lane VM (linear chain / front plans: results do not depend on the warm start, use krb_solve_dc_sweep)");
        return KRB_ERR_UNSUPPORTED;
    }
    if (P0.use_front) return run_front(plan_first, plan_steps, a);
    if (PS && PS->use_front) { set_error("transient: plan_tran uses the front executor but plan_dc does not (analyse both with KRB_EXEC_STAGED)"); return KRB_ERR_UNSUPPORTED; }
    if (!P0.dev) P0.dev = device_state_new();
    if (PS && !PS->dev) PS->dev = device_state_new();
    DeviceState &D0 = *P0.dev;
    {   // a plan's device buffers are tied to the device of its first solve; one solve at a time per plan (include/krets_b200.h)
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        if (D0.device < 0) D0.device = dev;
        if (PS && PS->dev->device < 0) PS->dev->device = dev;
        if (D0.device != dev || (PS && PS->dev->device != dev)) {
            set_error("this plan was first used on another CUDA device; analyse one plan per device");
            return KRB_ERR_INVALID_ARGUMENT;
        }
    }
    cudaStream_t st = (cudaStream_t)a.stream;
    const bool cplx = P0.is_complex;
    const size_t ts = cplx ? 16 : 8;
    const int n = P0.circuit.n;
    const int rows = a.rows;
    const int tpb = pick_tpb(P0);
    int rc;
    if ((rc = upload_plan(P0, D0, PS ? a.lanes : a.lanes * (long long)a.rows))) return rc;
    if (PS && (rc = upload_plan(*PS, *PS->dev, a.lanes * (long long)a.rows))) return rc;
    if (PS && PS->n_inputs != P0.n_inputs) { set_error("plan_dc and plan_tran must be analysed with the same varied elements"); return KRB_ERR_INVALID_ARGUMENT; }
    if (n == 0 || a.lanes <= 0) return KRB_OK;
    const bool chain = P0.use_chain && !PS;
    const int width = n + (PS ? (int)PS->state_elems.size() : 0);   // result-row columns incl. trapezoidal state
    const bool out_dev = (a.memory == KRB_MEM_DEVICE);
    if (a.n_select > 0 && a.layout != KRB_LAYOUT_POINT_MAJOR) { set_error("result selection delivers point-major rows only"); return KRB_ERR_UNSUPPORTED; }
    // the chain kernel can write point-major results itself when pivot row k is unknown k (no X staging, no pack kernel)
    // block-partitioned chain kernel: one CTA per lane, results delivered point-major from shared memory (any pivot order),
    // no factor planes in HBM.  KRB_CHAIN_SEQ=1 keeps the sequential kernel (tests, A/B measurements).
    const bool block = chain && P0.chain.block.valid && a.layout == KRB_LAYOUT_POINT_MAJOR && rows == 1 && getenv("KRB_CHAIN_SEQ") == nullptr;
    // result selection (extension): the block kernel writes only the selected unknowns itself when pivot row k is unknown k; every
    // other path delivers full rows into an internal buffer first and a gather kernel compacts them (device->host traffic shrinks
    // either way, which is the point)
    const int nsel = a.n_select;
    const bool sel = nsel > 0;
    const bool sel_in_kernel = sel && block && P0.chain.identity_order;
    const bool sel_gather = sel && !sel_in_kernel;
    const bool direct = block || (chain && P0.chain.identity_order && a.layout == KRB_LAYOUT_POINT_MAJOR && rows == 1);
    // the dense transient kernel delivers point-major rows itself: only the operating-point row is staged in planes
    const bool dense_direct = PS && rows > 1 && PS->dense.valid && a.layout == KRB_LAYOUT_POINT_MAJOR && getenv("KRB_NO_DENSE_TRAN") == nullptr;
    const int plane_rows = dense_direct ? 1 : rows;

    // ---- per-step source table (transient) ----
    if (PS) {
        DeviceState &DS = *PS->dev;
        if (DS.bstep_rows != rows) {
            // row r, iteration 0 sees the time key of the previous row, later iterations see r*h
            const int nw = (int)PS->wave_elems.size();
            const double h = PS->opts.time_step;
            DS.bstep_host.assign((size_t)rows * 2 * std::max(nw, 1), 0.0);
            for (int r = 1; r < rows; r++)
                for (int k = 0; k < nw; k++) {
                    const krb_element &el = PS->circuit.el[PS->wave_elems[k]];
                    DS.bstep_host[((size_t)r * 2 + 0) * nw + k] = transient_value_at(el, (double)(PS->opts.integrator == KRB_INTEGRATOR_BE_KRETS ? r - 1 : r) * h);   // stampable.rs:1043 + transient.rs:49 (lag only in reference mode)
                    DS.bstep_host[((size_t)r * 2 + 1) * nw + k] = transient_value_at(el, (double)r * h);         // transient.rs:95,105
                }
            if ((rc = DS.Bstep.ensure(DS.bstep_host.size() * 8))) return rc;
            CUDA_TRY(cudaMemcpy(DS.Bstep.p, DS.bstep_host.data(), DS.bstep_host.size() * 8, cudaMemcpyHostToDevice));
            DS.bstep_rows = rows;
        }
    }

    // ---- chunking: bound the per-call device footprint ----
    // block kernel without a shared dictionary: pre-stamped rows E[lane][T*M][4] + U2[lane][T] live in the G buffer (counted in planes)
    const bool block_stamped = block && !P0.chain.block.shared_dict;
    const int planes0 = block ? (block_stamped ? P0.chain.block.T * (4 * P0.chain.block.M + 1) : 0) : chain ? 2 * n + P0.chain.n_u2 : P0.n_planes, planesS = PS ? PS->n_planes : 0;
    const int nbuf = out_dev ? 1 : 2;
    const bool staged = !out_dev || sel_gather;      // full rows are packed into packed[b] before they leave (host copy / gather)
    const size_t per_lane = (size_t)(planes0 + planesS) * ts + (size_t)plane_rows * width * ts * (out_dev ? 1 : 2) + (staged ? (size_t)rows * width * ts * 2 : 0) +
                            (size_t)P0.n_inputs * 8 + (a.op_solution ? (size_t)n * 8 : 0) + 16;
    long long want = a.chunk_lanes > 0 ? a.chunk_lanes : (out_dev ? 65536 : 16384);
    want = std::min<long long>(want, ((a.lanes + tpb - 1) / tpb) * tpb);
    want = std::max<long long>(tpb, (want / tpb) * tpb);
    if (want > D0.chunk_lanes || (PS && want > PS->dev->chunk_lanes)) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        size_t reusable = D0.G.bytes + D0.X[0].bytes + D0.X[1].bytes + D0.packed[0].bytes + D0.packed[1].bytes + D0.IN.bytes + (PS ? PS->dev->G.bytes : 0);
        long long max_chunk = (long long)(((free_b + reusable) * 0.80) / (double)per_lane);
        max_chunk = (max_chunk / tpb) * tpb;
        if (max_chunk < tpb) { set_error("not enough device memory for one block of lanes"); return KRB_ERR_OUT_OF_MEMORY; }
        want = std::min(want, max_chunk);
    }
    const long long chunk = want;
    if ((rc = D0.IN.ensure(std::max<size_t>((size_t)P0.n_inputs, 1) * chunk * 8))) return rc;
    if ((rc = D0.G.ensure(std::max<size_t>((size_t)planes0, 1) * chunk * ts))) return rc;
    if (PS && (rc = PS->dev->G.ensure(std::max<size_t>((size_t)planesS, 1) * chunk * ts))) return rc;
    for (int b = 0; b < nbuf; b++) {
        if (!direct && (rc = D0.X[b].ensure((size_t)plane_rows * width * chunk * ts))) return rc;
        if (staged && !sel_in_kernel && (rc = D0.packed[b].ensure((size_t)rows * n * chunk * ts))) return rc;
        if (sel && !out_dev && (rc = D0.selbuf[b].ensure((size_t)rows * nsel * chunk * ts))) return rc;
    }
    if (sel) {
        if ((rc = D0.select.ensure((size_t)nsel * 4))) return rc;
        CUDA_TRY(cudaMemcpyAsync(D0.select.p, a.select, (size_t)nsel * 4, cudaMemcpyHostToDevice, (cudaStream_t)a.stream));
    }
    if ((rc = D0.status.ensure((size_t)chunk * 4 * 2))) return rc;
    if (a.iterations && (rc = D0.iters.ensure((size_t)chunk * 4 * 2))) return rc;
    if (a.op_solution && (rc = D0.XOP.ensure((size_t)n * chunk * 8))) return rc;
    D0.chunk_lanes = std::max(D0.chunk_lanes, chunk);
    if (PS) PS->dev->chunk_lanes = std::max(PS->dev->chunk_lanes, chunk);

    D0.events_used = 0;
    set_last_timer(sum_event_ms_thunk, &D0);
    auto next_events = [&](cudaEvent_t &e0, cudaEvent_t &e1) -> int {
        if (D0.events_used == D0.events.size()) {
            cudaEvent_t x, y;
            CUDA_TRY(cudaEventCreate(&x));
            CUDA_TRY(cudaEventCreate(&y));
            D0.events.push_back({x, y});
        }
        e0 = D0.events[D0.events_used].first;
        e1 = D0.events[D0.events_used].second;
        D0.events_used++;
        return KRB_OK;
    };

    Buf op_tmp;
    int chunk_idx = 0;
    for (long long lane0 = 0; lane0 < a.lanes; lane0 += chunk, ++chunk_idx) {
        const long long lanes = std::min<long long>(chunk, a.lanes - lane0);
        const long long Lpad = ((lanes + tpb - 1) / tpb) * tpb;
        const int b = out_dev ? 0 : (chunk_idx & 1);
        double *IN = (double *)D0.IN.p;
        bool mc_in_place = false;
        int *status_d = (int *)D0.status.p + (size_t)b * chunk;
        int *iters_d = a.iterations ? (int *)D0.iters.p + (size_t)b * chunk : nullptr;
        if (!out_dev && chunk_idx >= 2) CUDA_TRY(cudaStreamWaitEvent(st, D0.copy_done[b], 0));   // buffer set b is free again
        // ---- inputs ----
        {
        NvtxRange nvtx_inputs("krb lanes: inputs (sweep values / frequencies / Monte-Carlo planes)");
        {
            // Monte-Carlo planes normally occupy one contiguous run of input planes in the caller's order: one strided 2-D copy
            // instead of one copy per plane (a ladder with every resistor per-lane has 5000 of them: 20 ms of launches per chunk)
            const int nv = (int)P0.input_of_varied.size();
            bool run = nv > 1 && a.varied_values != nullptr;
            for (int q = 1; q < nv && run; q++) run = P0.input_of_varied[q] == P0.input_of_varied[0] + q;
            for (int q = 0; q < nv && run; q++) run = P0.input_of_varied[q] != P0.input_sweep && P0.input_of_varied[q] != P0.input_freq;
            // device-resident Monte-Carlo values that already have the kernels' layout ([plane][lane], one chunk, no lane padding) are read in
            // place by the pre-stamping kernel of the block chain executor — its only reader — instead of being copied into the plan's buffer
            mc_in_place = run && block_stamped && out_dev && lane0 == 0 && lanes == a.lanes && Lpad == lanes;
            if (run && !mc_in_place)
                CUDA_TRY(cudaMemcpy2DAsync(IN + (size_t)P0.input_of_varied[0] * Lpad, (size_t)Lpad * 8, a.varied_values + lane0, (size_t)a.lanes * 8,
                                           (size_t)lanes * 8, (size_t)nv, out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
            for (int k = 0; k < P0.n_inputs; k++) {
                const double *src = nullptr;
                if (k == P0.input_sweep && a.sweep_rows) continue;   // filled row by row in front of each launch
                if (k == P0.input_sweep || k == P0.input_freq) src = a.lane_param;
                else {
                    if (run) continue;
                    int vk = -1;
                    for (size_t q = 0; q < P0.input_of_varied.size(); q++) if (P0.input_of_varied[q] == k) { vk = (int)q; break; }
                    if (vk < 0 || !a.varied_values) { set_error("missing per-lane values for a varied element"); return KRB_ERR_INVALID_ARGUMENT; }
                    src = a.varied_values + (size_t)vk * a.lanes;
                }
                CUDA_TRY(cudaMemcpyAsync(IN + (size_t)k * Lpad, src + lane0, (size_t)lanes * 8, out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
            }
        }
        if (Lpad > lanes && P0.n_inputs > 0) {
            const long long work = (long long)P0.n_inputs * (Lpad - lanes);
            krb_fill_pad<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(IN, Lpad, lanes, P0.n_inputs);
            g_launches++;
        }
        if (a.op_solution) {
            const bool bc = (a.lanes_op == 1);
            const size_t cnt = bc ? (size_t)n : (size_t)lanes * n;
            const double *dsrc = a.op_solution + (bc ? 0 : (size_t)lane0 * n);
            if (!out_dev) {
                if ((rc = op_tmp.ensure(cnt * 8))) return rc;
                CUDA_TRY(cudaMemcpyAsync(op_tmp.p, dsrc, cnt * 8, cudaMemcpyHostToDevice, st));
                dsrc = (const double *)op_tmp.p;
            }
            const long long work = (long long)n * Lpad;
            krb_spread_op<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(dsrc, (double *)D0.XOP.p, n, Lpad, lanes, bc ? 1 : 0);
            g_launches++;
        }
        CUDA_TRY(cudaMemsetAsync(status_d, 0, (size_t)Lpad * 4, st));
        }

        // ---- compute ----
        {
        NvtxRange nvtx_compute(block ? "krb lanes: stamp + factor + solve (krb_chain_block_kernel)" : chain ? "krb lanes: stamp + factor + solve (krb_chain_kernel)"
                                                                                                              : "krb lanes: stamp + factor + solve (krb_lane_vm / krb_dense_tran)");
        cudaEvent_t e0, e1;
        if ((rc = next_events(e0, e1))) return rc;
        CUDA_TRY(cudaEventRecord(e0, st));
        if (block) {
            ChainBlockParams c{};
            c.rows = D0.block_rows; c.dict = D0.block_dict; c.stream = D0.chain_stream;
            c.unknown_of_row = P0.chain.identity_order ? nullptr : D0.chain_unknown;
            c.IN = IN; c.lanes = lanes; c.L = Lpad; c.n = n; c.input_freq = P0.input_freq; c.status = status_d;
            c.gen_rows = D0.block_gen; c.n_dict = P0.chain.block.n_dict; c.n_gen = (int)P0.chain.block.gen_rows.size();
            if (sel_in_kernel) {          // compact rows straight from the kernel
                c.OUT = out_dev ? (void *)((char *)a.x_out + (size_t)lane0 * (size_t)nsel * ts) : D0.selbuf[b].p;
                c.select = (const int *)D0.select.p; c.n_select = nsel;
            } else c.OUT = (out_dev && !sel_gather) ? (void *)((char *)a.x_out + (size_t)lane0 * (size_t)n * ts) : D0.packed[b].p;
            // one bulk asynchronous copy per lane (TMA engine) when pivot row k is unknown k; complex results need a 16-byte aligned
            // buffer (real ones are handled at any 8-byte alignment).  KRB_BLOCK_NO_BULK=1: per-thread stores (A/B measurements)
            static const bool no_bulk = getenv("KRB_BLOCK_NO_BULK") != nullptr;
            c.bulk_store = (P0.chain.identity_order && !sel_in_kernel && !no_bulk && (!cplx || ((size_t)c.OUT & 15u) == 0)) ? 1 : 0;
            if (block_stamped) {
                // (a) stamping as its own launch: thread per (pivot row, lane), coalesced inputs, records transposed to lane-major
                const int T = P0.chain.block.T, M = P0.chain.block.M;
                BlockStampParams sp{};
                sp.rows = D0.block_rows; sp.dict = D0.block_dict; sp.stream = D0.chain_stream; sp.IN = IN; sp.L = Lpad; sp.lanes = lanes;
                sp.input_freq = P0.input_freq; sp.slots = T * M; sp.T = T;
                sp.E = D0.G.p; sp.U2 = (char *)D0.G.p + (size_t)chunk * (size_t)T * M * 4 * ts;
                sp.INV = IN; sp.inv_base = 0; sp.inv_n = 0; sp.inv_L = Lpad;
                if (mc_in_place) { sp.INV = a.varied_values; sp.inv_base = (unsigned)P0.input_of_varied[0]; sp.inv_n = (unsigned)P0.input_of_varied.size(); sp.inv_L = a.lanes; }
                const dim3 grid((unsigned)(T * M / (cplx ? 16 : 32)), (unsigned)((lanes + 31) / 32)), blk(32, 8);
                if (cplx) krb_block_stamp_rows<true><<<grid, blk, 0, st>>>(sp);
                else krb_block_stamp_rows<false><<<grid, blk, 0, st>>>(sp);
                g_launches++;
                CUDA_TRY(cudaGetLastError());
                c.E = sp.E; c.U2 = sp.U2;
            }
            if ((rc = launch_chain_block(cplx, P0.chain.block.T, P0.chain.block.M, P0.chain.block.shared_dict, c, st))) return rc;
        } else if (chain) {
            ChainParams c{};
            c.rowinfo = D0.chain_rowinfo; c.stream = D0.chain_stream; c.unknown_of_row = D0.chain_unknown; c.IN = IN;
            c.UY = D0.G.p; c.X = D0.X[b].p; c.n = n; c.input_freq = P0.input_freq; c.n_u2 = P0.chain.n_u2; c.L = Lpad; c.status = status_d;
            c.u2_rows = D0.chain_u2_rows; c.out_lanes = lanes; c.runs = D0.chain_runs; c.n_runs = (int)P0.chain.runs.size();
            { const char *dbg = getenv("KRB_CHAIN_DEBUG"); c.debug = dbg ? atoi(dbg) : 0; }
            { const char *pf = getenv("KRB_CHAIN_PF"); c.pf_units = pf ? atoi(pf) : 256; }
            c.OUT = direct ? ((out_dev && !sel_gather) ? (void *)((char *)a.x_out + (size_t)lane0 * (size_t)n * ts) : D0.packed[b].p) : nullptr;
            if ((rc = launch_chain(cplx, tpb, direct, c, Lpad, st))) return rc;
        } else {
            VmParams p{};
            p.sec0 = D0.sec[0]; p.sec1 = D0.sec[1]; p.sec2 = D0.sec[2];
            if (P0.has_step_sections) p.sec0 = nullptr; else { p.sec1 = nullptr; p.sec2 = nullptr; }
            p.B = D0.B; p.Bstep = nullptr; p.n_wave = 0;
            p.IN = IN; p.G = D0.G.p; p.X = D0.X[b].p;
            p.Xprev0 = a.op_solution ? (const double *)D0.XOP.p : nullptr;
            p.n = n; p.row_width = width; p.L = Lpad;
            p.row_begin = 0; p.row_end = PS ? 1 : rows;
            p.nonlinear = (P0.circuit.nonlinear && !cplx) ? 1 : 0;
            p.max_iter = p.nonlinear ? std::max(1, a.cfg.maximum_iterations) : 1;
            p.guess_empty_at_iter0 = 1; p.pseudo_time = 0;
            p.h = P0.opts.time_step; p.reltol = a.cfg.relative_tolerance; p.iabstol = a.cfg.current_absolute_tolerance; p.vabstol = a.cfg.voltage_absolute_tolerance;
            p.status = status_d; p.iters = iters_d;
            if (width > n) CUDA_TRY(cudaMemsetAsync((char *)D0.X[b].p + (size_t)n * Lpad * ts, 0, (size_t)(width - n) * Lpad * ts, st));   // i_C(0) = 0
            if (a.sweep_rows) {
                // dc.rs:47-113: sweep points are the result rows of a lane, one launch per row (every lane of the chunk advances together).
                // Row 0 starts from the empty map (dc.rs:41), row r from row r - 1 (dc.rs:58,112); the "step" key of the result map
                // (dc.rs:97) takes part in the convergence check like the transient "time" key: |r - (r - 1)| against the voltage tolerance
                // (solver.rs:96-103, SURVEY B-5), which the kernels evaluate as |row h - (row - 1) h| with h = 1
                p.pseudo_time = 1; p.h = 1.0;
                for (int r = 0; r < rows; r++) {
                    krb_fill_plane<<<(unsigned)((Lpad + 255) / 256), 256, 0, st>>>(IN + (size_t)P0.input_sweep * Lpad, Lpad, a.sweep_start + ((double)r * a.sweep_step));   // dc.rs:48
                    g_launches++;
                    p.row_begin = r; p.row_end = r + 1;
                    p.guess_empty_at_iter0 = r == 0 ? 1 : 0;
                    if ((rc = launch_vm(cplx, tpb, p, Lpad, P0.n_slots, st, D0.jit))) return rc;
                }
            } else if ((rc = launch_vm(cplx, tpb, p, Lpad, P0.n_slots, st, D0.jit))) return rc;
            if (PS && rows > 1 && PS->dense.valid && getenv("KRB_NO_DENSE_TRAN") == nullptr) {
                // small linear circuit: native stepping kernel (row 0 = the operating point just written by the DC plan)
                DeviceState &DS = *PS->dev;
                DenseTranParams q{};
                q.elems = DS.dense_elems; q.mat_terms = DS.dense_mat; q.rhs_terms = DS.dense_rhs;
                q.n_mat = (int)PS->dense.mat_terms.size(); q.n_rhs = (int)PS->dense.rhs_terms.size();
                q.state_elems = DS.dense_state; q.col_order = DS.dense_col_order;
                q.n = n; q.ns = (int)PS->state_elems.size(); q.nw = (int)PS->wave_elems.size(); q.width = width;
                q.IN = IN; q.Bstep = (const double *)DS.Bstep.p; q.X = (double *)D0.X[b].p; q.L = Lpad; q.rows = rows;
                q.h_comp = PS->h_comp; q.status = status_d;
                q.propagate = getenv("KRB_DENSE_TRAN_LU") == nullptr;      // default: per-lane propagator steps
                q.OUT = dense_direct ? ((out_dev && !sel_gather) ? (double *)a.x_out + (size_t)lane0 * (size_t)rows * n : (double *)D0.packed[b].p) : nullptr;
                q.out_lanes = lanes;
                if ((rc = launch_dense(q, Lpad, st))) return rc;
            } else if (PS && rows > 1) {
                DeviceState &DS = *PS->dev;
                VmParams q = p;
                q.sec0 = DS.sec[0]; q.sec1 = DS.sec[1]; q.sec2 = DS.sec[2];
                if (PS->has_step_sections) q.sec0 = nullptr; else { q.sec1 = nullptr; q.sec2 = nullptr; }
                q.B = DS.B; q.Bstep = (const double *)DS.Bstep.p; q.n_wave = (int)PS->wave_elems.size();
                q.G = DS.G.p; q.Xprev0 = nullptr;
                q.row_begin = 1; q.row_end = rows;
                q.nonlinear = PS->circuit.nonlinear ? 1 : 0;
                q.max_iter = q.nonlinear ? std::max(1, a.cfg.maximum_iterations) : 1;
                q.guess_empty_at_iter0 = 0; q.pseudo_time = 1;
                q.h = PS->opts.time_step;
                q.iters = nullptr;
                if ((rc = launch_vm(false, tpb, q, Lpad, PS->n_slots, st, DS.jit))) return rc;
            }
        }
        CUDA_TRY(cudaEventRecord(e1, st));
        }

        // ---- delivery ----
        NvtxRange nvtx_deliver("krb lanes: deliver (pack / D2H)");
        const size_t per_point = (size_t)rows * n;
        const size_t per_point_out = (size_t)rows * (sel ? nsel : n);
        auto pack_full = [&](void *dst) -> int {      // plane-staged results -> point-major full rows
            dim3 grid((unsigned)((lanes + 31) / 32), (unsigned)((n + 31) / 32), (unsigned)rows), block(32, 8);
            if (cplx) krb_pack_point_major<double2><<<grid, block, 0, st>>>((const double2 *)D0.X[b].p, (double2 *)dst, rows, n, width, Lpad, lanes);
            else krb_pack_point_major<double><<<grid, block, 0, st>>>((const double *)D0.X[b].p, (double *)dst, rows, n, width, Lpad, lanes);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            return KRB_OK;
        };
        auto gather_selected = [&](void *dst) -> int {   // packed[b] (full rows) -> compact rows
            const long long points = lanes * rows, work = points * nsel;
            if (cplx) krb_select_columns<double2><<<(unsigned)((work + 255) / 256), 256, 0, st>>>((const double2 *)D0.packed[b].p, (double2 *)dst, (const int *)D0.select.p, n, nsel, points);
            else krb_select_columns<double><<<(unsigned)((work + 255) / 256), 256, 0, st>>>((const double *)D0.packed[b].p, (double *)dst, (const int *)D0.select.p, n, nsel, points);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            return KRB_OK;
        };
        const bool kernel_delivered = direct || dense_direct;   // the kernel wrote point-major rows itself (caller's buffer, packed[b] or selbuf[b])
        if (out_dev) {
            if (sel_gather) {
                if (!kernel_delivered && (rc = pack_full(D0.packed[b].p))) return rc;
                if ((rc = gather_selected((char *)a.x_out + (size_t)lane0 * per_point_out * ts))) return rc;
            } else if (kernel_delivered) {
                // already delivered by the kernel
            } else if (a.layout == KRB_LAYOUT_POINT_MAJOR) {
                if ((rc = pack_full((char *)a.x_out + (size_t)lane0 * per_point * ts))) return rc;
            } else {
                for (int rr = 0; rr < (width == n ? 1 : rows); rr++)   // state columns are skipped row by row
                    CUDA_TRY(cudaMemcpy2DAsync((char *)a.x_out + ((size_t)rr * n * a.lanes + (size_t)lane0) * ts, (size_t)a.lanes * ts, (char *)D0.X[b].p + (size_t)rr * width * Lpad * ts,
                                               (size_t)Lpad * ts, (size_t)lanes * ts, width == n ? per_point : (size_t)n, cudaMemcpyDeviceToDevice, st));
            }
            if (a.lane_status) CUDA_TRY(cudaMemcpyAsync(a.lane_status + lane0, status_d, (size_t)lanes * 4, cudaMemcpyDeviceToDevice, st));
            if (a.iterations) CUDA_TRY(cudaMemcpyAsync(a.iterations + lane0, iters_d, (size_t)lanes * 4, cudaMemcpyDeviceToDevice, st));
        } else {
            cudaStream_t cs = D0.copy_stream;
            if (a.layout == KRB_LAYOUT_POINT_MAJOR && !kernel_delivered && (rc = pack_full(D0.packed[b].p))) return rc;
            if (sel_gather && (rc = gather_selected(D0.selbuf[b].p))) return rc;
            CUDA_TRY(cudaEventRecord(D0.compute_done[b], st));
            CUDA_TRY(cudaStreamWaitEvent(cs, D0.compute_done[b], 0));
            if (a.layout == KRB_LAYOUT_POINT_MAJOR)
                CUDA_TRY(cudaMemcpyAsync((char *)a.x_out + (size_t)lane0 * per_point_out * ts, sel ? D0.selbuf[b].p : D0.packed[b].p, (size_t)lanes * per_point_out * ts, cudaMemcpyDeviceToHost, cs));
            else
                for (int rr = 0; rr < (width == n ? 1 : rows); rr++)
                    CUDA_TRY(cudaMemcpy2DAsync((char *)a.x_out + ((size_t)rr * n * a.lanes + (size_t)lane0) * ts, (size_t)a.lanes * ts, (char *)D0.X[b].p + (size_t)rr * width * Lpad * ts,
                                               (size_t)Lpad * ts, (size_t)lanes * ts, width == n ? per_point : (size_t)n, cudaMemcpyDeviceToHost, cs));
            if (a.lane_status) CUDA_TRY(cudaMemcpyAsync(a.lane_status + lane0, status_d, (size_t)lanes * 4, cudaMemcpyDeviceToHost, cs));
            if (a.iterations) CUDA_TRY(cudaMemcpyAsync(a.iterations + lane0, iters_d, (size_t)lanes * 4, cudaMemcpyDeviceToHost, cs));
            CUDA_TRY(cudaEventRecord(D0.copy_done[b], cs));
        }
    }
    if (!out_dev) {
        // host buffers must be complete on return
        CUDA_TRY(cudaStreamSynchronize(D0.copy_stream));
        CUDA_TRY(cudaStreamSynchronize(st));
        op_tmp.release();
    } else if (op_tmp.p) {
        CUDA_TRY(cudaStreamSynchronize(st));
        op_tmp.release();
    }
    return KRB_OK;
}

}  // namespace krb
