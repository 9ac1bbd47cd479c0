// Device runtime behind the C ABI: device buffers owned by plans, lane chunking, kernel launches, result packing.
// No CPU fallback: every entry point that computes fails with KRB_ERR_NO_DEVICE when there is no CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "plan.hpp"
#include "runtime.hpp"
#include "vm_kernel.cuh"
#include "chain_kernel.cuh"

namespace krb {

static thread_local std::string g_err;
void set_error(const std::string &s) { g_err = s; }
const char *get_error() { return g_err.c_str(); }

static std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }

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
    uint2 *sec[3] = {nullptr, nullptr, nullptr};
    ChainTermDev *chain_terms = nullptr;
    int *chain_unknown = nullptr;
    uint32_t *chain_u2_flags = nullptr;
    Buf B, Bstep, IN, G, X, XOP, packed, status, iters;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    size_t events_used = 0;
};

DeviceState *device_state_new() { return new DeviceState(); }
static double sum_event_ms(DeviceState *d);
static DeviceState *g_last_timed_fwd();
void device_state_free(DeviceState *d) {
    if (!d) return;
    if (g_last_timed_fwd() == d) { extern void krb_snapshot_last_ms(DeviceState *); krb_snapshot_last_ms(d); }
    for (int i = 0; i < 3; i++) if (d->sec[i]) cudaFree(d->sec[i]);
    if (d->chain_terms) cudaFree(d->chain_terms);
    if (d->chain_unknown) cudaFree(d->chain_unknown);
    if (d->chain_u2_flags) cudaFree(d->chain_u2_flags);
    d->B.release(); d->Bstep.release(); d->IN.release(); d->G.release(); d->X.release(); d->XOP.release();
    d->packed.release(); d->status.release(); d->iters.release();
    for (auto &e : d->events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    delete d;
}

static DeviceState *g_last_timed = nullptr;
static double g_last_ms = -1.0;   // kept when the plan that owned the events is destroyed
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
double last_kernel_ms() { return g_last_timed ? sum_event_ms(g_last_timed) : g_last_ms; }
static DeviceState *g_last_timed_fwd() { return g_last_timed; }
void krb_snapshot_last_ms(DeviceState *d) { g_last_ms = sum_event_ms(d); g_last_timed = nullptr; }

static int upload_programs(Plan &plan, DeviceState &d) {
    if (d.uploaded) return KRB_OK;
    for (int s = 0; s < 3; s++) {
        const auto &ops = plan.sec[s].ops;
        if (ops.empty()) continue;
        // pad with END ops: the interpreter prefetches one 32-op chunk ahead
        std::vector<uint32_t> padded(ops);
        const size_t n_ops = padded.size() / 2;
        const size_t total = ((n_ops + 31) / 32) * 32 + 64;
        padded.resize(total * 2, 0u);
        CUDA_TRY(cudaMalloc(&d.sec[s], padded.size() * sizeof(uint32_t)));
        CUDA_TRY(cudaMemcpy(d.sec[s], padded.data(), padded.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (plan.use_chain) {
        static_assert(sizeof(ChainTermDev) == sizeof(ChainTerm) && sizeof(ChainTerm) == 16, "term layout");
        CUDA_TRY(cudaMalloc(&d.chain_terms, plan.chain.terms.size() * sizeof(ChainTerm)));
        CUDA_TRY(cudaMemcpy(d.chain_terms, plan.chain.terms.data(), plan.chain.terms.size() * sizeof(ChainTerm), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(&d.chain_unknown, plan.chain.unknown_of_row.size() * sizeof(int)));
        CUDA_TRY(cudaMemcpy(d.chain_unknown, plan.chain.unknown_of_row.data(), plan.chain.unknown_of_row.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (plan.chain.n_u2 > 0) {
            const int nrows = (int)plan.chain.unknown_of_row.size();
            std::vector<uint32_t> bits((nrows + 31) / 32, 0u);
            int row = 0;
            for (const ChainTerm &t : plan.chain.terms) {
                if (row >= nrows) break;
                if (t.meta & (1u << 31)) { if (t.meta & (1u << 30)) bits[row >> 5] |= 1u << (row & 31); row++; }
            }
            CUDA_TRY(cudaMalloc(&d.chain_u2_flags, bits.size() * 4));
            CUDA_TRY(cudaMemcpy(d.chain_u2_flags, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
        }
    }
    d.uploaded = true;
    return KRB_OK;
}

template <bool CPLX, int TPB>
static int launch_chain_t(const ChainParams &p, long long Lpad, cudaStream_t st) {
    krb_chain_kernel<CPLX, TPB><<<(unsigned)(Lpad / TPB), TPB, 0, st>>>(p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return KRB_OK;
}
static int launch_chain(bool cplx, int tpb, const ChainParams &p, long long Lpad, cudaStream_t st) {
    if (cplx) {
        if (tpb == 32) return launch_chain_t<true, 32>(p, Lpad, st);
        if (tpb == 128) return launch_chain_t<true, 128>(p, Lpad, st);
        return launch_chain_t<true, 64>(p, Lpad, st);
    }
    if (tpb == 32) return launch_chain_t<false, 32>(p, Lpad, st);
    if (tpb == 128) return launch_chain_t<false, 128>(p, Lpad, st);
    return launch_chain_t<false, 64>(p, Lpad, st);
}

template <bool CPLX, int TPB>
static int launch_vm_t(const VmParams &p, long long Lpad, int n_slots, cudaStream_t st) {
    const size_t smem = (size_t)std::max(n_slots, 1) * TPB * (CPLX ? 16 : 8);
    static size_t configured = 0;
    if (smem > configured) {
        CUDA_TRY(cudaFuncSetAttribute(krb_lane_vm<CPLX, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        configured = std::max<size_t>(smem, 48 * 1024);
    }
    krb_lane_vm<CPLX, TPB><<<(unsigned)(Lpad / TPB), TPB, smem, st>>>(p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return KRB_OK;
}

static int launch_vm(bool cplx, int tpb, const VmParams &p, long long Lpad, int n_slots, cudaStream_t st) {
    if (cplx) {
        if (tpb == 32) return launch_vm_t<true, 32>(p, Lpad, n_slots, st);
        if (tpb == 128) return launch_vm_t<true, 128>(p, Lpad, n_slots, st);
        return launch_vm_t<true, 64>(p, Lpad, n_slots, st);
    }
    if (tpb == 32) return launch_vm_t<false, 32>(p, Lpad, n_slots, st);
    if (tpb == 128) return launch_vm_t<false, 128>(p, Lpad, n_slots, st);
    return launch_vm_t<false, 64>(p, Lpad, n_slots, st);
}

static int pick_tpb(const Plan &plan) {
    int t = plan.opts.threads_per_block;
    if (t == 32 || t == 64 || t == 128) return t;
    return 64;
}

// Copy `count` doubles per plane from a caller pointer (host or device) into rows of a padded device plane array.
static int load_planes(double *dst, long long Lpad, const double *src, long long src_stride, long long lane0, long long lanes,
                       int planes, bool src_device, cudaStream_t st) {
    if (planes <= 0) return KRB_OK;
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)Lpad * 8, src + lane0, (size_t)src_stride * 8, (size_t)lanes * 8, (size_t)planes,
                               src_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    return KRB_OK;
}

int run_lanes(Plan *plan_first, Plan *plan_steps, const RunArgs &a) {
    if (device_count() <= 0) { set_error("no CUDA device: krets_b200 has no CPU fallback"); return KRB_ERR_NO_DEVICE; }
    Plan &P0 = *plan_first;
    Plan *PS = plan_steps;
    if (!P0.dev) P0.dev = device_state_new();
    if (PS && !PS->dev) PS->dev = device_state_new();
    DeviceState &D0 = *P0.dev;
    cudaStream_t st = (cudaStream_t)a.stream;
    const bool cplx = P0.is_complex;
    const size_t ts = cplx ? 16 : 8;
    const int n = P0.circuit.n;
    const int rows = a.rows;
    const int tpb = pick_tpb(P0);
    int rc;
    if ((rc = upload_programs(P0, D0))) return rc;
    if (PS && (rc = upload_programs(*PS, *PS->dev))) return rc;
    if (PS && PS->n_inputs != P0.n_inputs) { set_error("plan_dc and plan_tran must be analysed with the same varied elements"); return KRB_ERR_INVALID_ARGUMENT; }
    if (n == 0 || a.lanes <= 0) return KRB_OK;

    // ---- lane-invariant tables ----
    std::vector<double> pool;
    evaluate_pool(P0, P0.circuit, P0.opts.time_step, pool);
    if ((rc = D0.B.ensure(std::max<size_t>(pool.size(), 1) * 8))) return rc;
    if (!pool.empty()) CUDA_TRY(cudaMemcpyAsync(D0.B.p, pool.data(), pool.size() * 8, cudaMemcpyHostToDevice, st));
    std::vector<double> pool_s, bstep;
    if (PS) {
        DeviceState &DS = *PS->dev;
        evaluate_pool(*PS, PS->circuit, PS->opts.time_step, pool_s);
        if ((rc = DS.B.ensure(std::max<size_t>(pool_s.size(), 1) * 8))) return rc;
        if (!pool_s.empty()) CUDA_TRY(cudaMemcpyAsync(DS.B.p, pool_s.data(), pool_s.size() * 8, cudaMemcpyHostToDevice, st));
        // source table: row r, iteration 0 sees the time key of the previous row, later iterations see r*h
        const int nw = (int)PS->wave_elems.size();
        const double h = PS->opts.time_step;
        bstep.assign((size_t)rows * 2 * std::max(nw, 1), 0.0);
        for (int r = 1; r < rows; r++)
            for (int k = 0; k < nw; k++) {
                const krb_element &el = PS->circuit.el[PS->wave_elems[k]];
                bstep[((size_t)r * 2 + 0) * nw + k] = transient_value_at(el, (double)(r - 1) * h);   // stampable.rs:1043 + transient.rs:49
                bstep[((size_t)r * 2 + 1) * nw + k] = transient_value_at(el, (double)r * h);         // transient.rs:95,105
            }
        if ((rc = DS.Bstep.ensure(bstep.size() * 8))) return rc;
        CUDA_TRY(cudaMemcpyAsync(DS.Bstep.p, bstep.data(), bstep.size() * 8, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));   // host vectors above go out of scope

    // ---- chunking: bound the per-call device footprint ----
    const bool chain = P0.use_chain && !PS;
    const int planes0 = chain ? 2 * n + P0.chain.n_u2 : P0.n_planes, planesS = PS ? PS->n_planes : 0;
    const size_t per_lane = (size_t)(planes0 + planesS) * ts + (size_t)rows * n * ts * 2 + (size_t)P0.n_inputs * 8 + (a.op_solution ? (size_t)n * 8 : 0) + 16;
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    size_t reusable = D0.G.bytes + D0.X.bytes + D0.packed.bytes + D0.IN.bytes + (PS ? PS->dev->G.bytes : 0);
    long long max_chunk = (long long)(((free_b + reusable) * 0.80) / (double)per_lane);
    max_chunk = (max_chunk / tpb) * tpb;
    if (max_chunk < tpb) { set_error("not enough device memory for one block of lanes"); return KRB_ERR_OUT_OF_MEMORY; }
    long long chunk_cap = a.chunk_lanes > 0 ? a.chunk_lanes : 65536;
    chunk_cap = std::max<long long>(tpb, (chunk_cap / tpb) * tpb);
    const long long chunk = std::min<long long>(std::min(max_chunk, chunk_cap), ((a.lanes + tpb - 1) / tpb) * tpb);

    if ((rc = D0.IN.ensure(std::max<size_t>((size_t)P0.n_inputs, 1) * chunk * 8))) return rc;
    if ((rc = D0.G.ensure(std::max<size_t>((size_t)planes0, 1) * chunk * ts))) return rc;
    if (PS && (rc = PS->dev->G.ensure(std::max<size_t>((size_t)planesS, 1) * chunk * ts))) return rc;
    if ((rc = D0.X.ensure((size_t)rows * n * chunk * ts))) return rc;
    if ((rc = D0.status.ensure((size_t)chunk * 4))) return rc;
    if (a.iterations && (rc = D0.iters.ensure((size_t)chunk * 4))) return rc;
    if (a.op_solution && (rc = D0.XOP.ensure((size_t)n * chunk * 8))) return rc;
    const bool out_dev = (a.memory == KRB_MEM_DEVICE);
    const bool need_packed = !out_dev;   // host delivery goes through a packed device buffer
    if (need_packed && (rc = D0.packed.ensure((size_t)rows * n * chunk * ts))) return rc;

    D0.events_used = 0;
    g_last_timed = &D0;
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

    for (long long lane0 = 0; lane0 < a.lanes; lane0 += chunk) {
        const long long lanes = std::min<long long>(chunk, a.lanes - lane0);
        const long long Lpad = ((lanes + tpb - 1) / tpb) * tpb;
        double *IN = (double *)D0.IN.p;
        // inputs
        for (int k = 0; k < P0.n_inputs; k++) {
            const double *src = nullptr;
            long long stride = a.lanes;
            if (k == P0.input_sweep || k == P0.input_freq) src = a.lane_param;
            else {
                int vk = -1;
                for (size_t q = 0; q < P0.input_of_varied.size(); q++) if (P0.input_of_varied[q] == k) { vk = (int)q; break; }
                if (vk < 0 || !a.varied_values) { set_error("missing per-lane values for a varied element"); return KRB_ERR_INVALID_ARGUMENT; }
                src = a.varied_values + (size_t)vk * a.lanes;
            }
            if ((rc = load_planes(IN + (size_t)k * Lpad, Lpad, src, stride, lane0, lanes, 1, out_dev, st))) return rc;
        }
        if (Lpad > lanes && P0.n_inputs > 0) {
            const long long work = (long long)P0.n_inputs * (Lpad - lanes);
            krb_fill_pad<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(IN, Lpad, lanes, P0.n_inputs);
            g_launches++;
        }
        if (a.op_solution) {
            const double *src = a.op_solution;
            Buf tmp;
            const bool bc = (a.lanes_op == 1);
            const size_t cnt = bc ? (size_t)n : (size_t)lanes * n;
            const double *dsrc = src + (bc ? 0 : (size_t)lane0 * n);
            if (!out_dev) {
                if ((rc = tmp.ensure(cnt * 8))) return rc;
                CUDA_TRY(cudaMemcpyAsync(tmp.p, dsrc, cnt * 8, cudaMemcpyHostToDevice, st));
                dsrc = (const double *)tmp.p;
            }
            const long long work = (long long)n * Lpad;
            krb_spread_op<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(dsrc, (double *)D0.XOP.p, n, Lpad, lanes, bc ? 1 : 0);
            g_launches++;
            CUDA_TRY(cudaStreamSynchronize(st));
            tmp.release();
        }
        CUDA_TRY(cudaMemsetAsync(D0.status.p, 0, (size_t)Lpad * 4, st));

        VmParams p{};
        p.sec0 = D0.sec[0]; p.sec1 = D0.sec[1]; p.sec2 = D0.sec[2];
        if (P0.has_step_sections) p.sec0 = nullptr; else { p.sec1 = nullptr; p.sec2 = nullptr; }
        p.B = (const double *)D0.B.p; p.Bstep = nullptr; p.n_wave = 0;
        p.IN = IN; p.G = D0.G.p; p.X = D0.X.p;
        p.Xprev0 = a.op_solution ? (const double *)D0.XOP.p : nullptr;
        p.n = n; p.L = Lpad;
        p.row_begin = 0; p.row_end = PS ? 1 : rows;
        p.nonlinear = P0.circuit.nonlinear ? 1 : 0;
        p.max_iter = p.nonlinear ? std::max(1, a.cfg.maximum_iterations) : 1;
        if (cplx) { p.nonlinear = 0; p.max_iter = 1; }
        p.guess_empty_at_iter0 = 1; p.pseudo_time = 0;
        p.h = P0.opts.time_step; p.reltol = a.cfg.relative_tolerance; p.iabstol = a.cfg.current_absolute_tolerance; p.vabstol = a.cfg.voltage_absolute_tolerance;
        p.status = (int *)D0.status.p; p.iters = a.iterations ? (int *)D0.iters.p : nullptr;
        cudaEvent_t e0, e1;
        if ((rc = next_events(e0, e1))) return rc;
        CUDA_TRY(cudaEventRecord(e0, st));
        if (chain) {
            ChainParams c{};
            c.terms = D0.chain_terms; c.unknown_of_row = D0.chain_unknown; c.IN = IN; c.UY = D0.G.p; c.X = D0.X.p;
            c.n = n; c.input_freq = P0.input_freq; c.n_u2 = P0.chain.n_u2; c.u2_flags = D0.chain_u2_flags; c.L = Lpad; c.status = (int *)D0.status.p;
            if ((rc = launch_chain(cplx, tpb, c, Lpad, st))) return rc;
        } else if ((rc = launch_vm(cplx, tpb, p, Lpad, P0.n_slots, st))) return rc;
        if (PS) {
            DeviceState &DS = *PS->dev;
            VmParams q = p;
            q.sec0 = DS.sec[0]; q.sec1 = DS.sec[1]; q.sec2 = DS.sec[2];
            if (PS->has_step_sections) q.sec0 = nullptr; else { q.sec1 = nullptr; q.sec2 = nullptr; }
            q.B = (const double *)DS.B.p; q.Bstep = (const double *)DS.Bstep.p; q.n_wave = (int)PS->wave_elems.size();
            q.G = DS.G.p; q.Xprev0 = nullptr;
            q.row_begin = 1; q.row_end = rows;
            q.nonlinear = PS->circuit.nonlinear ? 1 : 0;
            q.max_iter = q.nonlinear ? std::max(1, a.cfg.maximum_iterations) : 1;
            q.guess_empty_at_iter0 = 0; q.pseudo_time = 1;
            q.h = PS->opts.time_step;
            q.iters = nullptr;
            if (rows > 1 && (rc = launch_vm(false, tpb, q, Lpad, PS->n_slots, st))) return rc;
        }
        CUDA_TRY(cudaEventRecord(e1, st));

        // ---- delivery ----
        const size_t per_point = (size_t)rows * n;
        if (a.layout == KRB_LAYOUT_POINT_MAJOR) {
            void *dst = out_dev ? (void *)((char *)a.x_out + (size_t)lane0 * per_point * ts) : D0.packed.p;
            dim3 grid((unsigned)((lanes + 31) / 32), (unsigned)((n + 31) / 32), (unsigned)rows), block(32, 8);
            if (cplx) krb_pack_point_major<double2><<<grid, block, 0, st>>>((const double2 *)D0.X.p, (double2 *)dst, rows, n, Lpad, lanes);
            else krb_pack_point_major<double><<<grid, block, 0, st>>>((const double *)D0.X.p, (double *)dst, rows, n, Lpad, lanes);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            if (!out_dev) CUDA_TRY(cudaMemcpyAsync((char *)a.x_out + (size_t)lane0 * per_point * ts, D0.packed.p, (size_t)lanes * per_point * ts, cudaMemcpyDeviceToHost, st));
        } else {
            // lane-interleaved delivery: x_out[row][u][lane] with the caller's lane stride (= total lanes)
            CUDA_TRY(cudaMemcpy2DAsync((char *)a.x_out + (size_t)lane0 * ts, (size_t)a.lanes * ts, D0.X.p, (size_t)Lpad * ts, (size_t)lanes * ts, per_point,
                                       out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        }
        if (a.lane_status)
            CUDA_TRY(cudaMemcpyAsync(a.lane_status + lane0, D0.status.p, (size_t)lanes * 4, out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        if (a.iterations)
            CUDA_TRY(cudaMemcpyAsync(a.iterations + lane0, D0.iters.p, (size_t)lanes * 4, out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        if (!out_dev || lane0 + chunk < a.lanes) CUDA_TRY(cudaStreamSynchronize(st));   // staging buffers are reused by the next chunk
    }
    return KRB_OK;
}

}  // namespace krb
