// sm_100a lane-program interpreter: one thread = one solve lane, one warp = 32 consecutive lanes.
//
// Memory behaviour (the design point for B200's HBM3e):
//   * every per-lane array is lane-interleaved  P[idx][lane]  -> each op's global access is one fully coalesced
//     256 B (f64) / 512 B (c64, 128-bit per thread) segment per warp; nothing is ever gathered or scattered;
//   * working values live in a per-thread shared-memory column scr[slot][tid] (bank-conflict free); the host
//     compiler decides what stays on chip (Belady) so HBM only sees spilled U^/y^ values, once out and once back;
//   * the op stream is lane-invariant: a warp fetches 32 ops with one coalesced 256 B load (prefetched one chunk
//     ahead) and broadcasts them with shuffles; no divergence, no per-lane control flow except the NR mask.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/krets_b200.h"
#include "scalar.cuh"
#include "vm_ops.h"

namespace krb {

struct VmParams {
    const uint2 *sec0, *sec1, *sec2;   // programs (sec1 = factor-once, sec2 = per-step; null when unused)
    const double *B;                   // lane-invariant constants (static part)
    const double *Bstep;               // per-row source table: [row][2][n_wave]
    int n_wave;
    const double *IN;                  // [n_inputs][L]
    void *G;                           // [n_planes][L] of T
    void *X;                           // [rows][n][L] of T   (result rows, lane-interleaved)
    const double *Xprev0;              // real plane [n][L] used as "previous row" of row_begin (AC: OP solution); may be null
    int n;
    int row_width;                     // columns per result row: n unknowns + per-lane state columns (trapezoidal capacitor currents)
    long long L;                       // lane stride of every plane (padded lane count)
    int row_begin, row_end;
    int nonlinear, max_iter;
    int guess_empty_at_iter0;          // OP/DC: the first NR iteration sees an empty solution map (op.rs:30-31)
    int pseudo_time;                   // TRAN: the "time" key takes part in the convergence check
    double h, reltol, iabstol, vabstol;
    int *status;                       // [L] or null
    int *iters;                        // [L] or null
};

// diode.rs:42-70
__device__ __forceinline__ double diode_limit(double vd, double vcrit) { return fmin(fmax(vd, -vcrit), vcrit); }

// level-1 NMOS, nmosfet.rs:51-102 (extension: reached only through KRB_ELEM_FLAG_LEVEL1).  what: 0 g_m, 1 g_ds, 2 i_d.  The region
// tests and the evaluation order of every product are the reference's; single operations are rounded separately (no contraction).
__device__ __forceinline__ double nmos_level1(int what, double vgs, double vds, double vth, double beta, double lambda) {
    if (vgs <= vth) return 0.0;
    const double vov = __dsub_rn(vgs, vth);
    if (vds >= 0.0 && vds <= vov) {          // linear region
        if (what == 0) return __dmul_rn(beta, vds);
        if (what == 1) return __dmul_rn(beta, __dsub_rn(vov, vds));
        return __dmul_rn(beta, __dsub_rn(__dmul_rn(vov, vds), __ddiv_rn(__dmul_rn(vds, vds), 2.0)));
    }
    if (vds >= vov && vds >= 0.0) {          // saturation region
        const double clm = __dadd_rn(1.0, __dmul_rn(lambda, vds));
        if (what == 0) return __dmul_rn(__dmul_rn(beta, vov), clm);
        if (what == 1) return __dmul_rn(__dmul_rn(__ddiv_rn(beta, 2.0), lambda), __dmul_rn(vov, vov));
        return __dmul_rn(__dmul_rn(__ddiv_rn(beta, 2.0), __dmul_rn(vov, vov)), clm);
    }
    return 0.0;
}

// ---- shared-memory scratch access with 32-bit addresses (no generic-pointer arithmetic in the interpreter loop) ----
template <bool CPLX> struct ScrIO;
template <> struct ScrIO<false> {
    static __device__ __forceinline__ double ld(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
    static __device__ __forceinline__ void st(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
};
template <> struct ScrIO<true> {
    static __device__ __forceinline__ double2 ld(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory"); return v; }
    static __device__ __forceinline__ void st(uint32_t a, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory"); }
};

// asynchronous global -> shared copies straight into a scratch slot (LDGSTS); completion via KOP_WAIT
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void *g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <bool CPLX> __device__ __forceinline__ void async_real_to_slot(uint32_t saddr, const double *g) {
    cp_async8(saddr, g);
    if (CPLX) asm volatile("st.shared.f64 [%0], %1;" ::"r"(saddr + 8u), "d"(0.0) : "memory");   // (v, 0)
}

__device__ __forceinline__ double vm_rcp_seed(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
// 1/x to ~1 ulp: r0 (1 + e + e^2) with e = 1 - x r0 (MUFU seed error ~2^-20 -> e^3 < 2^-60); pivots are not bit-pinned
__device__ __forceinline__ double vm_fast_rcp(double x) {
    const double r0 = vm_rcp_seed(x);
    const double e = fma(-x, r0, 1.0);
    const double t = fma(e, e, e);
    return fma(r0, t, r0);
}
template <bool CPLX> struct VmRcp;
template <> struct VmRcp<false> {
    static __device__ __forceinline__ double rcp(double a, bool &bad) { bad = !(fabs(a) > 0.0) || !isfinite(a); return vm_fast_rcp(a); }
};
template <> struct VmRcp<true> {
    static __device__ __forceinline__ double2 rcp(double2 a, bool &bad) {
        const double ns = fma(a.x, a.x, a.y * a.y);
        bad = !(ns > 0.0) || !isfinite(ns);
        const double r = vm_fast_rcp(ns);
        return make_double2(a.x * r, -a.y * r);
    }
};

template <bool CPLX, int TPB>
__global__ void __launch_bounds__(TPB) krb_lane_vm(const VmParams p) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    typedef ScrIO<CPLX> IO;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t SLOT_SHIFT = (CPLX ? 4 : 3) + (TPB == 32 ? 5 : TPB == 64 ? 6 : 7);   // log2(TPB * sizeof(T))
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw) + threadIdx.x * (uint32_t)sizeof(T);   // slot s at sbase + (s << SLOT_SHIFT)
    const unsigned lane_id = threadIdx.x & 31u;
    const long long lane = (long long)blockIdx.x * TPB + threadIdx.x;
    const long long L = p.L;
    const double *__restrict__ IN = p.IN + lane;
    T *__restrict__ G = reinterpret_cast<T *>(p.G) + lane;
    const double *__restrict__ Bc = p.B;
    int status = 0;
    int it_used = 0;
    uint32_t grow = 0;   // running maximum of growth_key(u^) over every factorisation of the lane (scalar.cuh)

    for (int row = p.row_begin; row < p.row_end; ++row) {
        T *Xcur = reinterpret_cast<T *>(p.X) + (size_t)row * (size_t)p.row_width * (size_t)L + lane;
        const double *Xprev;
        if (row == p.row_begin && p.Xprev0) Xprev = p.Xprev0 + lane;
        else Xprev = reinterpret_cast<const double *>(reinterpret_cast<T *>(p.X) + (size_t)(row > 0 ? row - 1 : 0) * (size_t)p.row_width * (size_t)L) + lane;
        bool lane_active = true;
        const int n_pass = (p.sec1 && row == p.row_begin) ? 2 : 1;
        for (int pass = 0; pass < n_pass; ++pass) {
            const uint2 *__restrict__ prog = p.sec1 ? ((n_pass == 2 && pass == 0) ? p.sec1 : p.sec2) : p.sec0;
            const bool solve_pass = !(n_pass == 2 && pass == 0);
            for (int iter = 0; iter < p.max_iter; ++iter) {
                const bool guess_empty = p.guess_empty_at_iter0 && iter == 0;
                const double *Xguess = (iter == 0 && !p.guess_empty_at_iter0) ? Xprev : reinterpret_cast<const double *>(Xcur);
                const double *Bs = p.Bstep + ((size_t)row * 2 + (iter > 0 ? 1 : 0)) * (size_t)p.n_wave;
                bool notconv = false;
                // ---- run one program: 32-op chunks, one coalesced load per chunk (prefetched one chunk ahead), ops broadcast
                // by shuffle; the next op's words are fetched while the current one executes ----
                int pc = 0;
                uint2 cur = prog[lane_id];
                bool running = true;
                while (running) {
                    const uint2 nxt = prog[pc + 32 + lane_id];   // programs are padded with 64 END ops
                    uint32_t w0 = __shfl_sync(0xffffffffu, cur.x, 0);
                    uint32_t imm = __shfl_sync(0xffffffffu, cur.y, 0);
                    int t = 0;
#pragma unroll 1
                    while (t < 32) {
                        const int tn = t < 31 ? t + 1 : 31;
                        const uint32_t nw0 = __shfl_sync(0xffffffffu, cur.x, tn);     // next op (or the inline constant of an LDK)
                        const uint32_t nimm = __shfl_sync(0xffffffffu, cur.y, tn);
                        const uint32_t opc = w0 & 0xffu;
                        const uint32_t ad = sbase + (((w0 >> 8) & 0xffu) << SLOT_SHIFT);
                        const uint32_t aa = sbase + (((w0 >> 16) & 0xffu) << SLOT_SHIFT);
                        const uint32_t bf = w0 >> 24;
                        const uint32_t ab = sbase + (bf << SLOT_SHIFT);
                        int step = 1;
                        if (opc == KOP_FNMA) IO::st(ad, S::fnma(IO::ld(ad), IO::ld(aa), IO::ld(ab)));
                        else if (opc == KOP_LDK) { IO::st(ad, S::from_real(__hiloint2double((int)nimm, (int)nw0))); step = 2; }
                        else if (opc == KOP_MUL) IO::st(ad, S::mul(IO::ld(aa), IO::ld(ab)));
                        else if (opc == KOP_ADD) IO::st(ad, S::add(IO::ld(aa), IO::ld(ab)));
                        else if (opc == KOP_STG) G[(size_t)imm * L] = IO::ld(aa);
                        else if (opc == KOP_LDG) { if (CPLX) cp_async16(ad, G + (size_t)imm * L); else cp_async8(ad, G + (size_t)imm * L); }
                        else if (opc == KOP_WAIT) cp_async_wait_all();
                        else if (opc == KOP_END) { running = false; break; }
                        else switch (opc) {
                        case KOP_LDB: IO::st(ad, S::from_real(__ldg(Bc + imm))); break;
                        case KOP_LDBS: async_real_to_slot<CPLX>(ad, Bs + imm); break;
                        case KOP_LDI: async_real_to_slot<CPLX>(ad, IN + (size_t)imm * L); break;
                        case KOP_LDX: {
                            if ((bf & 1u) && guess_empty) IO::st(ad, S::from_real((bf & 2u) ? 0.5 : 0.0));   // empty solution map defaults
                            else async_real_to_slot<CPLX>(ad, ((bf & 1u) ? Xguess : Xprev) + (size_t)imm * L);
                            break;
                        }
                        case KOP_XOUT: {
                            const T v = IO::ld(aa);
                            if (!S::finite(v)) status |= KRB_LANE_NONFINITE;
                            if (!CPLX && p.nonlinear && !guess_empty && !(bf & 4u)) {
                                const double nv = S::re(v);
                                const double ov = Xguess[(size_t)imm * L];
                                const double diff = fabs(nv - ov);
                                const double scale = fmax(fabs(nv), fabs(ov));
                                const double atol = (bf & 1u) ? p.iabstol : p.vabstol;
                                if (!(diff <= p.reltol * scale + atol)) notconv = true;   // solver.rs:89-103
                            }
                            if (lane_active) Xcur[(size_t)imm * L] = v;
                            break;
                        }
                        case KOP_ZERO: IO::st(ad, S::zero()); break;
                        case KOP_MOV: IO::st(ad, IO::ld(aa)); break;
                        case KOP_NEG: IO::st(ad, S::neg(IO::ld(aa))); break;
                        case KOP_SUB: IO::st(ad, S::sub(IO::ld(aa), IO::ld(ab))); break;
                        case KOP_RCP: {
                            bool bad;
                            IO::st(ad, VmRcp<CPLX>::rcp(IO::ld(aa), bad));
                            if (bad) status |= KRB_LANE_ZERO_PIVOT;
                            break;
                        }
                        case KOP_MULG: {
                            const T v = S::mul(IO::ld(aa), IO::ld(ab));
                            IO::st(ad, v);
                            grow = max(grow, growth_key(v));
                            break;
                        }
                        case KOP_INV: IO::st(ad, S::from_real(__ddiv_rn(1.0, S::re(IO::ld(aa))))); break;
                        case KOP_DIVR: IO::st(ad, S::from_real(__ddiv_rn(S::re(IO::ld(aa)), S::re(IO::ld(ab))))); break;
                        case KOP_MULJ: IO::st(ad, S::mulj(IO::ld(aa))); break;
                        case KOP_DIOG: {   // (is/(n Vt)) * exp(vlim/(n Vt))
                            const double vd = S::re(IO::ld(aa));
                            const double vl = diode_limit(vd, __ldg(Bc + imm + 3));
                            IO::st(ad, S::from_real(__ldg(Bc + imm) * exp(vl / __ldg(Bc + imm + 1))));
                            break;
                        }
                        case KOP_DIOI: {   // is*(exp(vlim/(n Vt)) - 1) - g*vd
                            const double vd = S::re(IO::ld(aa));
                            const double vl = diode_limit(vd, __ldg(Bc + imm + 3));
                            const double ex = exp(vl / __ldg(Bc + imm + 1));
                            const double g = __ldg(Bc + imm) * ex;
                            const double id = __ldg(Bc + imm + 2) * (ex - 1.0);
                            IO::st(ad, S::from_real(__dsub_rn(id, __dmul_rn(g, vd))));
                            break;
                        }
                        case KOP_MOSGM:
                        case KOP_MOSGDS:
                        case KOP_MOSID:
                            IO::st(ad, S::from_real(nmos_level1((int)opc - (int)KOP_MOSGM, S::re(IO::ld(aa)), S::re(IO::ld(ab)), __ldg(Bc + imm),
                                                                __ldg(Bc + imm + 1), __ldg(Bc + imm + 2))));
                            break;
                        default: break;   // KOP_NOP
                        }
                        t += step;
                        if (step == 1) { w0 = nw0; imm = nimm; }
                        else if (t < 32) { w0 = __shfl_sync(0xffffffffu, cur.x, t); imm = __shfl_sync(0xffffffffu, cur.y, t); }
                    }
                    cur = nxt;
                    pc += 32;
                }
                // ---- NR bookkeeping (op.rs:65-81, dc.rs:99-110, transient.rs:97-109) ----
                if (!solve_pass || !p.nonlinear) break;
                bool pseudo_ok = true;
                if (p.pseudo_time && iter == 0) {   // |t_n - t_{n-1}| compared as a voltage
                    const double tn = (double)row * p.h, tp = (double)(row - 1) * p.h;
                    pseudo_ok = fabs(tn - tp) <= p.reltol * fmax(fabs(tn), fabs(tp)) + p.vabstol;
                }
                if (lane_active) it_used = iter + 1;
                const bool conv = !guess_empty && !notconv && pseudo_ok;
                if (lane_active && conv) lane_active = false;
                if (lane_active && iter == p.max_iter - 1) status |= KRB_LANE_NOT_CONVERGED;
                if (__all_sync(0xffffffffu, !lane_active)) break;
            }
        }
    }
    // linear circuits only: a diode conductance spans tens of decades over the Newton iterates, so |u^| says nothing about a
    // nonlinear lane; there the convergence check (solver.rs:76-104) on the next iterate is what guards the answer
    if (!p.nonlinear && grow >= KRB_GROWTH_KEY && grow < KRB_NONFINITE_KEY) status |= KRB_LANE_INACCURATE;
    if (p.status) p.status[lane] |= status;
    if (p.iters) p.iters[lane] = it_used;
}

// x_out[lane][row][u] (point-major, what the reference returns per solved point) <- X[row][u][lane]
// 32x32 tiles through shared memory so both sides are coalesced.
template <typename T>
__global__ void krb_pack_point_major(const T *__restrict__ X, T *__restrict__ out, int rows, int n, int width, long long L, long long lanes) {
    __shared__ T tile[32][33];
    const long long lane0 = (long long)blockIdx.x * 32;
    const int u0 = blockIdx.y * 32;
    const int row = blockIdx.z;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int u = u0 + j;
        const long long ln = lane0 + threadIdx.x;
        if (u < n && ln < lanes) tile[j][threadIdx.x] = X[((size_t)row * width + u) * (size_t)L + ln];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long ln = lane0 + j;
        const int u = u0 + threadIdx.x;
        if (u < n && ln < lanes) out[((size_t)ln * rows + row) * (size_t)n + u] = tile[threadIdx.x][j];
    }
}

// lane-interleaved copy with a different lane stride (padded -> caller's)
template <typename T>
__global__ void krb_copy_planes(const T *__restrict__ X, T *__restrict__ out, long long planes, long long L, long long lanes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = planes * lanes;
    if (i < total) {
        const long long pl = i / lanes, ln = i % lanes;
        out[pl * lanes + ln] = X[pl * L + ln];
    }
}

// IN[k][lane] <- replicate / pad helpers
__global__ void krb_fill_pad(double *plane, long long L, long long lanes, long long planes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long padn = L - lanes;
    if (padn <= 0) return;
    if (i < planes * padn) {
        const long long pl = i / padn, k = i % padn;
        plane[pl * L + lanes + k] = plane[pl * L];   // padded lanes repeat lane 0 (their results are never delivered)
    }
}

// one input plane <- the same value on every lane (sequential DC sweep: the sweep element's value of the current row)
__global__ void krb_fill_plane(double *__restrict__ plane, long long L, double value) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L) plane[i] = value;
}

// op_solution[lane][u] (point-major) or [u] broadcast  ->  plane [u][L]
__global__ void krb_spread_op(const double *__restrict__ op, double *__restrict__ plane, int n, long long L, long long lanes, int broadcast) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)n * L) {
        const long long u = i / L;
        long long ln = i % L;
        if (ln >= lanes) ln = 0;
        plane[i] = broadcast ? op[u] : op[ln * n + u];
    }
}

}  // namespace krb
