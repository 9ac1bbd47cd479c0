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

struct c64 { double re, im; };

template <bool CPLX> struct Sc;
template <> struct Sc<false> {
    typedef double T;
    static __device__ __forceinline__ T from_real(double v) { return v; }
    static __device__ __forceinline__ double re(T v) { return v; }
    static __device__ __forceinline__ T zero() { return 0.0; }
    static __device__ __forceinline__ T neg(T a) { return -a; }
    static __device__ __forceinline__ T add(T a, T b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ T sub(T a, T b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ T mul(T a, T b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ T fnma(T d, T a, T b) { return fma(-a, b, d); }
    static __device__ __forceinline__ T rcp(T a, bool &bad) { bad = !(fabs(a) > 0.0) || !isfinite(a); return 1.0 / a; }
    static __device__ __forceinline__ T mulj(T a) { return a; }
    static __device__ __forceinline__ bool finite(T a) { return isfinite(a); }
};
template <> struct Sc<true> {
    typedef double2 T;
    static __device__ __forceinline__ T from_real(double v) { return make_double2(v, 0.0); }
    static __device__ __forceinline__ double re(T v) { return v.x; }
    static __device__ __forceinline__ T zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ T neg(T a) { return make_double2(-a.x, -a.y); }
    static __device__ __forceinline__ T add(T a, T b) { return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
    static __device__ __forceinline__ T sub(T a, T b) { return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y)); }
    // num_complex-style products (no inf/nan recovery), exact for (x,0)*(y,0)
    static __device__ __forceinline__ T mul(T a, T b) {
        return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
    }
    static __device__ __forceinline__ T fnma(T d, T a, T b) {
        return make_double2(fma(a.y, b.y, fma(-a.x, b.x, d.x)), fma(-a.y, b.x, fma(-a.x, b.y, d.y)));
    }
    static __device__ __forceinline__ T rcp(T a, bool &bad) {
        const double ns = fma(a.x, a.x, a.y * a.y);
        bad = !(ns > 0.0) || !isfinite(ns);
        const double r = 1.0 / ns;
        return make_double2(a.x * r, -a.y * r);
    }
    static __device__ __forceinline__ T mulj(T a) { return make_double2(0.0, a.x); }
    static __device__ __forceinline__ bool finite(T a) { return isfinite(a.x) && isfinite(a.y); }
};

// diode.rs:42-70
__device__ __forceinline__ double diode_limit(double vd, double vcrit) { return fmin(fmax(vd, -vcrit), vcrit); }

template <bool CPLX, int TPB>
__global__ void __launch_bounds__(TPB) krb_lane_vm(const VmParams p) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scr = reinterpret_cast<T *>(smem_raw) + threadIdx.x;   // slot s lives at scr[s * TPB]
    const unsigned lane_id = threadIdx.x & 31u;
    const long long lane = (long long)blockIdx.x * TPB + threadIdx.x;
    const long long L = p.L;
    const double *__restrict__ IN = p.IN + lane;
    T *__restrict__ G = reinterpret_cast<T *>(p.G) + lane;
    const double *__restrict__ Bc = p.B;
    int status = 0;
    int it_used = 0;

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
                // ---- run one program ----
                int pc = 0;
                uint2 cur = prog[lane_id];
                bool running = true;
                while (running) {
                    const uint2 nxt = prog[pc + 32 + lane_id];   // programs are padded with 64 END ops
#pragma unroll 1
                    for (int t = 0; t < 32; ++t) {
                        const uint32_t w0 = __shfl_sync(0xffffffffu, cur.x, t);
                        const uint32_t imm = __shfl_sync(0xffffffffu, cur.y, t);
                        const uint32_t opc = w0 & 0xffu;
                        if (opc == KOP_END) { running = false; break; }
                        T *pd = scr + ((w0 >> 8) & 0xffu) * TPB;
                        const T *pa = scr + ((w0 >> 16) & 0xffu) * TPB;
                        const uint32_t bf = w0 >> 24;
                        const T *pb = scr + bf * TPB;
                        switch (opc) {
                        case KOP_LDB: *pd = S::from_real(__ldg(Bc + imm)); break;
                        case KOP_LDBS: *pd = S::from_real(__ldg(Bs + imm)); break;
                        case KOP_LDI: *pd = S::from_real(IN[(size_t)imm * L]); break;
                        case KOP_LDG: *pd = G[(size_t)imm * L]; break;
                        case KOP_STG: G[(size_t)imm * L] = *pa; break;
                        case KOP_LDX: {
                            double v;
                            if (bf & 1u) v = guess_empty ? ((bf & 2u) ? 0.5 : 0.0) : Xguess[(size_t)imm * L];
                            else v = Xprev[(size_t)imm * L];
                            *pd = S::from_real(v);
                            break;
                        }
                        case KOP_XOUT: {
                            const T v = *pa;
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
                        case KOP_ZERO: *pd = S::zero(); break;
                        case KOP_MOV: *pd = *pa; break;
                        case KOP_NEG: *pd = S::neg(*pa); break;
                        case KOP_ADD: *pd = S::add(*pa, *pb); break;
                        case KOP_SUB: *pd = S::sub(*pa, *pb); break;
                        case KOP_MUL: *pd = S::mul(*pa, *pb); break;
                        case KOP_FNMA: *pd = S::fnma(*pd, *pa, *pb); break;
                        case KOP_RCP: {
                            bool bad;
                            *pd = S::rcp(*pa, bad);
                            if (bad) status |= KRB_LANE_ZERO_PIVOT;
                            break;
                        }
                        case KOP_INV: *pd = S::from_real(__ddiv_rn(1.0, S::re(*pa))); break;
                        case KOP_DIVR: *pd = S::from_real(__ddiv_rn(S::re(*pa), S::re(*pb))); break;
                        case KOP_MULJ: *pd = S::mulj(*pa); break;
                        case KOP_DIOG: {   // (is/(n Vt)) * exp(vlim/(n Vt))
                            const double vd = S::re(*pa);
                            const double vl = diode_limit(vd, __ldg(Bc + imm + 3));
                            *pd = S::from_real(__ldg(Bc + imm) * exp(vl / __ldg(Bc + imm + 1)));
                            break;
                        }
                        case KOP_DIOI: {   // is*(exp(vlim/(n Vt)) - 1) - g*vd
                            const double vd = S::re(*pa);
                            const double vl = diode_limit(vd, __ldg(Bc + imm + 3));
                            const double ex = exp(vl / __ldg(Bc + imm + 1));
                            const double g = __ldg(Bc + imm) * ex;
                            const double id = __ldg(Bc + imm + 2) * (ex - 1.0);
                            *pd = S::from_real(__dsub_rn(id, __dmul_rn(g, vd)));
                            break;
                        }
                        default: break;
                        }
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
