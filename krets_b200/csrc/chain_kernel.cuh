// "Chain" executor: plans whose permuted matrix has lower bandwidth <= 1 and upper bandwidth <= 2 (ladders, RC lines
// and other path-like circuits, including the source row that couples two neighbours).
//
// One thread = one lane.  The forward sweep stamps row k from a lane-invariant ROW STREAM, eliminates it against the
// previous row kept in registers (Thomas-style recurrence on the LDU form  u^_k = sup_k / w_k,  y^_k = y_k / w_k) and
// streams u^_k, y^_k to lane-interleaved HBM planes; the backward sweep re-reads them in reverse through a
// double-buffered register prefetch ring and delivers x.  Per point the kernel moves 2n values out, 2n back and n
// results — a third of the 3-pass stamp/factor/solve traffic of SURVEY §8(d).
//
// Row stream (built by the host, compiler.cpp::build_chain_plan), 16-byte units, one broadcast load per warp:
//   FIXED row   : every entry is one folded lane-invariant constant (+ at most one j*w*K term)  ->  the record is just
//                 {sub, diag, sup1, rhs} (real: 2 units) or {re[4], K[4]} (complex: 4 units): no per-term decoding;
//   GENERIC row : a run of 16-byte terms (target, kind, input plane, K) for Monte-Carlo values, swept sources, rows
//                 with a second super-diagonal, several reactive elements on one node, ...
// Both keep the reference's left-to-right summation order of duplicate stamps (solver.rs:59-70).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/krets_b200.h"
#include "scalar.cuh"

namespace krb {

enum : uint32_t {
    CT_SUB = 0, CT_DIAG = 1, CT_SUP1 = 2, CT_SUP2 = 3, CT_RHS = 4,
    CK_CONST = 0,       // K                               (lane-invariant stamps, folded in stamp order on the host)
    CK_IN = 1,          // IN[plane]                       (swept / Monte-Carlo source value)
    CK_INV_IN = 2,      // 1.0 / IN[plane]                 (Monte-Carlo resistor, stampable.rs:232)
    CK_IN_OVER_K = 3,   // IN[plane] / K                   (Monte-Carlo C or L over h)
    CK_JW_K = 4,        // (0, ((2 pi) f) K)               (C / L admittance, stampable.rs:379-382, 715)
    CK_JW_IN = 5,       // (0, ((2 pi) f) IN[plane])
    CT_NEG = 1u << 8,
    // row info word (one uint32 per row)
    RI_FIXED = 1u,            // bit 0: FIXED record, else GENERIC
    RI_HAS_U2 = 2u,           // bit 1: the row owns a u2 plane (second super-diagonal)
    RI_NTERMS_SHIFT = 8       // GENERIC: number of terms
};

struct ChainParams {
    const uint32_t *rowinfo;     // [n + ...] per-row info (GENERIC rows: #terms, HAS_U2)
    const uint32_t *runs;        // run table: (count << 1) | is_fixed
    int n_runs;
    const uint4 *stream;         // row records, 16-byte units, padded with 4 zero units
    const int *unknown_of_row;   // pivot row k delivers unknown unknown_of_row[k]
    const double *IN;            // [n_inputs][L]
    void *UY;                    // [2n + n_u2][L] of T: plane 2k = u^_{k,k+1}, 2k+1 = y^_k, 2n + r = u^_{k,k+2} of the r-th flagged row
    void *X;                     // [n][L] of T (staging for the pack kernel / lane-interleaved delivery), or
    void *OUT;                   // point-major x_out[lane][unknown] written directly (when the pivot order is the unknown order)
    long long out_lanes;         // lanes actually delivered to OUT (lanes >= out_lanes are padding)
    const int *u2_rows;          // rows owning a u2 plane, ascending (n_u2 entries)
    int n;
    int input_freq;              // input plane holding the AC frequency, or -1
    int n_u2;
    long long L;                 // lane stride of every plane (padded lane count)
    int *status;
    int pf_units;                // L1 prefetch distance of the row stream, in 16-byte units (0 = off)
    int debug;                   // profiling only: 1 = no forward stores, 2 = skip backward, 4 = skip forward
};

__device__ __forceinline__ double rcp_seed(double x) {   // MUFU.RCP64H: ~2^-20 relative error, no slow path
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
// 1/x to ~1 ulp with a three-deep dependent chain: r0 (1 + e + e^2), e = 1 - x r0, error e^3 <= 2^-60.
// (The LU is not bit-pinned; stamps use exact IEEE division.)
__device__ __forceinline__ double fast_rcp(double x) {
    const double r0 = rcp_seed(x);
    const double e = fma(-x, r0, 1.0);
    const double t = fma(e, e, e);
    return fma(r0, t, r0);
}

template <bool CPLX> struct ChainSc;
template <> struct ChainSc<false> {
    typedef double T;
    static __device__ __forceinline__ T zero() { return 0.0; }
    static __device__ __forceinline__ T make(double re, double) { return re; }
    static __device__ __forceinline__ T fnma(T d, T a, T b) { return fma(-a, b, d); }
    static __device__ __forceinline__ T mul(T a, T b) { return a * b; }
    static __device__ __forceinline__ bool finite(T a) { return isfinite(a); }
    static __device__ __forceinline__ void add(T &acc, double re, double) { acc = __dadd_rn(acc, re); }
    // (a, b, c) / w with the quotients folded into the reciprocal's correction step: critical path = seed + 3 fma
    static __device__ __forceinline__ bool scale3(T w, T a, T b, T c, T &qa, T &qb, T &qc) {
        const double r0 = rcp_seed(w);
        const double e = fma(-w, r0, 1.0);
        const double t = fma(e, e, e);
        const double a0 = a * r0, b0 = b * r0, c0 = c * r0;
        qa = fma(a0, t, a0); qb = fma(b0, t, b0); qc = fma(c0, t, c0);
        return !(fabs(w) > 0.0) || !isfinite(w);
    }
};
template <> struct ChainSc<true> {
    typedef double2 T;
    static __device__ __forceinline__ T zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ T make(double re, double im) { return make_double2(re, im); }
    static __device__ __forceinline__ T fnma(T d, T a, T b) {
        return make_double2(fma(a.y, b.y, fma(-a.x, b.x, d.x)), fma(-a.y, b.x, fma(-a.x, b.y, d.y)));
    }
    static __device__ __forceinline__ T mul(T a, T b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
    static __device__ __forceinline__ bool finite(T a) { return isfinite(a.x) && isfinite(a.y); }
    static __device__ __forceinline__ void add(T &acc, double re, double im) { acc.x = __dadd_rn(acc.x, re); acc.y = __dadd_rn(acc.y, im); }
    static __device__ __forceinline__ bool scale3(T w, T a, T b, T c, T &qa, T &qb, T &qc) {
        const double ns = fma(w.x, w.x, w.y * w.y);
        const double r = fast_rcp(ns);
        const T inv = make_double2(w.x * r, -w.y * r);
        qa = mul(a, inv); qb = mul(b, inv); qc = mul(c, inv);
        return !(ns > 0.0) || !isfinite(ns);
    }
};

__device__ __forceinline__ double u4lo(const uint4 &v) { return __hiloint2double((int)v.y, (int)v.x); }
__device__ __forceinline__ double u4hi(const uint4 &v) { return __hiloint2double((int)v.w, (int)v.z); }

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// DIRECT: deliver point-major results straight from the backward sweep (pivot row k == unknown k), staging 32 lanes x R
// rows per warp through shared memory so that every store instruction writes 64-byte runs of the caller's rows.
template <bool CPLX, int TPB, bool DIRECT>
__global__ void __launch_bounds__(TPB, (TPB == 32 ? 14 : TPB == 64 ? 7 : 3)) krb_chain_kernel(const ChainParams p) {
    typedef ChainSc<CPLX> S;
    typedef typename S::T T;
    constexpr int FIX_UNITS = CPLX ? 4 : 2;
    const long long lane = (long long)blockIdx.x * TPB + threadIdx.x;
    const long long L = p.L;
    const double *__restrict__ IN = p.IN + lane;
    const int n = p.n;
    // lane-interleaved planes P[idx][lane]: at any moment the whole grid writes (forward) or reads (backward) the same
    // few planes, i.e. the GPU walks HBM sequentially (measured faster than per-block tiling of the planes)
    T *__restrict__ UY = reinterpret_cast<T *>(p.UY) + lane;
    T *__restrict__ X = reinterpret_cast<T *>(p.X) + lane;
    int status = 0;
    uint32_t grow = 0;                                // running maximum of growth_key(u^)

    double w = 0.0;                                   // (2.0 * PI) * f, rounded once as the reference does
    if (CPLX && p.input_freq >= 0) w = __dmul_rn(2.0 * 3.14159265358979323846264338327950288, IN[(size_t)p.input_freq * L]);

    // ---- forward: stamp + eliminate + forward-substitute, row by row --------------------------------------------
    // The host groups the rows into runs: a FIXED run is a tight loop over fixed-shape records (one-ahead register
    // prefetch, no per-row decoding); GENERIC rows (swept source, Monte-Carlo terms, second super-diagonal) go one by one.
    const uint4 *__restrict__ dp = p.stream;
    T pu1 = S::zero(), pu2 = S::zero(), py = S::zero();   // u^_{k-1,k}, u^_{k-1,k+1}, y^_{k-1}
    int r2 = 0;                                         // rank of the next row that owns a u2 plane (warp-uniform)
    T *__restrict__ up = UY;                            // &UY[2k][lane]
    const size_t up_step = (size_t)2 * (size_t)L;
    auto finish_row = [&](T e0, T e1, T e2, T e3, T e4, bool has_u2) {
        const T wk = S::fnma(e1, e0, pu1);            // w_k     = diag - sub * u^_{k-1,k}
        const T s1 = S::fnma(e2, e0, pu2);            // sup1'   = sup1 - sub * u^_{k-1,k+1}
        const T yk = S::fnma(e4, e0, py);             // y_k     = rhs  - sub * y^_{k-1}
        if (S::scale3(wk, s1, e3, yk, pu1, pu2, py)) status |= KRB_LANE_ZERO_PIVOT;
        grow = max(grow, growth_key(pu1));
        __stcs(up, pu1);
        __stcs(up + L, py);
        up += up_step;
        if (has_u2) { __stcs(&UY[(size_t)(2 * n + r2) * L], pu2); ++r2; grow = max(grow, growth_key(pu2)); }
    };
    int k = 0;
    const int n_runs = (p.debug & 4) ? 0 : p.n_runs;
    for (int run = 0; run < n_runs; ++run) {
        const uint32_t rd = __ldg(p.runs + run);
        const int cnt = (int)(rd >> 1);
        if (rd & 1u) {
            uint4 q[FIX_UNITS];
#pragma unroll
            for (int i = 0; i < FIX_UNITS; ++i) q[i] = __ldg(dp + i);
#pragma unroll 2
            for (int r = 0; r < cnt; ++r) {
                uint4 c[FIX_UNITS];
#pragma unroll
                for (int i = 0; i < FIX_UNITS; ++i) c[i] = q[i];
                dp += FIX_UNITS;
#pragma unroll
                for (int i = 0; i < FIX_UNITS; ++i) q[i] = __ldg(dp + i);   // next record (the stream is padded past the end)
                if ((r & 3) == 0 && p.pf_units > 0) prefetch_l1(dp + p.pf_units);
                if (CPLX) {
                    finish_row(S::make(u4lo(c[0]), __dmul_rn(w, u4lo(c[2]))), S::make(u4hi(c[0]), __dmul_rn(w, u4hi(c[2]))),
                               S::make(u4lo(c[1]), __dmul_rn(w, u4lo(c[FIX_UNITS - 1]))), S::zero(), S::make(u4hi(c[1]), 0.0), false);
                } else {
                    finish_row(S::make(u4lo(c[0]), 0.0), S::make(u4hi(c[0]), 0.0), S::make(u4lo(c[1]), 0.0), S::zero(), S::make(u4hi(c[1]), 0.0), false);
                }
            }
            k += cnt;
        } else {
            for (int r = 0; r < cnt; ++r, ++k) {
                const uint32_t ci = __ldg(p.rowinfo + k);
                const int nt = (int)(ci >> RI_NTERMS_SHIFT);
                T e0 = S::zero(), e1 = S::zero(), e2 = S::zero(), e3 = S::zero(), e4 = S::zero();
                for (int t = 0; t < nt; ++t) {
                    const uint4 cur = __ldg(dp + t);
                    const double K = u4hi(cur);
                    const uint32_t kind = (cur.x >> 4) & 0xfu;
                    double re = 0.0, im = 0.0;
                    if (kind == CK_CONST) re = K;
                    else if (kind == CK_JW_K) im = __dmul_rn(w, K);
                    else {
                        const double v = IN[(size_t)cur.y * L];
                        if (kind == CK_IN) re = v;
                        else if (kind == CK_INV_IN) re = __ddiv_rn(1.0, v);
                        else if (kind == CK_IN_OVER_K) re = __ddiv_rn(v, K);
                        else im = __dmul_rn(w, v);
                    }
                    if (cur.x & CT_NEG) { re = -re; im = -im; }
                    const uint32_t tgt = cur.x & 7u;   // warp-uniform: no divergence, accumulators stay in registers
                    if (tgt == CT_SUB) S::add(e0, re, im);
                    else if (tgt == CT_DIAG) S::add(e1, re, im);
                    else if (tgt == CT_SUP1) S::add(e2, re, im);
                    else if (tgt == CT_SUP2) S::add(e3, re, im);
                    else S::add(e4, re, im);
                }
                dp += nt;
                finish_row(e0, e1, e2, e3, e4, (ci & RI_HAS_U2) != 0);
            }
        }
    }

    // ---- backward: x_k = y^_k - u^_{k,k+1} x_{k+1} - u^_{k,k+2} x_{k+2}; reverse stream, double-buffered prefetch ----
    constexpr int R = CPLX ? 4 : 8;
    // DIRECT delivery: a warp collects FW consecutive unknowns of its 32 lanes in shared memory and then writes, per
    // instruction, FW*sizeof(T)-byte runs of the caller's point-major rows (longer runs = better DRAM page locality)
    constexpr int FW = CPLX ? 16 : 32;
    __shared__ T tile[DIRECT ? TPB / 32 : 1][DIRECT ? 32 : 1][FW + 1];
    int fill = 0;                                     // unknowns currently staged in the tile (warp-uniform)
    int tile_top = 0;                                 // unknown index of tile column 0 (columns go downwards)
    const int warp = threadIdx.x >> 5, lid = threadIdx.x & 31;
    const long long lane_w0 = lane - lid;             // first lane of this warp
    T x1 = S::zero(), x2 = S::zero();                 // x_{k+1}, x_{k+2}
    const size_t rs = (size_t)2 * (size_t)L;          // plane stride between consecutive rows, in T
    int nu = p.n_u2 - 1;                              // next (largest remaining) row that owns a u2 plane
    int next_u2 = nu >= 0 ? __ldg(p.u2_rows + nu) : -1;
    T ua[R], ya[R], ub[R], yb[R];
    auto load_group = [&](int k0, T *u, T *y) {       // rows k0 .. k0-R+1, all valid
        const T *base = UY + (size_t)k0 * rs;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            u[j] = __ldcs(base - (size_t)j * rs);
            y[j] = __ldcs(base - (size_t)j * rs + L);
        }
    };
    auto solve_row = [&](int k, T u, T y) -> T {
        T x = S::fnma(y, u, x1);
        if (k == next_u2) {                           // rare (e.g. the source's KCL row): not prefetched
            x = S::fnma(x, __ldcs(&UY[(size_t)(2 * n + nu) * L]), x2);
            --nu;
            next_u2 = nu >= 0 ? __ldg(p.u2_rows + nu) : -1;
        }
        x2 = x1;
        x1 = x;
        return x;
    };
    auto flush_tile = [&]() {
        // column c of the tile holds unknown tile_top - c; each pass writes one lane-group's run in ascending order
        __syncwarp();
        T *out = reinterpret_cast<T *>(p.OUT);
        constexpr int LANES_PER_PASS = 32 / FW > 0 ? 32 / FW : 1;
        constexpr int PASSES = 32 / LANES_PER_PASS;
#pragma unroll 4
        for (int pass = 0; pass < PASSES; ++pass) {
            const int l = pass * LANES_PER_PASS + lid / FW, c = FW - 1 - (lid % FW);
            const long long gl = lane_w0 + l;
            if (c < fill && gl < p.out_lanes) out[(size_t)gl * (size_t)n + (size_t)(tile_top - c)] = tile[warp][l][c];
        }
        __syncwarp();
        fill = 0;
    };
    auto deliver_group = [&](int k0, const T *u, const T *y) {
        T xs[R];
#pragma unroll
        for (int j = 0; j < R; ++j) xs[j] = solve_row(k0 - j, u[j], y[j]);
        if (DIRECT) {
            if (fill == 0) tile_top = k0;
#pragma unroll
            for (int j = 0; j < R; ++j) tile[warp][lid][fill + j] = xs[j];
            fill += R;
            if (fill == FW) flush_tile();
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) X[(size_t)__ldg(p.unknown_of_row + (k0 - j)) * L] = xs[j];
        }
    };
    int k0 = (p.debug & 2) ? -1 : n - 1;
    if (k0 >= R - 1) {
        load_group(k0, ua, ya);
        while (k0 >= 2 * R - 1 + R) {                 // two more full groups follow the current one
            load_group(k0 - R, ub, yb);
            deliver_group(k0, ua, ya);
            load_group(k0 - 2 * R, ua, ya);
            deliver_group(k0 - R, ub, yb);
            k0 -= 2 * R;
        }
        deliver_group(k0, ua, ya);
        k0 -= R;
        while (k0 >= R - 1) { load_group(k0, ua, ya); deliver_group(k0, ua, ya); k0 -= R; }
    }
    for (; k0 >= 0; --k0) {                           // tail: fewer than R rows
        const T x = solve_row(k0, __ldcs(&UY[(size_t)(2 * k0) * L]), __ldcs(&UY[(size_t)(2 * k0 + 1) * L]));
        if (DIRECT) {
            if (fill == 0) tile_top = k0;
            tile[warp][lid][fill++] = x;
            if (fill == FW) flush_tile();
        } else X[(size_t)__ldg(p.unknown_of_row + k0) * L] = x;
    }
    if (DIRECT && fill > 0) flush_tile();
    if (!S::finite(x1)) status |= KRB_LANE_NONFINITE;   // non-finite values propagate down the recurrence to x_0
    if (grow >= KRB_GROWTH_KEY && grow < KRB_NONFINITE_KEY) status |= KRB_LANE_INACCURATE;
    if (p.status) p.status[lane] |= status;
}

}  // namespace krb
