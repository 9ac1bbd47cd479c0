// "Chain" executor: plans whose permuted matrix has lower bandwidth <= 1 and upper bandwidth <= 2 (ladders, RC/RLC
// lines and other path-like circuits, including the source row that couples two neighbours).
//
// One thread = one lane.  The forward sweep stamps row k from a lane-invariant TERM STREAM (16 B per term, one broadcast
// load per warp, prefetched one term ahead), eliminates it against the previous row kept in registers (Thomas-style
// recurrence on the LDU form  u^_k = sup_k / w_k,  y^_k = y_k / w_k) and streams u^_k, y^_k to lane-interleaved HBM
// planes; the backward sweep re-reads them in reverse with an 8-row register prefetch ring and delivers x.  Per point
// the kernel moves 2n values out, 2n back and n results — a third of the 3-pass stamp/factor/solve traffic of SURVEY §8(d).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/krets_b200.h"
#include "vm_kernel.cuh"

namespace krb {

struct ChainTermDev { uint32_t meta; uint32_t plane; double K; };   // same layout as krb::ChainTerm (plan.hpp)

enum : uint32_t {
    CT_SUB = 0, CT_DIAG = 1, CT_SUP1 = 2, CT_SUP2 = 3, CT_RHS = 4,
    CK_CONST = 0,       // K                               (lane-invariant stamps, folded in stamp order on the host)
    CK_IN = 1,          // IN[plane]                       (swept / Monte-Carlo source value)
    CK_INV_IN = 2,      // 1.0 / IN[plane]                 (Monte-Carlo resistor, stampable.rs:232)
    CK_IN_OVER_K = 3,   // IN[plane] / K                   (Monte-Carlo C or L over h)
    CK_JW_K = 4,        // (0, ((2 pi) f) K)               (C / L admittance, stampable.rs:379-382, 715)
    CK_JW_IN = 5,       // (0, ((2 pi) f) IN[plane])
    CT_NEG = 1u << 8,
    CT_HAS_U2 = 1u << 30,   // (on the last term of a row) the row has a second super-diagonal: u2 lives in plane 2n + rank
    CT_LAST = 1u << 31
};

struct ChainParams {
    const ChainTermDev *terms;
    const int *unknown_of_row;   // pivot row k delivers unknown unknown_of_row[k]
    const uint32_t *u2_flags;    // bitset over rows (null when n_u2 == 0)
    const double *IN;            // [n_inputs][L]
    void *UY;                    // [2n + n_u2][L] of T: plane 2k = u^_{k,k+1}, 2k+1 = y^_k, 2n + r = u^_{k,k+2} of the r-th flagged row
    void *X;                     // [n][L] of T
    int n;
    int input_freq;              // input plane holding the AC frequency, or -1
    int n_u2;
    long long L;
    int *status;
};

__device__ __forceinline__ double fast_rcp(double x) {
    // MUFU.RCP64H seed + two Newton steps: <= 1-2 ulp, no slow path (the LU is not bit-pinned; stamps use exact division)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

template <bool CPLX> struct ChainSc;
template <> struct ChainSc<false> {
    typedef double T;
    static __device__ __forceinline__ T zero() { return 0.0; }
    static __device__ __forceinline__ T fnma(T d, T a, T b) { return fma(-a, b, d); }
    static __device__ __forceinline__ T mul(T a, T b) { return a * b; }
    static __device__ __forceinline__ T rcp(T a, bool &bad) { const double r = fast_rcp(a); bad = !isfinite(r); return r; }
    static __device__ __forceinline__ bool finite(T a) { return isfinite(a); }
};
template <> struct ChainSc<true> {
    typedef double2 T;
    static __device__ __forceinline__ T zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ T fnma(T d, T a, T b) {
        return make_double2(fma(a.y, b.y, fma(-a.x, b.x, d.x)), fma(-a.y, b.x, fma(-a.x, b.y, d.y)));
    }
    static __device__ __forceinline__ T mul(T a, T b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
    static __device__ __forceinline__ T rcp(T a, bool &bad) {
        const double ns = fma(a.x, a.x, a.y * a.y);
        const double r = fast_rcp(ns);
        bad = !isfinite(r);
        return make_double2(a.x * r, -a.y * r);
    }
    static __device__ __forceinline__ bool finite(T a) { return isfinite(a.x) && isfinite(a.y); }
};

template <bool CPLX>
__device__ __forceinline__ void chain_add(typename ChainSc<CPLX>::T &acc, double re, double im);
template <> __device__ __forceinline__ void chain_add<false>(double &acc, double re, double) { acc = __dadd_rn(acc, re); }
template <> __device__ __forceinline__ void chain_add<true>(double2 &acc, double re, double im) {
    acc.x = __dadd_rn(acc.x, re);
    acc.y = __dadd_rn(acc.y, im);
}

template <bool CPLX, int TPB>
__global__ void __launch_bounds__(TPB) krb_chain_kernel(const ChainParams p) {
    typedef ChainSc<CPLX> S;
    typedef typename S::T T;
    const long long lane = (long long)blockIdx.x * TPB + threadIdx.x;
    const long long L = p.L;
    const double *__restrict__ IN = p.IN + lane;
    T *__restrict__ UY = reinterpret_cast<T *>(p.UY) + lane;
    T *__restrict__ X = reinterpret_cast<T *>(p.X) + lane;
    const int n = p.n;
    int status = 0;

    double w = 0.0;                                   // (2.0 * PI) * f, rounded once as the reference does
    if (CPLX && p.input_freq >= 0) w = __dmul_rn(2.0 * 3.14159265358979323846264338327950288, IN[(size_t)p.input_freq * L]);

    // ---- forward: stamp + eliminate + forward-substitute, row by row ----
    const uint4 *__restrict__ tp = reinterpret_cast<const uint4 *>(p.terms);
    T pu1 = S::zero(), pu2 = S::zero(), py = S::zero();   // u^_{k-1,k}, u^_{k-1,k+1}, y^_{k-1}
    int r2 = 0;                                         // rank of the next row that owns a u2 plane (warp-uniform)
    uint4 t = __ldg(tp);
    for (int k = 0; k < n; ++k) {
        T e0 = S::zero(), e1 = S::zero(), e2 = S::zero(), e3 = S::zero(), e4 = S::zero();
        uint32_t meta;
        for (;;) {
            const uint4 cur = t;
            t = __ldg(++tp);                          // next term (the stream ends with one sentinel)
            const double K = __hiloint2double((int)cur.w, (int)cur.z);
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
            const uint32_t tgt = cur.x & 7u;            // warp-uniform: no divergence, accumulators stay in registers
            if (tgt == CT_SUB) chain_add<CPLX>(e0, re, im);
            else if (tgt == CT_DIAG) chain_add<CPLX>(e1, re, im);
            else if (tgt == CT_SUP1) chain_add<CPLX>(e2, re, im);
            else if (tgt == CT_SUP2) chain_add<CPLX>(e3, re, im);
            else chain_add<CPLX>(e4, re, im);
            meta = cur.x;
            if (meta & CT_LAST) break;
        }
        const T wk = S::fnma(e1, e0, pu1);            // w_k     = diag - sub * u^_{k-1,k}
        const T s1 = S::fnma(e2, e0, pu2);            // sup1'   = sup1 - sub * u^_{k-1,k+1}
        const T yk = S::fnma(e4, e0, py);             // y_k     = rhs  - sub * y^_{k-1}
        bool bad;
        const T inv = S::rcp(wk, bad);
        if (bad) status |= KRB_LANE_ZERO_PIVOT;
        pu1 = S::mul(s1, inv);
        pu2 = S::mul(e3, inv);
        py = S::mul(yk, inv);
        __stcs(&UY[(size_t)(2 * k) * L], pu1);
        __stcs(&UY[(size_t)(2 * k + 1) * L], py);
        if (meta & CT_HAS_U2) { __stcs(&UY[(size_t)(2 * n + r2) * L], pu2); ++r2; }
    }

    // ---- backward: x_k = y^_k - u^_{k,k+1} x_{k+1} - u^_{k,k+2} x_{k+2}, reverse stream with an 8-row prefetch ring ----
    constexpr int R = 8;
    T x1 = S::zero(), x2 = S::zero();                 // x_{k+1}, x_{k+2}
    const uint32_t *__restrict__ flags = p.u2_flags;  // bit per row: owns a u2 plane
    for (int k0 = n - 1; k0 >= 0; k0 -= R) {
        T u[R], y[R];
        int un[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int k = k0 - j;
            if (k >= 0) {
                u[j] = __ldcs(&UY[(size_t)(2 * k) * L]);
                y[j] = __ldcs(&UY[(size_t)(2 * k + 1) * L]);
                un[j] = __ldg(p.unknown_of_row + k);
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int k = k0 - j;
            if (k >= 0) {
                T x = S::fnma(y[j], u[j], x1);
                if (flags && ((__ldg(flags + (k >> 5)) >> (k & 31)) & 1u)) {   // rare (e.g. the source row): not prefetched
                    --r2;
                    x = S::fnma(x, __ldcs(&UY[(size_t)(2 * n + r2) * L]), x2);
                }
                if (!S::finite(x)) status |= KRB_LANE_NONFINITE;
                X[(size_t)un[j] * L] = x;
                x2 = x1;
                x1 = x;
            }
        }
    }
    if (p.status) p.status[lane] |= status;
}

}  // namespace krb
