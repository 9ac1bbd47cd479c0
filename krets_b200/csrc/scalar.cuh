// Scalar arithmetic shared by the kernels: real f64 and complex f64 (num_complex-style products, as the reference's c64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/krets_b200.h"

namespace krb {

struct c64 { double re, im; };

// Pivot growth on the integer pipe: high word of |v| (complex: of the larger component).  A lane whose largest scaled factor entry
// |u^| = |u / pivot| reaches 2^KRB_GROWTH_LIMIT_LOG2 gets KRB_LANE_INACCURATE (include/krets_b200.h): its static pivot cancelled that
// many bits against its row, where faer's per-factorisation partial pivoting (op.rs:50) would have exchanged rows.
__device__ __forceinline__ uint32_t growth_key(double v) { return (uint32_t)__double2hiint(v) & 0x7fffffffu; }
__device__ __forceinline__ uint32_t growth_key(double2 v) { return max(growth_key(v.x), growth_key(v.y)); }
constexpr uint32_t KRB_GROWTH_KEY = (uint32_t)(1023 + KRB_GROWTH_LIMIT_LOG2) << 20;
constexpr uint32_t KRB_NONFINITE_KEY = 0x7ff00000u;


// 1/x to ~1 ulp without the IEEE division's long instruction sequence (and its slow-path call): MUFU.RCP64H seed (2^-20) and two fused
// correction steps, r0 (1 + e + e^2) with e = 1 - x r0.  Exponents near the ends of the range (the seed flushes / overflows) take the
// exact division.  Used for pivot reciprocals on dependent chains; the LU is not bit-pinned (values are held to 1e-9).
__device__ __forceinline__ double krb_fast_rcp(double x) {
    const uint32_t e = ((uint32_t)__double2hiint(x) >> 20) & 0x7ffu;
    if (e - 24u > 1999u) return 1.0 / x;
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double er = fma(-x, r0, 1.0);
    return fma(r0, fma(er, er, er), r0);
}

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
    static __device__ __forceinline__ T rcpf(T a, bool &bad) { bad = !(fabs(a) > 0.0) || !isfinite(a); return krb_fast_rcp(a); }
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
    static __device__ __forceinline__ T rcpf(T a, bool &bad) {
        const double ns = fma(a.x, a.x, a.y * a.y);
        bad = !(ns > 0.0) || !isfinite(ns);
        const double r = krb_fast_rcp(ns);
        return make_double2(a.x * r, -a.y * r);
    }
    static __device__ __forceinline__ T mulj(T a) { return make_double2(0.0, a.x); }
    static __device__ __forceinline__ bool finite(T a) { return isfinite(a.x) && isfinite(a.y); }
};

// result selection (extension, krb_io.select): out[point][k] = x[point][select[k]], point = (lane, row)
template <typename V>
__global__ void krb_select_columns(const V *__restrict__ x, V *__restrict__ out, const int *__restrict__ select, int n, int n_select, long long points) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < points * n_select) {
        const long long pt = i / n_select;
        out[i] = x[pt * n + __ldg(select + (int)(i - pt * n_select))];
    }
}

}  // namespace krb
