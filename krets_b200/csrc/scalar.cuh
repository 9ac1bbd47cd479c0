// Scalar arithmetic shared by the kernels: real f64 and complex f64 (num_complex-style products, as the reference's c64).
#pragma once
#include <cuda_runtime.h>

namespace krb {

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

}  // namespace krb
