// TEST INFRASTRUCTURE ONLY: lets g++ compile the text of a circuit-specialised lane kernel (krb_plan_jit_source with
// host_mode = 1, krets_b200/csrc/vm_jit.cpp) as ordinary host C++, so that the CPU test tier can run the generated statements
// lane by lane against the oracle (tests/jit_host.py).  Device builds never see this file.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
// round-to-nearest single operations: host doubles are IEEE already; volatile keeps g++ from contracting a*b+c into an fma
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline int __double2hiint(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (int)(b >> 32); }
static inline double __longlong_as_double(long long b) { double v; std::memcpy(&v, &b, 8); return v; }
// the device seeds its pivot reciprocal with MUFU.RCP64H (2^-20) and corrects twice; an exact seed makes the corrections no-ops
static inline double krb_rcp_seed(double x) { return 1.0 / x; }
using std::exp;
using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
using std::isfinite;
