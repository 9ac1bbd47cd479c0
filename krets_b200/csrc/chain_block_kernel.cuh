// Block-partitioned chain executor: the same plans as chain_kernel.cuh (permuted matrix with lower bandwidth <= 1 and
// upper bandwidth <= 2), but one CTA of T threads solves one lane and NOTHING except the per-lane inputs and the n results
// touches HBM (SURVEY §8(d)'s fused floor  B_floor = P + s*n  instead of the sequential kernel's  s*5n).
//
//   thread t owns pivot rows [t*M, (t+1)*M): M-1 interior rows and one separator (the block's last row)
//   1. stamp + eliminate the interior rows top-down inside the thread; the coupling to the previous block's separator is
//      carried as a fill column l^ (Wang's partition method):   x_j + u^_j x_{j+1} (+ u2^_j x_{j+2}) + l^_j s_{t-1} = y^_j
//   2. eliminate bottom-up:                                     x_j = p_j - r_j s_t - q_j s_{t-1}
//   3. the separator rows, with x_{first of block t+1} substituted, form a T x T tridiagonal system in s_0..s_{T-1}:
//      parallel cyclic reduction in shared memory (log2 T steps)
//   4. interiors by substitution, rows staged in shared memory and written point-major with full-width coalesced stores.
//
// u^ / y^ (later r / p) live in registers, l^ (later q, then x) in a per-thread shared-memory column with an odd stride.
// Stamped values are bit-identical to chain_kernel.cuh (same records, same summation order); the elimination order differs,
// so unknowns agree with the sequential kernel to ~M ulp-amplifications (host check in compiler.cpp::build_chain_block
// admits a plan only if both orders agree to 5e-10 at the nominal inputs; SURVEY §8(c): values are held to 1e-9).
#pragma once
#include "chain_kernel.cuh"

namespace krb {

struct ChainBlockParams {
    const uint2 *rows;           // [M][T]: .x = FIXED ? (dict index << 1) | 1 : (stream unit offset << 1); .y = (#terms << 8) | HAS_U2
    const uint4 *dict;           // unique FIXED records (2 units real / 4 units complex)
    const uint4 *stream;         // GENERIC term runs (the sequential kernel's record stream)
    const int *unknown_of_row;   // nullptr when pivot row k delivers unknown k
    const double *IN;            // [n_inputs][L]
    void *OUT;                   // point-major x_out[lane][unknown]
    long long lanes;
    long long L;                 // lane stride of the input planes
    int n;
    int input_freq;
    int *status;
};

// exponent field all-zero (0 / denormal) or all-one (inf / nan): integer pipe only
__device__ __forceinline__ bool bad_pivot(double w) {
    const uint32_t hi = (uint32_t)__double2hiint(w);
    return ((hi & 0x7ff00000u) - 0x00100000u) >= 0x7fe00000u;
}

template <bool CPLX> struct BlockSc;
template <> struct BlockSc<false> {
    typedef double T;
    // reciprocal of the pivot to ~1 ulp, returned as (r0, t):  q / w = fma(q*r0, t, q*r0)
    struct Inv { double r0, t; };
    static __device__ __forceinline__ Inv inv(double w, bool &bad) {
        bad = bad_pivot(w);
        const double r0 = rcp_seed(w);
        const double e = fma(-w, r0, 1.0);
        return Inv{r0, fma(e, e, e)};
    }
    static __device__ __forceinline__ double div(double a, const Inv &i) { const double a0 = a * i.r0; return fma(a0, i.t, a0); }
    static __device__ __forceinline__ double neg(double a) { return -a; }
    static __device__ __forceinline__ double one() { return 1.0; }
};
template <> struct BlockSc<true> {
    typedef double2 T;
    struct Inv { double2 c; };    // conj(w) / |w|^2
    static __device__ __forceinline__ Inv inv(double2 w, bool &bad) {
        const double ns = fma(w.x, w.x, w.y * w.y);
        bad = bad_pivot(ns);
        const double r = fast_rcp(ns);
        return Inv{make_double2(w.x * r, -w.y * r)};
    }
    static __device__ __forceinline__ double2 div(double2 a, const Inv &i) { return ChainSc<true>::mul(a, i.c); }
    static __device__ __forceinline__ double2 neg(double2 a) { return make_double2(-a.x, -a.y); }
    static __device__ __forceinline__ double2 one() { return make_double2(1.0, 0.0); }
};

template <bool CPLX, int T, int M>
__global__ void __launch_bounds__(T, CPLX ? 1 : 2) krb_chain_block_kernel(const ChainBlockParams p) {
    typedef ChainSc<CPLX> S;
    typedef BlockSc<CPLX> B;
    typedef typename S::T V;
    constexpr int MI = M - 1;               // interior rows per thread
    constexpr int XS = M + 1;               // stride of a thread's shared-memory column (odd: conflict-free)
    constexpr int FIX_UNITS = CPLX ? 4 : 2;
    extern __shared__ __align__(16) unsigned char krb_block_smem[];
    V *xq = reinterpret_cast<V *>(krb_block_smem);   // [T][XS]: l^_j -> q_j -> x_j
    V *red = xq + T * XS;                              // [2][3][T]: cyclic-reduction rows (a, c, d), double-buffered
    V *fst = red + 3 * T;                              // [3][T]: (p, r, q) of each block's first row; aliases buffer 1
    const int t = threadIdx.x;
    V *mycol = xq + t * XS;
    const long long L = p.L;
    const int n = p.n;

    for (long long lane = blockIdx.x; lane < p.lanes; lane += gridDim.x) {
        const double *__restrict__ IN = p.IN + lane;
        double w = 0.0;                              // (2.0 * PI) * f, rounded once as the reference does
        if (CPLX && p.input_freq >= 0) w = __dmul_rn(2.0 * 3.14159265358979323846264338327950288, __ldg(IN + (size_t)p.input_freq * L));
        bool bad = false;
        V u[MI], y[MI];
        V lprev = S::zero(), pu2 = S::zero(), u2v = S::zero();
        bool pu2_live = false;
        int ju = -1;                                 // local index of this block's row with a second super-diagonal
        V sepA = S::zero(), sepB = S::zero(), sepD = S::zero(), sepS = S::zero();

        // ---- 1. stamp + top-down elimination of the interior rows; separator row partially reduced ----
#pragma unroll
        for (int j = 0; j < M; ++j) {
            const uint2 rw = __ldg(p.rows + j * T + t);
            V e0, e1, e2, e3 = S::zero(), e4;
            bool has_u2 = false;
            if (rw.x & 1u) {
                const uint4 *d = p.dict + (size_t)(rw.x >> 1) * FIX_UNITS;
                const uint4 c0 = __ldg(d), c1 = __ldg(d + 1);
                if (CPLX) {
                    const uint4 c2 = __ldg(d + 2), c3 = __ldg(d + 3);
                    e0 = S::make(u4lo(c0), __dmul_rn(w, u4lo(c2)));
                    e1 = S::make(u4hi(c0), __dmul_rn(w, u4hi(c2)));
                    e2 = S::make(u4lo(c1), __dmul_rn(w, u4lo(c3)));
                    e4 = S::make(u4hi(c1), 0.0);
                } else {
                    e0 = S::make(u4lo(c0), 0.0); e1 = S::make(u4hi(c0), 0.0); e2 = S::make(u4lo(c1), 0.0); e4 = S::make(u4hi(c1), 0.0);
                }
            } else {
                // GENERIC row: ordered term run (same decoding and summation order as chain_kernel.cuh)
                const uint4 *dp = p.stream + (rw.x >> 1);
                const int nt = (int)(rw.y >> RI_NTERMS_SHIFT);
                has_u2 = (rw.y & RI_HAS_U2) != 0;
                e0 = S::zero(); e1 = S::zero(); e2 = S::zero(); e4 = S::zero();
                for (int q = 0; q < nt; ++q) {
                    const uint4 cur = __ldg(dp + q);
                    const double K = u4hi(cur);
                    const uint32_t kind = (cur.x >> 4) & 0xfu;
                    double re = 0.0, im = 0.0;
                    if (kind == CK_CONST) re = K;
                    else if (kind == CK_JW_K) im = __dmul_rn(w, K);
                    else {
                        const double v = __ldg(IN + (size_t)cur.y * L);
                        if (kind == CK_IN) re = v;
                        else if (kind == CK_INV_IN) re = __ddiv_rn(1.0, v);
                        else if (kind == CK_IN_OVER_K) re = __ddiv_rn(v, K);
                        else im = __dmul_rn(w, v);
                    }
                    if (cur.x & CT_NEG) { re = -re; im = -im; }
                    const uint32_t tgt = cur.x & 7u;
                    // the target differs from thread to thread: select the accumulator, add once, select back
                    V acc = tgt == CT_SUB ? e0 : tgt == CT_DIAG ? e1 : tgt == CT_SUP1 ? e2 : tgt == CT_SUP2 ? e3 : e4;
                    S::add(acc, re, im);
                    e0 = tgt == CT_SUB ? acc : e0;
                    e1 = tgt == CT_DIAG ? acc : e1;
                    e2 = tgt == CT_SUP1 ? acc : e2;
                    e3 = tgt == CT_SUP2 ? acc : e3;
                    e4 = tgt >= CT_RHS ? acc : e4;
                }
            }
            if (j < MI) {
                V wk, s1, yk, lk;
                if (j == 0) { wk = e1; s1 = e2; yk = e4; lk = e0; }
                else {
                    wk = S::fnma(e1, e0, u[j - 1]);
                    s1 = pu2_live ? S::fnma(e2, e0, pu2) : e2;
                    yk = S::fnma(e4, e0, y[j - 1]);
                    lk = B::neg(S::mul(e0, lprev));
                }
                bool b1;
                const typename B::Inv iv = B::inv(wk, b1);
                bad = bad || b1;
                u[j] = B::div(s1, iv);
                y[j] = B::div(yk, iv);
                lprev = B::div(lk, iv);
                mycol[j] = lprev;
                pu2_live = has_u2;
                if (has_u2) { pu2 = B::div(e3, iv); u2v = pu2; ju = j; }
            } else {
                // separator row: sub * x_{last interior} substituted from the forward-eliminated last interior row
                sepA = B::neg(S::mul(e0, lprev));
                sepB = S::fnma(e1, e0, u[MI - 1]);
                sepD = S::fnma(e4, e0, y[MI - 1]);
                sepS = e2;
            }
        }

        // ---- 2. bottom-up: x_j = p_j - r_j s_t - q_j s_{t-1}  (p in y[], r in u[], q in the shared column) ----
        {
            // last interior row: the virtual separator row below it is (p, r, q) = (0, -1, 0)
            V qn = mycol[MI - 1];                     // q_{M-2} = l^_{M-2}
            // p_{M-2} = y^, r_{M-2} = u^: already in place
            if (__any_sync(0xffffffffu, ju >= 0)) {
                V p2 = S::zero(), r2 = B::neg(B::one()), q2 = S::zero();   // row j+2 (starts at the virtual separator)
#pragma unroll
                for (int j = MI - 2; j >= 0; --j) {
                    const V uj = u[j];
                    V pj = S::fnma(y[j], uj, y[j + 1]);
                    V rj = B::neg(S::mul(uj, u[j + 1]));
                    V qj = S::fnma(mycol[j], uj, qn);
                    if (j == ju) { pj = S::fnma(pj, u2v, p2); rj = S::fnma(rj, u2v, r2); qj = S::fnma(qj, u2v, q2); }
                    p2 = y[j + 1]; r2 = u[j + 1]; q2 = qn;
                    y[j] = pj; u[j] = rj; qn = qj;
                    mycol[j] = qj;
                }
            } else {
#pragma unroll
                for (int j = MI - 2; j >= 0; --j) {
                    const V uj = u[j];
                    y[j] = S::fnma(y[j], uj, y[j + 1]);
                    u[j] = B::neg(S::mul(uj, u[j + 1]));
                    qn = S::fnma(mycol[j], uj, qn);
                    mycol[j] = qn;
                }
            }
            fst[t] = y[0]; fst[T + t] = u[0]; fst[2 * T + t] = qn;
        }
        __syncthreads();

        // ---- 3. separator system: A s_{t-1} + B s_t + C s_{t+1} = D, parallel cyclic reduction ----
        V a, c, d;
        {
            V pn = S::zero(), rn = S::zero(), qn = S::zero();
            if (t + 1 < T) { pn = fst[t + 1]; rn = fst[T + t + 1]; qn = fst[2 * T + t + 1]; }
            const V bb = S::fnma(sepB, sepS, qn);
            const V cc = B::neg(S::mul(sepS, rn));
            const V dd = S::fnma(sepD, sepS, pn);
            bool b1;
            const typename B::Inv iv = B::inv(bb, b1);
            bad = bad || b1;
            a = B::div(sepA, iv); c = B::div(cc, iv); d = B::div(dd, iv);
        }
        red[t] = a; red[T + t] = c; red[2 * T + t] = d;
        __syncthreads();
        int cur = 0;
#pragma unroll
        for (int s = 1; s < T; s <<= 1) {
            const V *src = red + cur * 3 * T;
            V *dst = red + (cur ^ 1) * 3 * T;
            V am = S::zero(), cm = S::zero(), dm = S::zero(), ap = S::zero(), cp = S::zero(), dq = S::zero();
            if (t >= s) { am = src[t - s]; cm = src[T + t - s]; dm = src[2 * T + t - s]; }
            if (t + s < T) { ap = src[t + s]; cp = src[T + t + s]; dq = src[2 * T + t + s]; }
            const V b2 = S::fnma(S::fnma(B::one(), a, cm), c, ap);
            const V a2 = B::neg(S::mul(a, am));
            const V c2 = B::neg(S::mul(c, cp));
            const V d2 = S::fnma(S::fnma(d, a, dm), c, dq);
            bool b1;
            const typename B::Inv iv = B::inv(b2, b1);
            bad = bad || b1;
            a = B::div(a2, iv); c = B::div(c2, iv); d = B::div(d2, iv);
            dst[t] = a; dst[T + t] = c; dst[2 * T + t] = d;
            __syncthreads();
            cur ^= 1;
        }
        const V st = d;
        const V sm = t > 0 ? (red + cur * 3 * T)[2 * T + t - 1] : S::zero();

        // ---- 4. interiors, staged in the shared column, then coalesced point-major delivery ----
#pragma unroll
        for (int j = 0; j < MI; ++j) mycol[j] = S::fnma(S::fnma(y[j], u[j], st), mycol[j], sm);
        mycol[MI] = st;
        int status = bad ? KRB_LANE_ZERO_PIVOT : 0;
        if (!S::finite(st)) status |= KRB_LANE_NONFINITE;
        if (status && p.status) atomicOr(p.status + lane, status);
        __syncthreads();
        V *__restrict__ out = reinterpret_cast<V *>(p.OUT) + (size_t)lane * (size_t)n;
        if (p.unknown_of_row == nullptr) {
            for (int i = t; i < n; i += T) __stcs(out + i, xq[(i / M) * XS + (i % M)]);
        } else {
            for (int i = t; i < n; i += T) __stcs(out + __ldg(p.unknown_of_row + i), xq[(i / M) * XS + (i % M)]);
        }
        __syncthreads();                              // the columns and the reduction buffers are reused by the next lane
    }
}

}  // namespace krb
