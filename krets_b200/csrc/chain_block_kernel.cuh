// Block-partitioned chain executor: the same plans as chain_kernel.cuh (permuted matrix with lower bandwidth <= 1 and
// upper bandwidth <= 2), but one CTA of T threads solves one lane and NOTHING except the per-lane inputs and the n results
// touches HBM (SURVEY §8(d)'s fused floor  B_floor = P + s*n  instead of the sequential kernel's  s*5n).
//
//   thread t owns pivot rows [t*M, (t+1)*M): M-1 interior rows and one separator (the block's last row)
//   1. stamp + eliminate the interior rows top-down inside the thread; the coupling to the previous block's separator is
//      carried as a fill column l^ (Wang's partition method):   x_j + u^_j x_{j+1} (+ u2^_j x_{j+2}) + l^_j s_{t-1} = y^_j.
//      The relation of the block's FIRST row to its two separators,  x_0 = P - R s_t - Q s_{t-1},  is accumulated in the same
//      sweep (x_0 = accP - accQ s_{t-1} + fA x_j + fB x_{j+1}, row j substituted as soon as it is eliminated): no bottom-up pass
//   2. the separator rows, with x_{first of block t+1} substituted, form a T x T tridiagonal system in s_0..s_{T-1}:
//      two levels of cyclic reduction (odd rows absorb the even ones, rows kept de-interleaved by parity so that every
//      access is unit-stride), parallel cyclic reduction on the remaining T/4 rows by a quarter of the warps (log2(T/4)
//      steps in shared memory), eliminated rows by substitution
//   3. interiors by back substitution  x_j = (y^_j - l^_j s_{t-1}) - u^_j x_{j+1} (- u2^_j x_{j+2}),  written to a second
//      shared-memory buffer in pivot-row order — the lane's result vector, contiguous — and delivered by ONE bulk asynchronous
//      copy (cp.async.bulk.global.shared::cta, the TMA engine) issued by thread 0: the CTA starts the next lane at once and the
//      delivery costs no LDS / STG issue slots (ncu r2a: the per-thread delivery loop was 19 % of all warp samples and 23 % of
//      the shared-memory wavefronts, 38 % of them bank conflicts of the padded column layout).
//
// u^ / y^ live in registers, l^ in a per-thread shared-memory column with an odd stride (conflict-free), x in the result buffer.  Register budget: 2 (M-1)
// values of u^ / y^ leave ~50 of the 128 registers (2 CTAs / SM) for everything else, so lane-invariant address arithmetic is
// deliberately kept INSIDE the lane loop (opaque copies of t / the packed entry indices): hoisted, it was spilled and every
// lane paid ~90 local-memory accesses per thread (ncu: 33.6 M -> 1.8 M local loads per 65 536-lane launch, -18 % run time).
// Stamped values are bit-identical to chain_kernel.cuh (same records, same summation order); the elimination order differs,
// so unknowns agree with the sequential kernel to ~M ulp-amplifications (host check in compiler.cpp::build_chain_block
// admits a plan only if both orders agree to 5e-10 at the nominal inputs; SURVEY §8(c): values are held to 1e-9).
#pragma once
#include "chain_kernel.cuh"

namespace krb {

struct ChainBlockParams {
    const uint2 *rows;           // [M][T]: .x = FIXED ? (dict index << 1) | 1 : (stream unit offset << 1); .y = BR_* fields
    const uint4 *dict;           // unique FIXED records (2 units real / 4 units complex)
    const uint4 *stream;         // GENERIC term runs (the sequential kernel's record stream)
    const int *unknown_of_row;   // nullptr when pivot row k delivers unknown k
    const int *gen_rows;         // shared-dictionary mode: the GENERIC pivot rows in slot order (n_gen entries)
    const double *IN;            // [n_inputs][L]
    void *OUT;                   // point-major x_out[lane][unknown]
    long long lanes;
    long long L;                 // lane stride of the input planes
    int n;
    int input_freq;
    int n_dict;                  // unique FIXED records
    int n_gen;                   // GENERIC rows (shared-dictionary mode: n_dict + n_gen <= the kernel's dictionary capacity)
    int *status;
    int bulk_store;              // deliver with cp.async.bulk (needs unknown_of_row == nullptr; complex: OUT 16-byte aligned)
    const void *E;               // !SD: pre-stamped rows E[lane][M][T][4] = (sub, diag, sup, rhs) written by krb_block_stamp_rows
    const void *U2;              // !SD: U2[lane][T] second super-diagonal of the (at most one) u2 row of each block
    const int *select;           // extension (krb_io.select): deliver only these unknowns, OUT = x_out[lane][n_select]; needs pivot row k == unknown k
    int n_select;
};

// row word .y: bit 1 HAS_U2, bits 2..15 #terms, bits 16..31 slot of a GENERIC row (shared-dictionary mode)
enum : uint32_t { BR_HAS_U2 = 2u, BR_NT_SHIFT = 2, BR_NT_MASK = 0x3fffu, BR_SLOT_SHIFT = 16 };
constexpr int KRB_BLOCK_DMAX = 256;   // entries of the shared-memory dictionary (unique FIXED records + GENERIC rows), real plans
constexpr int KRB_BLOCK_DMAX_CPLX = 64;   // complex plans: two 80 KB lane buffers leave room for 64 entries (227 KB per CTA)
template <bool CPLX> struct BlockDmax { static constexpr int value = CPLX ? KRB_BLOCK_DMAX_CPLX : KRB_BLOCK_DMAX; };
// shared-memory footprint in elements of V (runtime.cu sizes the launch with the same expression)
template <bool CPLX> struct BlockShfl;
template <> struct BlockShfl<false> {
    static __device__ __forceinline__ double up(double v, int s) { return __shfl_up_sync(0xffffffffu, v, s); }
    static __device__ __forceinline__ double down(double v, int s) { return __shfl_down_sync(0xffffffffu, v, s); }
};
template <> struct BlockShfl<true> {
    static __device__ __forceinline__ double2 up(double2 v, int s) { return make_double2(__shfl_up_sync(0xffffffffu, v.x, s), __shfl_up_sync(0xffffffffu, v.y, s)); }
    static __device__ __forceinline__ double2 down(double2 v, int s) { return make_double2(__shfl_down_sync(0xffffffffu, v.x, s), __shfl_down_sync(0xffffffffu, v.y, s)); }
};
template <bool CPLX, int T, int M, bool SD> constexpr size_t block_smem_elems() {
    return (size_t)T * ((M & 1) ? M : M + 1) + ((size_t)T * M + 2) + (size_t)6 * T + 16 + (SD ? (size_t)5 * BlockDmax<CPLX>::value : 0);
}

// Pivot health on the integer pipe.  The lane keeps ONE running maximum, of growth_key(u^) = the high word of |u^| = |sup / pivot|
// (complex: of the larger component):
//   * a pivot that is zero, denormal, inf or nan makes u^ non-finite (rcp.approx.ftz -> inf/0 -> nan in the correction step), i.e.
//     key >= 0x7ff00000                                               -> KRB_LANE_ZERO_PIVOT (the reference: DecompositionFailed)
//   * |u^| >= 2^KRB_GROWTH_LIMIT_LOG2: the static pivot cancelled that many bits against its row, where faer's per-factorisation
//     partial pivoting (op.rs:50) would have exchanged rows           -> KRB_LANE_INACCURATE
// Pivots whose quotients are not all kept (the separator system's last PCR step) are tested directly (pivot_key).
__device__ __forceinline__ uint32_t pivot_key(double w) { return ((uint32_t)__double2hiint(w) & 0x7ff00000u) - 0x00100000u; }
constexpr uint32_t KRB_BAD_PIVOT_KEY = 0x7fe00000u;

#ifndef KRB_OLD_FULL_BARRIER
#define KRB_BLOCK_ISSUER(NW) (((NW) / 2 - 1) * 32)
#else
#define KRB_BLOCK_ISSUER(NW) 0
#endif

template <bool CPLX> struct BlockSc;
template <> struct BlockSc<false> {
    typedef double T;
    // reciprocal of the pivot to ~1 ulp, returned as (r0, t):  q / w = fma(q*r0, t, q*r0)
#ifndef KRB_OLD_RCP8
    struct Inv { double r; };     // the full reciprocal once (r0 (1 + e + e^2), ~1 ulp), quotients = a * r: 6 instead of 8 FP64 operations per row (r2g: -6.6 %)
#else
    struct Inv { double r0, t; };
#endif
    static __device__ __forceinline__ uint32_t bad_pivot(double w) { return pivot_key(w) >= KRB_BAD_PIVOT_KEY ? KRB_NONFINITE_KEY : 0u; }
    static __device__ __forceinline__ Inv inv(double w) {
        const double r0 = rcp_seed(w);
        const double e = fma(-w, r0, 1.0);
#ifndef KRB_OLD_RCP8
        return Inv{fma(r0, fma(e, e, e), r0)};
    }
    static __device__ __forceinline__ double div(double a, const Inv &i) { return a * i.r; }
#else
        return Inv{r0, fma(e, e, e)};
    }
    static __device__ __forceinline__ double div(double a, const Inv &i) { const double a0 = a * i.r0; return fma(a0, i.t, a0); }
#endif
    static __device__ __forceinline__ double neg(double a) { return -a; }
    static __device__ __forceinline__ double one() { return 1.0; }
    static __device__ __forceinline__ double shfl_up1(double a) { return __shfl_up_sync(0xffffffffu, a, 1); }
};
template <> struct BlockSc<true> {
    typedef double2 T;
    struct Inv { double2 c; };    // conj(w) / |w|^2
    static __device__ __forceinline__ uint32_t bad_pivot(double2 w) { return pivot_key(fma(w.x, w.x, w.y * w.y)) >= KRB_BAD_PIVOT_KEY ? KRB_NONFINITE_KEY : 0u; }
    static __device__ __forceinline__ Inv inv(double2 w) {
        const double ns = fma(w.x, w.x, w.y * w.y);
        const double r = fast_rcp(ns);
        return Inv{make_double2(w.x * r, -w.y * r)};
    }
    static __device__ __forceinline__ double2 div(double2 a, const Inv &i) { return ChainSc<true>::mul(a, i.c); }
    static __device__ __forceinline__ double2 neg(double2 a) { return make_double2(-a.x, -a.y); }
    static __device__ __forceinline__ double2 one() { return make_double2(1.0, 0.0); }
    static __device__ __forceinline__ double2 shfl_up1(double2 a) {
        return make_double2(__shfl_up_sync(0xffffffffu, a.x, 1), __shfl_up_sync(0xffffffffu, a.y, 1));
    }
};

// One GENERIC row: ordered term run, same decoding and summation order as chain_kernel.cuh (bit-identical stamps).
// INV / inv_base / inv_n (KEEP_INV callers only): input planes [inv_base, inv_base + inv_n) are read from INV (lane stride inv_L) instead
// of IN — the caller's own Monte-Carlo array when it already has the kernels' layout, so that it is never copied into the plan's buffer.
template <bool CPLX, bool KEEP_INV = false>
__device__ __forceinline__ void block_eval_terms(const uint4 *__restrict__ dp, int nt, const double *__restrict__ IN, long long L, double w,
                                                 long long prefetch_ahead, typename ChainSc<CPLX>::T &e0, typename ChainSc<CPLX>::T &e1,
                                                 typename ChainSc<CPLX>::T &e2, typename ChainSc<CPLX>::T &e3, typename ChainSc<CPLX>::T &e4,
                                                 const double *__restrict__ INV = nullptr, uint32_t inv_base = 0, uint32_t inv_n = 0, long long inv_L = 0) {
    typedef ChainSc<CPLX> S;
    typedef typename S::T V;
    e0 = S::zero(); e1 = S::zero(); e2 = S::zero(); e3 = S::zero(); e4 = S::zero();
    // KEEP_INV (the pre-stamping kernel of Monte-Carlo plans): a two-terminal element stamps its conductance twice into a node's row
    // (diagonal, then off-diagonal: consecutive terms of the run, stampable.rs:229-244): the second 1.0 / value is the first one again,
    // bit for bit, so it is kept instead of redone (half of the IEEE divisions of a ladder row)
    uint32_t inv_plane = 0xffffffffu;
    double inv_val = 0.0;
    uint4 ahead = make_uint4(0u, 0u, 0u, 0u);
    if (KEEP_INV && nt > 0) ahead = __ldg(dp);      // KEEP_INV: term q + 1 is fetched while term q is evaluated (the stream ends with 4 zero units)
    for (int q = 0; q < nt; ++q) {
        uint4 cur;
        if (KEEP_INV) { cur = ahead; ahead = __ldg(dp + q + 1); }
        else cur = __ldg(dp + q);
        const double K = u4hi(cur);
        const uint32_t kind = (cur.x >> 4) & 0xfu;
        double re = 0.0, im = 0.0;
        if (kind == CK_CONST) re = K;
        else if (kind == CK_JW_K) im = __dmul_rn(w, K);
        else if (KEEP_INV && kind == CK_INV_IN && cur.y == inv_plane) re = inv_val;
        else {
            const double *src = IN + (size_t)cur.y * L;
            if (KEEP_INV && cur.y - inv_base < inv_n) src = INV + (size_t)(cur.y - inv_base) * inv_L;
            const double v = __ldg(src);
            if (prefetch_ahead > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + prefetch_ahead));
            if (kind == CK_IN) re = v;
            else if (kind == CK_INV_IN) {
                // 1.0 / value (stampable.rs:232): __drcp_rn is the same correctly rounded quotient as __ddiv_rn(1.0, v) in half the instructions
                if (KEEP_INV) { re = __drcp_rn(v); inv_plane = cur.y; inv_val = re; }
                else re = __ddiv_rn(1.0, v);
            }
            else if (kind == CK_IN_OVER_K) re = __ddiv_rn(v, K);
            else im = __dmul_rn(w, v);
        }
        if (cur.x & CT_NEG) { re = -re; im = -im; }
        const uint32_t tgt = cur.x & 7u;
        if (KEEP_INV) {   // the pre-stamping kernel: a warp = 32 lanes of ONE row, so the target is warp-uniform and this branches without diverging
            if (tgt == CT_SUB) S::add(e0, re, im);
            else if (tgt == CT_DIAG) S::add(e1, re, im);
            else if (tgt == CT_SUP1) S::add(e2, re, im);
            else if (tgt == CT_SUP2) S::add(e3, re, im);
            else S::add(e4, re, im);
            continue;
        }
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

// ---- pre-stamped rows for plans without a shared dictionary ---------------------------------------------------------------
template <bool CPLX> struct BlockRec;
template <> struct BlockRec<false> {
    struct __align__(16) T { double v[4]; };          // (sub, diag, sup, rhs)
    static __device__ __forceinline__ void load(const T *p, double &a, double &b, double &c, double &d) {
        const double2 lo = __ldcs(reinterpret_cast<const double2 *>(p)), hi = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
        a = lo.x; b = lo.y; c = hi.x; d = hi.y;
    }
};
template <> struct BlockRec<true> {
    struct __align__(16) T { double2 v[4]; };
    static __device__ __forceinline__ void load(const T *p, double2 &a, double2 &b, double2 &c, double2 &d) {
        const double2 *q = reinterpret_cast<const double2 *>(p);
        a = __ldcs(q); b = __ldcs(q + 1); c = __ldcs(q + 2); d = __ldcs(q + 3);
    }
};

struct BlockStampParams {
    const uint2 *rows;           // [M][T] row words (as ChainBlockParams::rows)
    const uint4 *dict, *stream;
    const double *IN;            // [n_inputs][L]
    long long L, lanes;
    int input_freq, slots;       // slots = T * M
    int T;
    void *E, *U2;
    const double *INV;           // planes [inv_base, inv_base + inv_n) come from here (lane stride inv_L); inv_n = 0: everything from IN
    unsigned inv_base, inv_n;
    long long inv_L;
};

// (a) of the north star for Monte-Carlo chain plans: one thread stamps one pivot row of one lane — ordered term run, exact IEEE
// divisions, same summation order as sum_triplets (solver.rs:59-70) — from the lane-interleaved inputs (a warp = 32 consecutive
// lanes of one row: coalesced), and the tile is transposed through shared memory so that every lane receives 32 consecutive
// records (1 / 2 KB) at once.  blockDim = (32, 8); grid = (slots / 32, ceil(lanes / 32)).
// resident CTAs per SM the register allocation must allow.  Measured on ladder_mc (r6b, one gpurun call): 5 CTAs (48 registers, 28 bytes
// of spills) 4.34 ms per step, 6 CTAs (40 registers, 60 bytes) 4.75 ms, no bound (the allocator takes what it likes) 4.99 ms; the
// complex instantiation keeps its natural 80 registers (3 CTAs)
#ifndef KRB_STAMP_MIN_BLOCKS
#define KRB_STAMP_MIN_BLOCKS (CPLX ? 3 : 5)
#endif
template <bool CPLX>
__global__ void __launch_bounds__(256, KRB_STAMP_MIN_BLOCKS) krb_block_stamp_rows(const BlockStampParams p) {
    typedef ChainSc<CPLX> S;
    typedef typename S::T V;
    constexpr int FIX_UNITS = CPLX ? 4 : 2;
    constexpr int RT = CPLX ? 16 : 32;               // rows per tile (shared memory: 4 * RT * 33 elements)
    __shared__ V tile[RT * 4][33];
    const int lx = threadIdx.x;
    const long long lane0 = (long long)blockIdx.y * 32, lane = lane0 + lx;
    const bool live = lane < p.lanes;
    const double *__restrict__ IN = p.IN + (live ? lane : 0);
    const double *__restrict__ INV = p.INV + (live ? lane : 0);
    double w = 0.0;
    if (CPLX && p.input_freq >= 0) w = __dmul_rn(2.0 * 3.14159265358979323846264338327950288, __ldg(IN + (size_t)p.input_freq * p.L));
    const int slot0 = blockIdx.x * RT;
    for (int r = threadIdx.y; r < RT; r += blockDim.y) {
        const int slot = slot0 + r;                   // = j * T + t
        const uint2 rw = __ldg(p.rows + slot);
        V e0, e1, e2, e3 = S::zero(), e4;
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
            block_eval_terms<CPLX, true>(p.stream + (rw.x >> 1), (int)((rw.y >> BR_NT_SHIFT) & BR_NT_MASK), IN, p.L, w, 0, e0, e1, e2, e3, e4,
                                         INV, p.inv_base, p.inv_n, p.inv_L);
            if ((rw.y & BR_HAS_U2) && live) reinterpret_cast<V *>(p.U2)[(size_t)lane * p.T + slot % p.T] = e3;
        }
        tile[r * 4 + 0][lx] = e0; tile[r * 4 + 1][lx] = e1; tile[r * 4 + 2][lx] = e2; tile[r * 4 + 3][lx] = e4;
    }
    __syncthreads();
    V *__restrict__ E = reinterpret_cast<V *>(p.E);
    for (int ll = threadIdx.y; ll < 32; ll += blockDim.y) {
        const long long ln = lane0 + ll;
        if (ln >= p.lanes) break;
        V *dst = E + ((size_t)ln * (size_t)p.slots + (size_t)slot0) * 4;
        for (int q = lx; q < RT * 4; q += 32) __stcs(dst + q, tile[q][ll]);
    }
}

// SD (shared dictionary): the plan's unique FIXED records and its GENERIC rows together fit KRB_BLOCK_DMAX entries (ladders,
// RC lines, swept sources: a handful).  Every row then reads its stamped (sub, diag, sup, rhs) from a shared-memory entry:
// FIXED entries are lane-invariant (real) or refreshed per lane with the lane's w (complex: (re, w*K) is formed once per
// entry instead of once per row), GENERIC entries are evaluated per lane by one thread each — for the NEXT lane, by threads of
// the upper-half warps while the lower warps reduce the separator system, so the main loop is branch-free and no warp waits
// for a term run.
// !SD: FIXED rows read the global dictionary (L1), warps owning GENERIC rows decode them inline (Monte-Carlo plans).
template <bool CPLX, int T, int M, bool SD>
__global__ void __launch_bounds__(T, T >= 512 ? 1 : 2) krb_chain_block_kernel(const ChainBlockParams p) {
    typedef ChainSc<CPLX> S;
    typedef BlockSc<CPLX> B;
    typedef typename S::T V;
    constexpr int MI = M - 1;               // interior rows per thread
    constexpr int XS = (M & 1) ? M : M + 1; // stride of a thread's l^ column (odd: conflict-free)
    constexpr int FIX_UNITS = CPLX ? 4 : 2;
    constexpr int NW = T / 32;
    constexpr int DMAX = BlockDmax<CPLX>::value;
    constexpr int RPAD = 8;                 // the odd-parity halves of the reduction buffers start 8 elements later: threads t, t+1 hit different banks
    extern __shared__ __align__(16) unsigned char krb_block_smem[];
    V *xq = reinterpret_cast<V *>(krb_block_smem);   // [T][XS]: l^_j
    V *xo = xq + T * XS;                               // [T * M + 2]: the lane's results in pivot-row order (+ one element of alignment shift)
    V *red = xo + T * M + 2;                           // [2][3][T] + pads: separator rows (a, c, d) by parity, then the reduced system
    V *fst = red + 3 * T + RPAD;                       // [3][T] + pad: (P, R, Q) of each block's first row; aliases the reduced system
    V *sd = red + 6 * T + 2 * RPAD;                    // SD: [DMAX][4] stamped (sub, diag, sup, rhs) per entry
    V *pe3 = sd + 4 * DMAX;                            // SD: [DMAX] second super-diagonal of a GENERIC entry
    const int t = threadIdx.x;
    V *mycol = xq + t * XS;
    const long long L = p.L;
    const int n = p.n;

    // The rows a thread owns do not depend on the lane: dictionary / entry indices are packed two per register once.
    constexpr int PKW = SD ? 8 : 16;         // SD: entry indices < KRB_BLOCK_DMAX = 256, four per register
    constexpr int PKN = 32 / PKW;
    uint32_t pk[(M + PKN - 1) / PKN];
    bool my_generic = false;
    int ju = -1, gu = 0;                     // SD: local index / entry of this block's row with a second super-diagonal
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const uint2 rw = __ldg(p.rows + j * T + t);
        uint32_t idx = rw.x >> 1;
        if (!(rw.x & 1u)) {
            my_generic = true;
            if (SD) idx = (uint32_t)p.n_dict + (rw.y >> BR_SLOT_SHIFT);
            if (rw.y & BR_HAS_U2) { ju = j; gu = (int)idx; }
        }
        if (j % PKN) pk[j / PKN] |= (idx & ((1u << PKW) - 1u)) << ((j % PKN) * PKW); else pk[j / PKN] = idx & ((1u << PKW) - 1u);
    }
    const bool warp_generic = __any_sync(0xffffffffu, my_generic);
    const bool warp_u2 = __any_sync(0xffffffffu, ju >= 0);
    const uint32_t u2mask = __reduce_or_sync(0xffffffffu, ju >= 0 ? 1u << ju : 0u);   // local indices of the warp's u2 rows
    // SD: every thread of the warp stamps its M rows from ONE entry (the periodic interior of a ladder / line): the entry is
    // read once per lane instead of once per row
    bool one_entry = SD;
#pragma unroll
    for (int j = 1; j < M; ++j)
        one_entry = one_entry && ((pk[j / PKN] >> ((j % PKN) * PKW)) & ((1u << PKW) - 1u)) == (pk[0] & ((1u << PKW) - 1u));
    const bool warp_one_entry = SD && !warp_u2 && __all_sync(0xffffffffu, one_entry);
    (void)warp_generic; (void)warp_u2; (void)u2mask; (void)gu; (void)sd; (void)pe3;

    // SD: (re)build the entries for `lane`.  The owners are the threads of the UPPER half of the warps - the ones that sit out the
    // reduction of the separator system (step 2) - so that the global loads of a GENERIC entry (term run + per-lane inputs, two
    // to four dependent L2 round trips) run beside that reduction instead of in front of it, where every other warp of the CTA
    // waited for the owner (ncu, c64: 11.7 % of all warp samples at the barrier behind it).  Owner s = t - T/2 takes entries
    // e(s), e(s) + T/2, ... with e(s) = (s % 32) * (NW/2) + s / 32 (spread over the upper warps).  What an owner needs is looked
    // up again for every lane (from an opaque copy of t): kept across the lane loop it would cost registers of the
    // elimination's budget, i.e. local-memory traffic in every thread.
    static_assert(DMAX <= 256, "8-bit entry indices");
    constexpr int TH = T / 2, TQ = T / 4;
    auto refresh = [&](long long lane, double w, long long ahead) {
        int tt = t;
        asm volatile("" : "+r"(tt));
        if (tt < TH) return;
        tt -= TH;
        for (int my_entry = (tt & 31) * (NW / 2) + (tt >> 5); my_entry < p.n_dict + p.n_gen; my_entry += TH) {
            if (my_entry < p.n_dict) {
                if (CPLX) {
                    const uint4 *g = p.dict + (size_t)my_entry * FIX_UNITS;
                    const uint4 c0 = __ldg(g), c1 = __ldg(g + 1), c2 = __ldg(g + 2), c3 = __ldg(g + 3);
                    sd[4 * my_entry + 0] = S::make(u4lo(c0), __dmul_rn(w, u4lo(c2)));
                    sd[4 * my_entry + 1] = S::make(u4hi(c0), __dmul_rn(w, u4hi(c2)));
                    sd[4 * my_entry + 2] = S::make(u4lo(c1), __dmul_rn(w, u4lo(c3)));
                    sd[4 * my_entry + 3] = S::make(u4hi(c1), 0.0);
                }
            } else {
                const int row = __ldg(p.gen_rows + (my_entry - p.n_dict));
                const uint2 rw = __ldg(p.rows + (row % M) * T + row / M);
                V e0, e1, e2, e3, e4;
                block_eval_terms<CPLX>(p.stream + (rw.x >> 1), (int)((rw.y >> BR_NT_SHIFT) & BR_NT_MASK), p.IN + lane, L, w, ahead, e0, e1, e2, e3, e4);
                sd[4 * my_entry + 0] = e0; sd[4 * my_entry + 1] = e1; sd[4 * my_entry + 2] = e2; sd[4 * my_entry + 3] = e4;
                pe3[my_entry] = e3;
            }
        }
    };
    // barriers among the lower warps only (ids 1, 2): the warps that have nothing to reduce do not take part
    auto bar_lower = [](int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); };
    const double TWO_PI = 2.0 * 3.14159265358979323846264338327950288;
    if (SD) {
        if (!CPLX)
            for (int d = t; d < p.n_dict; d += T) {
                const uint4 *g = p.dict + (size_t)d * FIX_UNITS;
                const uint4 c0 = __ldg(g), c1 = __ldg(g + 1);
                sd[4 * d + 0] = S::make(u4lo(c0), 0.0); sd[4 * d + 1] = S::make(u4hi(c0), 0.0);
                sd[4 * d + 2] = S::make(u4lo(c1), 0.0); sd[4 * d + 3] = S::make(u4hi(c1), 0.0);
            }
        if ((long long)blockIdx.x < p.lanes) {
            double w0 = 0.0;
            if (CPLX && p.input_freq >= 0) w0 = __dmul_rn(TWO_PI, __ldg(p.IN + blockIdx.x + (size_t)p.input_freq * L));
            refresh(blockIdx.x, w0, (long long)blockIdx.x + gridDim.x < p.lanes ? (long long)gridDim.x : 0);
        }
        __syncthreads();
    }

    for (long long lane = blockIdx.x; lane < p.lanes; lane += gridDim.x) {
        const double *__restrict__ IN = p.IN + lane;
        const long long next = lane + gridDim.x;
        double w = 0.0, f_next = 1.0;                // w = (2.0 * PI) * f, rounded once as the reference does
        if (CPLX && p.input_freq >= 0) {
            if (!SD) w = __dmul_rn(TWO_PI, __ldg(IN + (size_t)p.input_freq * L));
            else if (next < p.lanes) f_next = __ldg(p.IN + next + (size_t)p.input_freq * L);   // consumed after the first barrier
        }
        // keep what is derived from the packed indices (20 shared-memory addresses) inside the loop: hoisted, it is spilled
#pragma unroll
        for (int i = 0; i < (M + PKN - 1) / PKN; ++i) asm volatile("" : "+r"(pk[i]));
        uint32_t grow = 0;                           // running maximum of growth_key(u^) (see pivot_key / growth_key above)
        V u[MI], y[MI];
        V lprev = S::zero(), pu2 = S::zero(), u2v = S::zero();
        V sepA = S::zero(), sepB = S::zero(), sepD = S::zero(), sepS = S::zero();

        // ---- 1. stamp + top-down elimination of the interior rows; separator row partially reduced ----
        // one interior row:  w_k = diag - sub u^_{k-1},  u^ = sup1'/w_k,  y^ = (rhs - sub y^_{k-1})/w_k,  l^ = -sub l^_{k-1}/w_k
        V accP = S::zero(), accQ = S::zero(), fA = B::one(), fAprev = B::one(), fB = S::zero();
        auto interior = [&](int j, V e0, V e1, V s1, V e4) {
            V wk, yk, lk;
            if (j == 0) { wk = e1; yk = e4; lk = e0; }
            else {
                wk = S::fnma(e1, e0, u[j > 0 ? j - 1 : 0]);
                yk = S::fnma(e4, e0, y[j > 0 ? j - 1 : 0]);
                lk = B::neg(S::mul(e0, lprev));
            }
            const typename B::Inv iv = B::inv(wk);
            u[j] = B::div(s1, iv);
            grow = max(grow, growth_key(u[j]));
            y[j] = B::div(yk, iv);
            lprev = B::div(lk, iv);
            mycol[j] = lprev;
            // x_0 = accP - accQ s_{t-1} + fA x_j + fB x_{j+1}: substitute row j (replaces a bottom-up pass over the block)
            if (j == 0) { accP = y[0]; accQ = lprev; fAprev = B::one(); fA = B::neg(u[0]); }
            else {
                accP = S::fnma(accP, B::neg(fA), y[j]);
                accQ = S::fnma(accQ, B::neg(fA), lprev);
                fAprev = fA;
                fA = S::fnma(fB, fA, u[j]);
            }
            fB = S::zero();
            return iv;
        };
        // separator row: sub * x_{last interior} substituted from the forward-eliminated last interior row
        auto separator = [&](V e0, V e1, V e2, V e4) {
            sepA = B::neg(S::mul(e0, lprev));
            sepB = S::fnma(e1, e0, u[MI - 1]);
            sepD = S::fnma(e4, e0, y[MI - 1]);
            sepS = e2;
        };
        if (SD) {
            if (warp_one_entry) {
                const V *e = sd + 4 * (pk[0] & ((1u << PKW) - 1u));
                const V c0 = e[0], c1 = e[1], c2 = e[2], c3 = e[3];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    if (j < MI) interior(j, c0, c1, c2, c3);
                    else separator(c0, c1, c2, c3);
                }
            } else {
                // entries fetched exactly one row ahead (the fence keeps the compiler from hoisting more of them: each costs
                // 4 values of the register budget that u^ / y^ need).  Rows with a second super-diagonal (at most one per
                // thread) are found through the warp's mask of their local indices: a warp-uniform test per row, the
                // per-thread comparison only on the one or two rows where some thread of the warp needs it.
                auto entry = [&](int j) { return sd + 4 * ((pk[j / PKN] >> ((j % PKN) * PKW)) & ((1u << PKW) - 1u)); };
                const V *e = entry(0);
                V c0 = e[0], c1 = e[1], c2 = e[2], c3 = e[3];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    V n0 = c0, n1 = c1, n2 = c2, n3 = c3;
                    if (j + 1 < M) { e = entry(j + 1); n0 = e[0]; n1 = e[1]; n2 = e[2]; n3 = e[3]; }
                    asm volatile("" ::: "memory");
                    if (j < MI) {
                        V s1 = c2;
                        if (j > 0 && ((u2mask >> (j - 1)) & 1u)) {
                            if (j - 1 == ju) s1 = S::fnma(s1, c0, pu2);            // fill from the row above's second super-diagonal
                        }
                        const typename B::Inv iv = interior(j, c0, c1, s1, c3);
                        if ((u2mask >> j) & 1u) {
                            if (j == ju) { pu2 = B::div(pe3[gu], iv); u2v = pu2; fB = B::neg(S::mul(fAprev, pu2)); }
                        }
                    } else separator(c0, c1, c2, c3);
                    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
                }
            }
        } else {
            // !SD (Monte-Carlo plans: most rows carry per-lane terms): the rows were stamped by krb_block_stamp_rows into
            // E[lane][M][T][4] — thousands of term runs with IEEE divisions per lane belong in a thread-per-(row, lane) kernel with
            // coalesced lane-interleaved inputs, not on this kernel's dependent chain.  A row is one 32-byte (real) / 64-byte
            // (complex) record per thread, a warp reads 1 / 2 KB contiguous; records are fetched one row ahead and the next lane's
            // block is prefetched into L2 (296 resident lanes x 172 KB fit it).
            typedef typename BlockRec<CPLX>::T Rec;
            int tt = t;
            asm volatile("" : "+r"(tt));
            const Rec *__restrict__ Eb = reinterpret_cast<const Rec *>(p.E) + (size_t)lane * (size_t)(T * M) + tt;
            if (next < p.lanes) {
                const char *nb = reinterpret_cast<const char *>(reinterpret_cast<const Rec *>(p.E) + (size_t)next * (size_t)(T * M));
                constexpr int LINES = (int)(((size_t)T * M * sizeof(Rec) + 127) / 128);
                for (int ln = tt; ln < LINES; ln += T) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + (size_t)ln * 128));
            }
            V e3v = S::zero();
            if (ju >= 0) e3v = __ldg(reinterpret_cast<const V *>(p.U2) + (size_t)lane * T + tt);
            V c0, c1, c2, c3;
            BlockRec<CPLX>::load(Eb, c0, c1, c2, c3);
#pragma unroll
            for (int j = 0; j < M; ++j) {
                V n0 = c0, n1 = c1, n2 = c2, n3 = c3;
                if (j + 1 < M) BlockRec<CPLX>::load(Eb + (j + 1) * T, n0, n1, n2, n3);
                asm volatile("" ::: "memory");
                if (j < MI) {
                    V s1 = c2;
                    if (j > 0 && ((u2mask >> (j - 1)) & 1u)) {
                        if (j - 1 == ju) s1 = S::fnma(s1, c0, pu2);            // fill from the row above's second super-diagonal
                    }
                    const typename B::Inv iv = interior(j, c0, c1, s1, c3);
                    if ((u2mask >> j) & 1u) {
                        if (j == ju) { pu2 = B::div(e3v, iv); u2v = pu2; fB = B::neg(S::mul(fAprev, pu2)); }
                    }
                } else separator(c0, c1, c2, c3);
                c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            }
        }

        // (P, R, Q) of the block's first row were accumulated top-down inside step 1: there is no bottom-up pass
        fst[t] = accP; fst[T + t] = B::neg(fA); fst[2 * T + t] = accQ;
        __syncthreads();

        // ---- 2. separator system: A s_{t-1} + B s_t + C s_{t+1} = D ----
        V a, c, d;
        {
            V pn = S::zero(), rn = S::zero(), qn = S::zero();
            if (t + 1 < T) { pn = fst[t + 1]; rn = fst[T + t + 1]; qn = fst[2 * T + t + 1]; }
            const V bb = S::fnma(sepB, sepS, qn);
            const V cc = B::neg(S::mul(sepS, rn));
            const V dd = S::fnma(sepD, sepS, pn);
            const typename B::Inv iv = B::inv(bb);
            a = B::div(sepA, iv); c = B::div(cc, iv); d = B::div(dd, iv);
            grow = max(grow, max(growth_key(a), growth_key(c)));
        }
        V st, sm;
        // Two levels of cyclic reduction, then parallel cyclic reduction on the surviving rows only: the T/2 threads of the lower
        // warps own compact row t' = (2t'+1), the lowest T/4 the rows that survive the second level; rows are stored
        // de-interleaved by parity so that every access stays unit-stride.  Against PCR on all T rows the shared-memory traffic of
        // the separator system drops by ~65 % (measured: one level -13 % (c64) / -5 % (f64) run time, the second another -4 % /
        // -0.6 %).  The upper warps rebuild the dictionary entries for the next lane meanwhile (nobody reads them any more for this
        // lane) and meet the others at the barrier in front of step 3.
        V *ini = red;                                  // [parity][a, c, d][TH] (odd half RPAD later); the odd part's `a` row later holds the solution
        V *inio = ini + 3 * TH + RPAD;                 // the odd rows
        V *sol = inio;
        const int hk = t >> 1, par = t & 1;
        { V *w_ = par ? inio : ini; w_[hk] = a; w_[TH + hk] = c; w_[2 * TH + hk] = d; }
        __syncthreads();
        auto reduce_row = [&](V am, V cm, V dm, V ap, V cp, V dq, bool need_ac) {
            const V b2 = S::fnma(S::fnma(B::one(), a, cm), c, ap);
            const V d2 = S::fnma(S::fnma(d, a, dm), c, dq);
            const typename B::Inv iv = B::inv(b2);
            if (!need_ac) grow = max(grow, B::bad_pivot(b2));
            if (need_ac) {
                const V a2 = B::neg(S::mul(a, am));
                const V c2 = B::neg(S::mul(c, cp));
                a = B::div(a2, iv); c = B::div(c2, iv);
                grow = max(grow, max(growth_key(a), growth_key(c)));
            }
            d = B::div(d2, iv);
        };
        if (t >= TH) {
            if (SD && next < p.lanes) refresh(next, CPLX ? __dmul_rn(TWO_PI, f_next) : 0.0, next + gridDim.x < p.lanes ? (long long)gridDim.x : 0);
        } else {
            V *ini2 = fst;                             // [parity][a, c, d][TQ] (aliases fst, odd half RPAD later); odd part's `a` row: level-2 solution
            V *ini2o = ini2 + 3 * TQ + RPAD;
            V *sol2 = ini2o;
            V *pa = ini2o + 3 * TQ, *pb = pa + 3 * TQ; // double-buffered [a, c, d][TQ]
            {
                a = inio[t]; c = inio[TH + t]; d = inio[2 * TH + t];                   // row 2t+1
                V ap = S::zero(), cp = S::zero(), dq = S::zero();
                if (t + 1 < TH) { ap = ini[t + 1]; cp = ini[TH + t + 1]; dq = ini[2 * TH + t + 1]; }   // row 2t+2
                reduce_row(ini[t], ini[TH + t], ini[2 * TH + t], ap, cp, dq, true);    // row 2t
                { V *w_ = par ? ini2o : ini2; w_[hk] = a; w_[TQ + hk] = c; w_[2 * TQ + hk] = d; }
            }
            bar_lower(1, TH);
            if (t < TQ) {
                a = ini2o[t]; c = ini2o[TQ + t]; d = ini2o[2 * TQ + t];                // compact row 2t+1
                V ap = S::zero(), cp = S::zero(), dq = S::zero();
                if (t + 1 < TQ) { ap = ini2[t + 1]; cp = ini2[TQ + t + 1]; dq = ini2[2 * TQ + t + 1]; }
                reduce_row(ini2[t], ini2[TQ + t], ini2[2 * TQ + t], ap, cp, dq, true);
                pa[t] = a; pa[TQ + t] = c; pa[2 * TQ + t] = d;
                bar_lower(2, TQ);
#ifndef KRB_OLD_SMEM_PCR
                // real plans (TQ = 64): a third cyclic-reduction level leaves 32 rows, which ONE warp reduces by parallel cyclic reduction
                // in registers (neighbour rows by shuffle: no shared-memory round trips, no barriers) instead of six shared-memory PCR steps
                // with a two-warp barrier each; the even rows of the 64 follow by substitution
                if (!CPLX && TQ == 64) {
                    if (t < 32) {
                        const V ea = pa[2 * t], ec = pa[TQ + 2 * t], ed = pa[2 * TQ + 2 * t];            // row 2t: kept for the substitution
                        a = pa[2 * t + 1]; c = pa[TQ + 2 * t + 1]; d = pa[2 * TQ + 2 * t + 1];
                        V ap3 = S::zero(), cp3 = S::zero(), dq3 = S::zero();
                        if (2 * t + 2 < TQ) { ap3 = pa[2 * t + 2]; cp3 = pa[TQ + 2 * t + 2]; dq3 = pa[2 * TQ + 2 * t + 2]; }
                        reduce_row(ea, ec, ed, ap3, cp3, dq3, true);
#pragma unroll
                        for (int s = 1; s < 32; s <<= 1) {
                            V am = BlockShfl<CPLX>::up(a, s), cm = BlockShfl<CPLX>::up(c, s), dm = BlockShfl<CPLX>::up(d, s);
                            V ap2 = BlockShfl<CPLX>::down(a, s), cp2 = BlockShfl<CPLX>::down(c, s), dq2 = BlockShfl<CPLX>::down(d, s);
                            if (t < s) { am = S::zero(); cm = S::zero(); dm = S::zero(); }
                            if (t + s >= 32) { ap2 = S::zero(); cp2 = S::zero(); dq2 = S::zero(); }
                            reduce_row(am, cm, dm, ap2, cp2, dq2, 2 * s < 32);
                        }
                        V xm = BlockShfl<CPLX>::up(d, 1);
                        if (t == 0) xm = S::zero();
                        sol2[2 * t] = S::fnma(S::fnma(ed, ea, xm), ec, d);
                        sol2[2 * t + 1] = d;
                    }
                } else
#endif
#pragma unroll
                for (int s = 1; s < TQ; s <<= 1) {
                    V am = S::zero(), cm = S::zero(), dm = S::zero(), ap2 = S::zero(), cp2 = S::zero(), dq2 = S::zero();
                    if (t >= s) { am = pa[t - s]; cm = pa[TQ + t - s]; dm = pa[2 * TQ + t - s]; }
                    if (t + s < TQ) { ap2 = pa[t + s]; cp2 = pa[TQ + t + s]; dq2 = pa[2 * TQ + t + s]; }
                    reduce_row(am, cm, dm, ap2, cp2, dq2, 2 * s < TQ);
                    if (2 * s < TQ) { pb[t] = a; pb[TQ + t] = c; pb[2 * TQ + t] = d; }
                    else sol2[t] = d;
                    bar_lower(2, TQ);
                    V *sw = pa; pa = pb; pb = sw;
                }
            }
            bar_lower(1, TH);                           // the level-2 solution is visible to the T/2 compact rows
            V x1 = sol2[hk];                            // compact rows: odd ones are solved, even ones by substitution
            if (par == 0) {
                const V xl2 = hk > 0 ? sol2[hk - 1] : S::zero();
                x1 = S::fnma(S::fnma(ini2[2 * TQ + hk], ini2[hk], xl2), ini2[TQ + hk], x1);
            }
            sol[t] = x1;
            // KRB_BLOCK_WAIT_BULK: the previous lane's bulk copy has had a whole elimination + reduction to read xo; make sure of it
            // before the barrier that lets every thread overwrite xo
            if (t == KRB_BLOCK_ISSUER(NW)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
        // even rows by substitution from their own (a, c, d); a thread's left separator: odd rows get it from the lane below
        st = sol[hk];
        {
            const V xl = (par == 0 && hk > 0) ? sol[hk - 1] : S::zero();
            if (par == 0) st = S::fnma(S::fnma(ini[2 * TH + hk], ini[hk], xl), ini[TH + hk], st);
            const V below = B::shfl_up1(st);
            sm = par ? below : xl;
        }

        // ---- 3. interiors into the result buffer, then ONE bulk asynchronous copy of the lane's vector ----
        // back substitution  x_j = (y^_j - l^_j s_{t-1}) - u^_j x_{j+1} (- u2^_j x_{j+2}),  x_{M-1} = s_t
        V *__restrict__ out = reinterpret_cast<V *>(p.OUT) + (size_t)lane * (size_t)(p.n_select > 0 ? p.n_select : n);
        // real: the lane's first byte may sit at 8 mod 16; the result buffer is shifted by one element then, so that shared and
        // global addresses stay congruent mod 16 (bulk copies need 16-byte aligned addresses and sizes)
        const int shift = CPLX ? 0 : (int)((reinterpret_cast<size_t>(out) >> 3) & 1);
        {
            int tt = t;
            asm volatile("" : "+r"(tt));
            V *__restrict__ xcol = xo + shift + tt * M;
            V xn = st, x2 = st;
            const uint32_t m4 = u2mask;
            V lj = mycol[MI - 1];                     // l^ of row j - 1 is loaded while row j is computed (c64: -2.5 %)
#ifndef KRB_OLD_NO_BSPLIT
            if (m4 == 0u) {                           // (r2g: -2.1 %) no row of this warp has a second super-diagonal: a test-free loop
#pragma unroll
                for (int j = MI - 1; j >= 0; --j) {
                    V ln = lj;
                    if (j > 0) ln = mycol[j - 1];
                    const V xj = S::fnma(S::fnma(y[j], lj, sm), u[j], xn);
                    lj = ln;
                    xcol[j] = xj;
                    xn = xj;
                }
            } else
#endif
            {
#pragma unroll
                for (int j = MI - 1; j >= 0; --j) {
                    V ln = lj;
                    if (j > 0) ln = mycol[j - 1];
                    V xj = S::fnma(S::fnma(y[j], lj, sm), u[j], xn);
                    lj = ln;
                    if ((m4 >> j) & 1u) {
                        if (j == ju) xj = S::fnma(xj, u2v, x2);
                    }
                    xcol[j] = xj;
                    x2 = xn; xn = xj;
                }
            }
            xcol[MI] = st;
        }
        int status = grow >= KRB_NONFINITE_KEY ? KRB_LANE_ZERO_PIVOT : grow >= KRB_GROWTH_KEY ? KRB_LANE_INACCURATE : 0;
        if (!S::finite(st)) status |= KRB_LANE_NONFINITE;
        if (status && p.status) atomicOr(p.status + lane, status);
        const V *__restrict__ xv = xo + shift;        // pivot row i at xv[i]
#ifndef KRB_OLD_FULL_BARRIER
        // (r2g: -1 %) only the warp that issues the bulk copy waits for the vector; the others signal and start the next lane
        constexpr int ISSUER = (NW / 2 - 1) * 32;     // a thread of a warp with periodic interior rows (the fast elimination path)
        if (p.bulk_store) {
            if ((t >> 5) == (ISSUER >> 5)) asm volatile("bar.sync 3, %0;" ::"r"(T) : "memory");
            else asm volatile("bar.arrive 3, %0;" ::"r"(T) : "memory");
        } else __syncthreads();
#else
        constexpr int ISSUER = 0;
        __syncthreads();                              // the whole vector is in xo
#endif
        if (p.bulk_store) {
            if (t == ISSUER) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the async proxy
                const int head = shift;                                    // real, odd-aligned lane: element 0 by a plain store
                const int cnt = n - head;
                const int tail = CPLX ? 0 : (cnt & 1);                     // real: an odd count leaves one element for a plain store
                if (head) __stcs(out, xv[0]);
                if (tail) __stcs(out + (n - 1), xv[n - 1]);
                const uint32_t bytes = (uint32_t)(cnt - tail) * (uint32_t)sizeof(V);
                if (bytes) {
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(xv + head);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + head), "r"(src), "r"(bytes) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (p.n_select > 0) {
            // extension: only the selected unknowns leave the chip
            for (int i = t; i < p.n_select; i += T) __stcs(out + i, xv[__ldg(p.select + i)]);
        } else {
            // per-thread delivery (permuted pivot order, or a caller's buffer the bulk copy cannot address): pivot row i is at xv[i],
            // consecutive threads read consecutive elements (conflict-free) and store full-width runs
            int tt = t;
            asm volatile("" : "+r"(tt));
            if (p.unknown_of_row == nullptr) {
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const int i = tt + k * T;
                    if (i < n) __stcs(out + i, xv[i]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const int i = tt + k * T;
                    if (i < n) __stcs(out + __ldg(p.unknown_of_row + i), xv[i]);
                }
            }
        }
        // no barrier here: the next lane's elimination touches only the threads' own l^ columns; its separator buffers were last
        // read before the barrier in front of step 3, and xo is rewritten only in the next lane's step 3, three barriers from now
        // (a pending bulk copy is waited for just before that barrier, see KRB_BLOCK_WAIT_BULK)
    }
    if (t == KRB_BLOCK_ISSUER(NW)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the copies that read it
}

}  // namespace krb
