// sm_100a kernels of the front executor: batched supernodal multifrontal LU over independent solve lanes, for circuits
// whose factor is far too large for the one-thread-per-lane executors (RLC meshes: n ~ 1e5, nnz(L+U) ~ 1e7 per lane).
//
// Replaces, per lane, faer's sp_lu() + solve() (op.rs:48-58, dc.rs:81-90, ac.rs:67-95, transient.rs:74-83) on the
// shared static pattern.  Parallelism: grid = (fronts of one tree level, lanes); one CTA factorises one dense front of
// one lane.  Layout: lane-major — every lane owns a contiguous factor store, inside it fronts are laid out in level order
// and each front's L block is column-major, so a CTA streams its front with unit stride.
//
//   krb_front_expand   per-lane element values EV[lane][element] (Monte-Carlo inputs or nominal)
//   krb_front_stamp    (a) stamps the CSC value array A[lane][entry] in stamp order (== sum_triplets, solver.rs:59-70)
//   krb_front_rhs      right-hand side in pivot order incl. the companion-model history terms (stampable.rs:470-505, 806-815)
//   krb_front_factor   (b) assemble (stamps + children's contribution blocks), dense partial LU, write L/U panels + contribution block
//                      - fronts that fit in shared memory: right-looking rank-1 updates on the staged block
//                      - larger fronts: blocked (panel in shared memory, 4x2 register tiles on the trailing block in L2/HBM)
//   krb_front_forward  (c) level-scheduled forward substitution  (leaves -> root)
//   krb_front_backward (c) level-scheduled backward substitution (root -> leaves), delivers x point-major
//   krb_front_trap_state  trapezoidal capacitor currents carried between steps
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.hpp"
#include "scalar.cuh"

namespace krb {

struct FrontDesc {            // one per front, 64 B
    int p, q, piv0, idx_off;  // idx_off: into idx_unknown / rel
    int child_begin, child_end, asm_begin, asm_end;
    long long lu_off, cb_off, fw_off;   // fw_off < 0: the front is staged in shared memory
    int cbv_off, pad;         // pad: column-block width of krb_front_factor_ll for this front (its launch's shared memory)
};
struct FrontChild {           // one per tree edge, indexed like the child list: what a parent needs from the child, 32 B
    int q, idx_off, cbv_off, pad;
    long long cb_off, pad2;
};
struct FrontElem {            // one per circuit element, 40 B
    double value, ac_amp;
    int ip, im, ic, in_plane; // in_plane >= 0: per-lane value from input plane
    int wave_col, state_idx;  // column of the per-step source table; trapezoidal state column
};
struct FrontTerm { int elem; int kind; };   // kind = SrcKind | (neg << 8)

struct FrontParams {
    const FrontDesc *fronts;
    const int *list;          // fronts of this launch
    const int *idx_unknown, *rel, *child, *asm_ent, *asm_pos, *col_order;
    const FrontChild *childrec;
    const void *A;            // [lanes][a_stride]
    void *LU, *CB, *FW, *CBV, *BP;   // per-lane stores
    long long a_stride, lu_stride, cb_stride, fw_stride, cbv_stride;
    int n;
    int smem_elems;           // capacity of the dynamic shared memory in elements of T
    int use_dmma;             // real f64, 512-thread classes: trailing update on the FP64 tensor cores
    int *status;
    long long lanes;          // lanes of this launch (WARP-mode kernels bound-check their lane)
    int *IPIV;                // [lanes][n] row chosen for pivot k inside its front (local row index; == k when no swap)
    void *X;                  // result rows, point-major: X[lane * x_lane_stride + x_row_off + unknown]
    long long x_lane_stride, x_row_off;
    int stage_off;            // solves: > 0 = the front's factor block is first copied to shared memory (cp.async, every load in flight at
                              // once: ONE memory round trip per front instead of one per 8-column group) at this offset of the slot
};

struct FrontStampParams {
    const FrontElem *elems;
    const FrontTerm *terms;
    const int *term_ptr;      // CSR: matrix entry (stamp) or pivot position (rhs) -> terms
    int count;                // entries / positions
    int n_elems;
    const double *IN;         // [n_inputs][L] lane-interleaved inputs
    long long L;
    double *EV;               // [lanes][n_elems]
    void *OUT;                // A[lanes][count] or BP[lanes][count]
    int input_freq;
    double h_comp;
    const double *Bstep;      // [rows][2][n_wave]
    int n_wave, row;
    const double *Xprev;      // previous result row (real): Xprev[lane * x_lane_stride + unknown]
    long long x_lane_stride;
    double *ST;               // [lanes][n_state] trapezoidal state
    int n_state;
};

__global__ void krb_front_expand(const FrontStampParams p) {
    const long long lane = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_elems) return;
    const FrontElem el = p.elems[e];
    p.EV[lane * p.n_elems + e] = el.in_plane >= 0 ? p.IN[(size_t)el.in_plane * p.L + lane] : el.value;
}

template <bool CPLX>
__global__ void krb_front_stamp(const FrontStampParams p) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const long long lane = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.count) return;
    const double *ev = p.EV + lane * p.n_elems;
    T acc = S::zero();
    const int t0 = p.term_ptr[e], t1 = p.term_ptr[e + 1];
    for (int t = t0; t < t1; t++) {
        const FrontTerm tm = p.terms[t];
        const double v = ev[tm.elem];
        T c;
        switch (tm.kind & 0xff) {
        case S_ONE: c = S::from_real(1.0); break;
        case S_G: c = S::from_real(__ddiv_rn(1.0, v)); break;                       // stampable.rs:232
        case S_VAL: c = S::from_real(v); break;
        case S_VOH: c = S::from_real(__ddiv_rn(v, p.h_comp)); break;                // stampable.rs:449, 781
        case S_JWV: {                                                                // (0, 2.0*PI*f*value)  stampable.rs:379-382
            const double w = __dmul_rn(2.0 * 3.14159265358979323846264338327950288, p.IN[(size_t)p.input_freq * p.L + lane]);
            c = S::mulj(S::from_real(__dmul_rn(w, v)));
            break;
        }
        default: c = S::zero(); break;
        }
        if (tm.kind & 0x100) c = S::neg(c);
        acc = (t == t0) ? c : S::add(acc, c);
    }
    reinterpret_cast<T *>(p.OUT)[lane * p.count + e] = acc;
}

template <bool CPLX>
__global__ void krb_front_rhs(const FrontStampParams p) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const long long lane = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.count) return;
    const double *ev = p.EV + lane * p.n_elems;
    const double *xp = p.Xprev ? p.Xprev + lane * p.x_lane_stride : nullptr;
    T acc = S::zero();
    const int t0 = p.term_ptr[k], t1 = p.term_ptr[k + 1];
    for (int t = t0; t < t1; t++) {
        const FrontTerm tm = p.terms[t];
        const FrontElem el = p.elems[tm.elem];
        const double v = ev[tm.elem];
        double c = 0.0;
        switch (tm.kind & 0xff) {
        case S_VAL: c = v; break;                                                    // stampable.rs:567, 1017
        case S_ACAMP: c = el.ac_amp; break;                                          // :1031
        case S_WAVE: c = p.Bstep[((size_t)p.row * 2) * p.n_wave + el.wave_col]; break;   // :1043-1048 (table built on the host)
        case S_CAP_HIST: {                                                           // (C/h) * (v+_prev - v-_prev)   :480-493
            const double coef = __ddiv_rn(v, p.h_comp);
            double dv;
            if (el.ip >= 0 && el.im >= 0) dv = __dsub_rn(xp[el.ip], xp[el.im]);
            else if (el.ip >= 0) dv = xp[el.ip];
            else dv = -xp[el.im];
            c = __dmul_rn(coef, dv);
            break;
        }
        case S_IND_HIST: c = __dmul_rn(__ddiv_rn(v, p.h_comp), xp[el.ic]); break;  // (L/h) * i_prev   :806-812
        case S_CAP_IPREV: c = p.ST[lane * p.n_state + el.state_idx]; break;
        case S_IND_VPREV:
            if (el.ip >= 0 && el.im >= 0) c = __dsub_rn(xp[el.ip], xp[el.im]);
            else if (el.ip >= 0) c = xp[el.ip];
            else if (el.im >= 0) c = -xp[el.im];
            break;
        default: break;
        }
        if (tm.kind & 0x100) c = -c;
        acc = (t == t0) ? S::from_real(c) : S::add(acc, S::from_real(c));
    }
    reinterpret_cast<T *>(p.OUT)[lane * p.count + k] = acc;
}

// i_n = (2C/h) * ((v+ - v-)_n - (v+ - v-)_{n-1}) - i_{n-1}
struct FrontTrapParams {
    const FrontElem *elems;
    const int *state_elems;
    int n_state, n_elems;
    const double *EV;
    double h_comp;
    const double *Xnew, *Xprev;
    long long x_lane_stride;
    double *ST;
};
__global__ void krb_front_trap_state(const FrontTrapParams p) {
    const long long lane = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.n_state) return;
    const int e = p.state_elems[j];
    const FrontElem el = p.elems[e];
    const double g = __ddiv_rn(p.EV[lane * p.n_elems + e], p.h_comp);
    const double *xn = p.Xnew + lane * p.x_lane_stride, *xo = p.Xprev + lane * p.x_lane_stride;
    double vn, vp;
    if (el.ip >= 0 && el.im >= 0) { vn = __dsub_rn(xn[el.ip], xn[el.im]); vp = __dsub_rn(xo[el.ip], xo[el.im]); }
    else if (el.ip >= 0) { vn = xn[el.ip]; vp = xo[el.ip]; }
    else if (el.im >= 0) { vn = -xn[el.im]; vp = -xo[el.im]; }
    else return;
    double *st = p.ST + lane * p.n_state + j;
    *st = __dsub_rn(__dmul_rn(g, __dsub_rn(vn, vp)), *st);
}

// Threshold partial pivoting inside the fully-summed block of a front (rows k..p-1 of column k): the static pivot is kept
// unless another fully-summed row is more than 8x larger (or the static pivot is zero / non-finite).  Per lane, like the
// reference's faer LU picks pivots per factorisation; rows outside the front are never touched, so the symbolic plan holds.
template <bool CPLX> struct FrontAbs;
template <> struct FrontAbs<false> {
    static __device__ __forceinline__ double mag(double v) { return fabs(v); }
    static __device__ __forceinline__ double abs1(double v) { return fabs(v); }
    static constexpr double THRESH = 8.0;
    static constexpr uint32_t LOG2T = 3;
};
template <> struct FrontAbs<true> {
    static __device__ __forceinline__ double mag(double2 v) { return fma(v.x, v.x, v.y * v.y); }
    static __device__ __forceinline__ double abs1(double2 v) { return fabs(v.x) + fabs(v.y); }
    static constexpr double THRESH = 64.0;
    static constexpr uint32_t LOG2T = 6;
};

__device__ __forceinline__ double front_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double2 front_shfl(double2 v, int src) { return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)); }

// warp 0 scans col[k+1 .. p) (unit stride) and publishes the chosen row in *s_ip
template <bool CPLX, typename T>
__device__ __forceinline__ void front_pivot_search(const T *col, int k, int p, int tid, int *s_ip) {
    if (tid < 32) {
        double best = -1.0;
        int bi = k;
        for (int i = k + 1 + tid; i < p; i += 32) {
            const double a = FrontAbs<CPLX>::mag(col[i]);
            if (a > best) { best = a; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (tid == 0) {
            const double d = FrontAbs<CPLX>::mag(col[k]);
            const bool keep = isfinite(d) && d > 0.0 && !(best > FrontAbs<CPLX>::THRESH * d);
            *s_ip = keep ? k : bi;
        }
    }
}

// ---- (b) numeric factorisation of one front ---------------------------------------------------------------------
// ---- FP64 tensor-core (DMMA) trailing update ---------------------------------------------------------------------------
// The rank-nb update C -= L U of a large front is the one genuinely dense contraction on this path.  With FMA register tiles it
// is bound by operand delivery (8 shared-memory loads per 16 FMAs per thread); mma.sync.m8n8k4.f64 needs one A and one B value
// per thread for 8 FMAs' worth of work and a quarter of the issue slots.  Fragment layout (PTX ISA, m8n8k4 .f64): g = lane >> 2,
// t = lane & 3; A[g][t], B[t][g], C[g][2t + {0, 1}].  16 warps as 4 x 4, each warp owns a 32 x 16 block of C (4 x 2 MMA tiles).
__device__ __forceinline__ void front_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// C(i, j) -= sum_{k < K} Lt(i, k) Up(k, j), i < rows, j < cols; blockDim must be (32, 16)
__device__ __forceinline__ void front_dmma_update(double *C, int ldc, int rows, int cols, const double *Lt, int ldl, const double *Up, int ldu, int K,
                                                  int lane, int warp) {
    const int g = lane >> 2, t = lane & 3, wr = warp & 3, wc = warp >> 2;
    for (int j0 = 0; j0 < cols; j0 += 64)
        for (int i0 = 0; i0 < rows; i0 += 128) {
            const int rbase = i0 + wr * 32, cbase = j0 + wc * 16;
            if (rbase >= rows || cbase >= cols) continue;                 // whole warp leaves together
            double c[4][2][2];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const int r = rbase + a * 8 + g, cc = cbase + b * 8 + 2 * t + i;
                        c[a][b][i] = (r < rows && cc < cols) ? C[r + (size_t)ldc * cc] : 0.0;
                    }
            int ra[4], cb[2];                                             // clamped fragment coordinates (results of padding discarded)
#pragma unroll
            for (int a = 0; a < 4; a++) { const int r = rbase + a * 8 + g; ra[a] = r < rows ? r : rows - 1; }
#pragma unroll
            for (int b = 0; b < 2; b++) { const int cc = cbase + b * 8 + g; cb[b] = cc < cols ? cc : cols - 1; }
            for (int k0 = 0; k0 < K; k0 += 4) {
                const int kk = k0 + t;
                const bool kv = kk < K;
                const int kc = kv ? kk : K - 1;
                double af[4], bf[2];
#pragma unroll
                for (int a = 0; a < 4; a++) { const double v = Lt[ra[a] + (size_t)ldl * kc]; af[a] = kv ? -v : 0.0; }
#pragma unroll
                for (int b = 0; b < 2; b++) { const double v = Up[kc + (size_t)ldu * cb[b]]; bf[b] = kv ? v : 0.0; }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++) front_dmma(c[a][b][0], c[a][b][1], af[a], bf[b]);
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const int r = rbase + a * 8 + g, cc = cbase + b * 8 + 2 * t + i;
                        if (r < rows && cc < cols) C[r + (size_t)ldc * cc] = c[a][b][i];
                    }
        }
}

// blockDim = (TX, TY); TX consecutive threads walk a column (unit stride), TY columns at a time.
//
// One algorithm for every front size (right-looking, blocked, LAPACK getrf order):
//   for each panel of nb pivot columns:  factor the panel (threshold pivoting inside the fully-summed rows, rank-1 updates
//   inside the panel only)  ->  interchanges + U12 = L11^-1 A12 on the columns right of it (one thread per column)  ->
//   rank-nb update of the trailing block with 4 x 4 register tiles (each trailing element is read and written once per panel).
// Fronts that fit are staged in shared memory whole (odd leading dimension: conflict-free column-per-thread access); larger
// fronts live in the per-lane workspace (L2/HBM) and only the panel's L columns and U rows are copied to shared memory.
template <bool CPLX, int R>   // R x R register tiles in the trailing update: 4 for the large classes, 2 (more resident CTAs) for small fronts
__global__ void __launch_bounds__(R == 4 ? 512 : 128, R == 4 ? 1 : 6) krb_front_factor(const FrontParams P) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    extern __shared__ __align__(16) unsigned char front_smem[];
    T *sm = reinterpret_cast<T *>(front_smem);
    const int f = P.list[blockIdx.x];
    const long long lane = blockIdx.y;
    const FrontDesc d = P.fronts[f];
    const int p = d.p, q = d.q, m = p + q;
    const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
    const int tid = ty * TX + tx, NT = TX * TY;
    const bool in_smem = d.fw_off < 0;
    T *F = in_smem ? sm : reinterpret_cast<T *>(P.FW) + lane * P.fw_stride + d.fw_off;
    const int ldf = in_smem ? (m | 1) : m;          // F(i, j) = F[i + ldf * j]
    bool bad = false;
    uint32_t grow = 0;   // largest multiplier |l_ik| = |a_ik / pivot| seen by this thread (scalar.cuh: growth_key)
    const bool use_dmma = (TX == 32 && TY == 16 && P.use_dmma);
    __shared__ int s_ip;
    __shared__ int s_pv[32];
    int *ipiv = P.IPIV + lane * P.n + d.piv0;

    // ---- assemble: zero, stamped entries, children's contribution blocks (extend-add, children in ascending order) ----
    for (int j = ty; j < m; j += TY)
        for (int i = tx; i < m; i += TX) F[i + (size_t)ldf * j] = S::zero();
    __syncthreads();
    {
        const T *A = reinterpret_cast<const T *>(P.A) + lane * P.a_stride;
        for (int t = d.asm_begin + tid; t < d.asm_end; t += NT) {
            const int pos = P.asm_pos[t], c = pos / m, r = pos - c * m;
            F[r + (size_t)ldf * c] = A[P.asm_ent[t]];
        }
    }
    __syncthreads();
    for (int ci = d.child_begin; ci < d.child_end; ci++) {
        const FrontChild c = P.childrec[ci];
        const int qc = c.q;
        const T *CBc = reinterpret_cast<const T *>(P.CB) + lane * P.cb_stride + c.cb_off;
        const int *rel = P.rel + c.idx_off;
        for (int j = ty; j < qc; j += TY) {
            const size_t rj = (size_t)rel[j] * ldf;
            for (int i = tx; i < qc; i += TX) { T *dst = F + rel[i] + rj; *dst = S::add(*dst, CBc[i + (size_t)qc * j]); }
        }
        __syncthreads();
    }

    // ---- blocked elimination ----
    int NB;
    if (in_smem) NB = CPLX ? 4 : 8;
    else { NB = (P.smem_elems - m) / (2 * m); NB = NB > 32 ? 32 : NB; }   // wide panels: the trailing block streams through L2/HBM once per panel
    for (int kb = 0; kb < p; kb += NB) {
        const int nb = (p - kb) < NB ? (p - kb) : NB;
        const int mk = m - kb, tw = mk - nb;
        // panel views: Lp(i, k) = Lp[i + ldl * k] (mk x nb), Up(k, j) = Up[k + ldu * j] (nb x tw)
        T *Lp, *Up;
        int ldl, ldu;
        if (in_smem) { Lp = F + kb + (size_t)ldf * kb; ldl = ldf; Up = F + kb + (size_t)ldf * (kb + nb); ldu = ldf; }
        else {
            Lp = sm; ldl = mk; Up = sm + (size_t)mk * nb; ldu = nb | 1;
            for (int k = ty; k < nb; k += TY)
                for (int i = tx; i < mk; i += TX) Lp[i + ldl * k] = F[(kb + i) + (size_t)m * (kb + k)];
            for (int j = tid; j < tw; j += NT)
                for (int k = 0; k < nb; k++) Up[k + ldu * j] = F[(kb + k) + (size_t)m * (kb + nb + j)];
            __syncthreads();
        }
        // panel factorisation: interchanges are applied to the panel columns and to the finished L columns on the left at
        // once; the columns right of the panel get them below, before their U rows are formed
        for (int k = 0; k < nb; k++) {
            int ip = k;
            if (kb + k + 1 < p) {
                front_pivot_search<CPLX>(Lp + (size_t)ldl * k, k, p - kb, tid, &s_ip);
                __syncthreads();
                ip = s_ip;
                if (ip != k) {
                    for (int j = tid; j < nb; j += NT) { const T t = Lp[k + ldl * j]; Lp[k + ldl * j] = Lp[ip + ldl * j]; Lp[ip + ldl * j] = t; }
                    for (int j = tid; j < kb; j += NT) {
                        T *a = F + (kb + k) + (size_t)ldf * j, *b = F + (kb + ip) + (size_t)ldf * j;
                        const T t = *a; *a = *b; *b = t;
                    }
                }
            }
            if (tid == 0) { ipiv[kb + k] = ip + kb; s_pv[k] = ip; }
            __syncthreads();
            bool b1;
            const T rp = S::rcp(Lp[k + ldl * k], b1);
            bad |= b1;
            for (int i = k + 1 + tid; i < mk; i += NT) { const T l = S::mul(Lp[i + ldl * k], rp); Lp[i + ldl * k] = l; grow = max(grow, growth_key(l)); }
            __syncthreads();
            for (int j = k + 1 + ty; j < nb; j += TY) {
                const T u = Lp[k + ldl * j];
                for (int i = k + 1 + tx; i < mk; i += TX) Lp[i + ldl * j] = S::fnma(Lp[i + ldl * j], Lp[i + ldl * k], u);
            }
            __syncthreads();
        }
        // columns right of the panel, one thread per column: interchanges in order, then U12 = L11^-1 A12 (unit lower)
        for (int j = tid; j < tw; j += NT) {
            T *uc = Up + (size_t)ldu * j;
            for (int k = 0; k < nb; k++) {
                const int ip = s_pv[k];
                if (ip != k) {
                    T *other = (in_smem || ip < nb) ? uc + ip : F + (kb + ip) + (size_t)m * (kb + nb + j);
                    const T t = uc[k]; uc[k] = *other; *other = t;
                }
            }
            for (int k = 0; k < nb; k++) {
                const T u = uc[k];
                for (int r = k + 1; r < nb; r++) uc[r] = S::fnma(uc[r], Lp[r + ldl * k], u);
            }
        }
        if (!in_smem) {   // panels back to the front (final L and U values)
            for (int k = ty; k < nb; k += TY)
                for (int i = tx; i < mk; i += TX) F[(kb + i) + (size_t)m * (kb + k)] = Lp[i + ldl * k];
            for (int j = tid; j < tw; j += NT)
                for (int k = 0; k < nb; k++) F[(kb + k) + (size_t)m * (kb + nb + j)] = Up[k + ldu * j];
        }
        __syncthreads();
        // trailing block C(i, j) -= sum_k Lp(nb + i, k) Up(k, j): FP64 tensor cores for the large real classes, else R x R register
        // tiles with C loaded before the k loop
        if (!CPLX && R == 4 && use_dmma) {
            front_dmma_update(reinterpret_cast<double *>(F) + (kb + nb) + (size_t)ldf * (kb + nb), ldf, tw, tw, reinterpret_cast<const double *>(Lp) + nb, ldl,
                              reinterpret_cast<const double *>(Up), ldu, nb, tx, ty);
        } else {
            T *C = F + (kb + nb) + (size_t)ldf * (kb + nb);
            const T *Lt = Lp + nb;
            for (int j0 = 0; j0 < tw; j0 += R * TY)
                for (int i0 = 0; i0 < tw; i0 += R * TX) {
                    int ii[R], jj[R];
#pragma unroll
                    for (int a = 0; a < R; a++) { ii[a] = i0 + tx + a * TX; jj[a] = j0 + ty + a * TY; }
                    if (ii[0] >= tw || jj[0] >= tw) continue;
                    T c[R][R];
#pragma unroll
                    for (int b = 0; b < R; b++)
#pragma unroll
                        for (int a = 0; a < R; a++) c[a][b] = (ii[a] < tw && jj[b] < tw) ? C[ii[a] + (size_t)ldf * jj[b]] : S::zero();
                    // clamp the indices of partial tiles (their results are discarded) so the inner loop is branch-free
                    int il[R], jl[R];
#pragma unroll
                    for (int a = 0; a < R; a++) { il[a] = ii[a] < tw ? ii[a] : ii[0]; jl[a] = jj[a] < tw ? jj[a] : jj[0]; }
                    for (int k = 0; k < nb; k++) {
                        T l[R], u[R];
#pragma unroll
                        for (int a = 0; a < R; a++) { l[a] = Lt[il[a] + (size_t)ldl * k]; u[a] = Up[k + (size_t)ldu * jl[a]]; }
#pragma unroll
                        for (int b = 0; b < R; b++)
#pragma unroll
                            for (int a = 0; a < R; a++) c[a][b] = S::fnma(c[a][b], l[a], u[b]);
                    }
#pragma unroll
                    for (int b = 0; b < R; b++)
#pragma unroll
                        for (int a = 0; a < R; a++)
                            if (ii[a] < tw && jj[b] < tw) C[ii[a] + (size_t)ldf * jj[b]] = c[a][b];
                }
        }
        __syncthreads();
    }

    // ---- write out: L block (m x p, reciprocal pivots on the diagonal), U12 (p x q), contribution block (q x q) ----
    T *LUf = reinterpret_cast<T *>(P.LU) + lane * P.lu_stride + d.lu_off;
    for (int k = ty; k < p; k += TY)
        for (int i = tx; i < m; i += TX) {
            T v = F[i + (size_t)ldf * k];
            if (i == k) { bool b1; v = S::rcp(v, b1); }
            LUf[i + (size_t)m * k] = v;
        }
    T *U12 = LUf + (size_t)m * p;
    for (int j = ty; j < q; j += TY)
        for (int k = tx; k < p; k += TX) U12[k + (size_t)p * j] = F[k + (size_t)ldf * (p + j)];
    T *CBf = reinterpret_cast<T *>(P.CB) + lane * P.cb_stride + d.cb_off;
    for (int j = ty; j < q; j += TY)
        for (int i = tx; i < q; i += TX) CBf[i + (size_t)q * j] = F[(p + i) + (size_t)ldf * (p + j)];
    if (bad && tid == 0 && P.status) atomicOr(P.status + lane, KRB_LANE_ZERO_PIVOT);
    // pivoting is restricted to the front's fully-summed rows: a multiplier of 2^24 or more means no acceptable pivot was among them
    if (grow >= KRB_GROWTH_KEY && grow < KRB_NONFINITE_KEY && P.status) atomicOr(P.status + lane, KRB_LANE_INACCURATE);
}

// ---- (b') left-looking factorisation of a LARGE front: one column block at a time resident in shared memory --------------------
// Fronts that do not fit shared memory used to live in a per-lane global workspace whose trailing block was re-read and re-written
// once per 15-column panel (a 451 x 451 complex front moved 27x its size through DRAM: ncu r1, 26 % of the DRAM peak with the FP64
// pipe 25 % active).  Here the front is processed by column blocks of W columns (m x W resident in shared memory, W = 51 real /
// 29 complex for m = 451):
//   for each block J:   assemble its columns straight into shared memory (stamped entries + the children's contribution-block
//                       columns that land in it: no global workspace at all)
//                       apply the interchanges of all earlier pivots
//                       for each earlier pivot panel K:  U_KJ = L_KK^-1 C_K  (8 lanes per column, warp-synchronous)
//                                                        C_below -= L_below,K U_KJ  on the FP64 tensor cores (DMMA m8n8k4; complex =
//                                                        4 real DMMAs), A operands = finished L columns streamed from the factor store
//                       factorise the block's own pivot columns in sub-panels of 8: every thread owns ONE ROW of the sub-panel in
//                       registers, ONE barrier per pivot column (warp arg-max + published candidate rows, threshold pivoting as
//                       in krb_front_factor), then the sub-panel's U rows and a DMMA update of the rest of the block
//                       write the block to the factor store (L | U11, U12) and the contribution block
// Each front entry is written once and every finished L panel is read once per later block: ~(2 + p/W ... ) x the front size instead
// of 27x, and the rank-k updates run out of shared memory + registers.  blockDim = (32, NW), m <= 32 NW.  The same kernel serves the
// mid-size fronts (32 < m <= 128 on 4 warps with 56 KB: four CTAs per SM; m <= 256 on 8 warps with 112 KB): DMMA updates and one
// barrier per pivot instead of the register-tile kernel's three, at four times its residency.
#ifdef KRB_LL_PROFILE
// per-phase cycles of thread 0, summed over all CTAs (profiling builds only: tools/build_variant.sh ... -DKRB_LL_PROFILE)
__device__ unsigned long long krb_ll_prof[24];   // [warps 16 | 8 | 4][phase]
#define LLP(i) do { if (tid == 0) { const long long t1_ = clock64(); llp[i] += t1_ - t0_; t0_ = t1_; } } while (0)
#else
#define LLP(i) do { } while (0)
#endif
constexpr int FRONT_LL_AUX = 384;    // elements of T reserved at the start of the dynamic shared memory (published rows, arg-max slots)
constexpr int FRONT_LL_SPW = 8;      // sub-panel width

__device__ __forceinline__ int front_lower_bound(const int *__restrict__ a, int lo, int hi, int key) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid; }
    return lo;
}
// column-block width for a front of m rows (ldc = padded leading dimension) in `elems` elements of shared memory: W columns of the
// front + a W x W staging block for L_KK; at most 64 (one pass of the 4 x 4 warp grid), then balanced over the blocks
__host__ __device__ inline int front_ll_width(int m, int ldc, long long elems) {
    if ((long long)m * ldc <= elems) return m;           // the whole front is one block (no L_KK staging): mid-size fronts
    int W = 1;
    while (W < 64 && (long long)(W + 1) * ldc + (long long)(W + 1) * (W + 1) <= elems) W++;
    const int nblk = (m + W - 1) / W;
    return (m + nblk - 1) / nblk;
}
__host__ __device__ inline int front_ll_ldc(int m) { return m + ((4 - m) & 7); }   // == 4 (mod 8): conflict-free DMMA B-fragment reads

// C(i, j) -= sum_{k < K} Lt(i, k) Up(k, j) on the FP64 tensor cores, operands prefetched one k-step ahead.  NW warps as WR x WC, each warp a
// 32 x 16 block of C (4 x 2 MMA tiles); real: 8 DMMAs per k-step of 4, complex: 32 (re/im parts as four real products).
template <bool CPLX, int NW> struct FrontMma;
template <int NW> struct FrontMma<false, NW> {
    static __device__ __forceinline__ void update(double *C, int ldc, int rows, int cols, const double *Lt, int ldl, const double *Up, int ldu, int K, int lane, int warp) {
        // warp grid WR x WC by the number of columns: narrow blocks (complex: W = 29) must not leave half of the warps without a tile
        const int WC = cols <= 16 ? 1 : cols <= 32 ? 2 : 4, WR = NW / WC;
        const int g = lane >> 2, t = lane & 3, wr = warp % WR, wc = warp / WR;
        for (int j0 = 0; j0 < cols; j0 += 16 * WC)
            for (int i0 = 0; i0 < rows; i0 += 32 * WR) {
                const int rbase = i0 + wr * 32, cbase = j0 + wc * 16;
                if (rbase >= rows || cbase >= cols) continue;
                double c[4][2][2];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++)
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const int r = rbase + a * 8 + g, cc = cbase + b * 8 + 2 * t + i;
                            c[a][b][i] = (r < rows && cc < cols) ? C[r + (size_t)ldc * cc] : 0.0;
                        }
                const double *ap[4], *bp[2];
#pragma unroll
                for (int a = 0; a < 4; a++) { const int r = rbase + a * 8 + g; ap[a] = Lt + (r < rows ? r : rows - 1); }
#pragma unroll
                for (int b = 0; b < 2; b++) { const int cc = cbase + b * 8 + g; bp[b] = Up + (size_t)ldu * (cc < cols ? cc : cols - 1); }
                double af[4], bf[2];
                {
                    const bool kv = t < K;
                    const int kc = kv ? t : K - 1;
#pragma unroll
                    for (int a = 0; a < 4; a++) { const double v = ap[a][(size_t)ldl * kc]; af[a] = kv ? -v : 0.0; }
#pragma unroll
                    for (int b = 0; b < 2; b++) { const double v = bp[b][kc]; bf[b] = kv ? v : 0.0; }
                }
                for (int k0 = 0; k0 < K; k0 += 4) {
                    double an[4], bn[2];
                    {
                        const int kk = k0 + 4 + t;
                        const bool kv = kk < K;
                        const int kc = kv ? kk : K - 1;
#pragma unroll
                        for (int a = 0; a < 4; a++) { const double v = ap[a][(size_t)ldl * kc]; an[a] = kv ? -v : 0.0; }
#pragma unroll
                        for (int b = 0; b < 2; b++) { const double v = bp[b][kc]; bn[b] = kv ? v : 0.0; }
                    }
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int b = 0; b < 2; b++) front_dmma(c[a][b][0], c[a][b][1], af[a], bf[b]);
#pragma unroll
                    for (int a = 0; a < 4; a++) af[a] = an[a];
#pragma unroll
                    for (int b = 0; b < 2; b++) bf[b] = bn[b];
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++)
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const int r = rbase + a * 8 + g, cc = cbase + b * 8 + 2 * t + i;
                            if (r < rows && cc < cols) C[r + (size_t)ldc * cc] = c[a][b][i];
                        }
            }
    }
};
template <int NW> struct FrontMma<true, NW> {
    static __device__ __forceinline__ void update(double2 *C, int ldc, int rows, int cols, const double2 *Lt, int ldl, const double2 *Up, int ldu, int K, int lane, int warp) {
        const int WC = cols <= 16 ? 1 : cols <= 32 ? 2 : 4, WR = NW / WC;
        const int g = lane >> 2, t = lane & 3, wr = warp % WR, wc = warp / WR;
        for (int j0 = 0; j0 < cols; j0 += 16 * WC)
            for (int i0 = 0; i0 < rows; i0 += 32 * WR) {
                const int rbase = i0 + wr * 32, cbase = j0 + wc * 16;
                if (rbase >= rows || cbase >= cols) continue;
                double cr[4][2][2], ci[4][2][2];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++)
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const int r = rbase + a * 8 + g, cc = cbase + b * 8 + 2 * t + i;
                            const double2 v = (r < rows && cc < cols) ? C[r + (size_t)ldc * cc] : make_double2(0.0, 0.0);
                            cr[a][b][i] = v.x; ci[a][b][i] = v.y;
                        }
                const double2 *ap[4], *bp[2];
#pragma unroll
                for (int a = 0; a < 4; a++) { const int r = rbase + a * 8 + g; ap[a] = Lt + (r < rows ? r : rows - 1); }
#pragma unroll
                for (int b = 0; b < 2; b++) { const int cc = cbase + b * 8 + g; bp[b] = Up + (size_t)ldu * (cc < cols ? cc : cols - 1); }
                // A fragments (finished L columns: global memory / L2) are fetched one k-step ahead, B fragments (shared memory) at use:
                // prefetching both costs 24 more registers and spills inside this loop (ncu r2o: 9 % of the samples on spill reloads)
                double2 af[4], bf[2];
                {
                    const bool kv = t < K;
                    const int kc = kv ? t : K - 1;
#pragma unroll
                    for (int a = 0; a < 4; a++) { const double2 v = ap[a][(size_t)ldl * kc]; af[a] = kv ? v : make_double2(0.0, 0.0); }
                }
                for (int k0 = 0; k0 < K; k0 += 4) {
                    double2 an[4];
                    {
                        const int kk = k0 + 4 + t;
                        const bool kv = kk < K;
                        const int kc = kv ? kk : K - 1;
#pragma unroll
                        for (int a = 0; a < 4; a++) { const double2 v = ap[a][(size_t)ldl * kc]; an[a] = kv ? v : make_double2(0.0, 0.0); }
                    }
                    {
                        const int kk = k0 + t;
                        const bool kv = kk < K;
                        const int kc = kv ? kk : K - 1;
#pragma unroll
                        for (int b = 0; b < 2; b++) { const double2 v = bp[b][kc]; bf[b] = kv ? v : make_double2(0.0, 0.0); }
                    }
                    // C -= A B:  C_re += (-A_re) B_re + A_im B_im,  C_im += (-A_re) B_im + (-A_im) B_re
#pragma unroll
                    for (int a = 0; a < 4; a++) {
                        const double nar = -af[a].x, ai = af[a].y, nai = -af[a].y;
#pragma unroll
                        for (int b = 0; b < 2; b++) {
                            front_dmma(cr[a][b][0], cr[a][b][1], nar, bf[b].x);
                            front_dmma(cr[a][b][0], cr[a][b][1], ai, bf[b].y);
                            front_dmma(ci[a][b][0], ci[a][b][1], nar, bf[b].y);
                            front_dmma(ci[a][b][0], ci[a][b][1], nai, bf[b].x);
                        }
                    }
#pragma unroll
                    for (int a = 0; a < 4; a++) af[a] = an[a];
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++)
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const int r = rbase + a * 8 + g, cc = cbase + b * 8 + 2 * t + i;
                            if (r < rows && cc < cols) C[r + (size_t)ldc * cc] = make_double2(cr[a][b][i], ci[a][b][i]);
                        }
            }
    }
};

// NW warps per CTA: 16 (m <= 512, one CTA per SM), 8 (m <= 256, two), 4 (m <= 128, four).  RPT = 2 (real plans, 16 warps): a thread owns a
// SECOND row, 512 below its first, in the sub-panel factorisation — fronts up to 1024 rows whose pivot block has at most 512 (the second
// rows are never pivot candidates: they only receive the updates).  DC plans of inductor-heavy meshes have such fronts (config 5: 694 rows).
template <bool CPLX, int NW, int RPT = 1>
__global__ void __launch_bounds__(NW * 32, 16 / NW) krb_front_factor_ll(const FrontParams P) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    constexpr int SPW = FRONT_LL_SPW, NT = NW * 32;
    extern __shared__ __align__(16) unsigned char front_smem[];
    T *aux = reinterpret_cast<T *>(front_smem);
    // published rows of the sub-panel factorisation, double-buffered by the parity of the pivot column
    T *slot_rows = aux;                                  // [2][NW][SPW] a warp's best candidate row
    T *rowk = aux + 2 * NW * SPW;                        // [2][SPW]     the row at the static pivot position
    T *rp_slot = rowk + 2 * SPW;                         // [2]          reciprocal of the entry at the static pivot position
    double *slot_mag = reinterpret_cast<double *>(rp_slot + 2);                    // [2][NW]
    int *slot_idx = reinterpret_cast<int *>(slot_mag + 2 * NW);                    // [2][NW]
    int *s_nswap = slot_idx + 2 * NW;
    T *Cs = aux + FRONT_LL_AUX;
    const int f = P.list[blockIdx.x];
    const long long lane = blockIdx.y;
    const FrontDesc d = P.fronts[f];
    const int p = d.p, q = d.q, m = p + q;
    const int lx = threadIdx.x, wy = threadIdx.y, tid = wy * 32 + lx;
    const int ldc = front_ll_ldc(m);
    const int W = d.pad;                                 // front_ll_width(m, ldc, shared memory of this launch), computed on the host
    T *Ls = Cs + (size_t)ldc * W;                        // [W][W] strictly lower part of L_KK
    bool bad = false;
    uint32_t grow = 0;
    int *ipiv = P.IPIV + lane * P.n + d.piv0;
    T *LUf = reinterpret_cast<T *>(P.LU) + lane * P.lu_stride + d.lu_off;
    T *U12 = LUf + (size_t)m * p;
    T *CBf = reinterpret_cast<T *>(P.CB) + lane * P.cb_stride + d.cb_off;
    const T *A = reinterpret_cast<const T *>(P.A) + lane * P.a_stride;
    if (tid == 0) *s_nswap = 0;
    if (NW == 4) {
        // mid-size fronts have ~4 small children each: their contribution blocks and index lists are requested from HBM now, so that
        // the per-child chains of dependent loads below (record -> relative indices -> block) find them in L2
        for (int ci = d.child_begin + wy; ci < d.child_end; ci += NW) {
            const FrontChild c = P.childrec[ci];
            const char *cb = reinterpret_cast<const char *>(reinterpret_cast<const T *>(P.CB) + lane * P.cb_stride + c.cb_off);
            const size_t bytes = (size_t)c.q * c.q * sizeof(T);
            for (size_t o = (size_t)lx * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(cb + o));
            if (lx * 32 < c.q) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.rel + c.idx_off + lx * 32));
        }
    }
#ifdef KRB_LL_PROFILE
    long long llp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0_ = clock64();
#endif

    for (int c0 = 0; c0 < m; c0 += W) {
        const int wj = (m - c0) < W ? (m - c0) : W;
        // ---- (a) assemble the block: zero, stamped entries (sorted by position), children's contribution-block columns ----
        for (int e = tid; e < ldc * wj; e += NT) Cs[e] = S::zero();
        __syncthreads();
        // (a front's list of stamped entries is short: every block scans it instead of searching its range, a chain of dependent loads)
        for (int t = d.asm_begin + tid; t < d.asm_end; t += NT) {
            const int pos = __ldg(P.asm_pos + t), c = pos / m, r = pos - c * m;
            if (c >= c0 && c < c0 + wj) Cs[r + (size_t)ldc * (c - c0)] = A[__ldg(P.asm_ent + t)];
        }
        __syncthreads();
        // children: warp w owns the block's columns w, w + NW, ... — two children that add into the same entry are handled by the same warp
        // in program order, so the children need no block-wide barrier between them (ncu/cycle profile r2s: the assembly was a third of
        // the mid-size fronts' time, a chain of dependent loads and barriers per child)
        for (int ci = d.child_begin; ci < d.child_end; ci++) {
            const FrontChild c = P.childrec[ci];
            const int qc = c.q;
            const int *rel = P.rel + c.idx_off;
            const T *CBc = reinterpret_cast<const T *>(P.CB) + lane * P.cb_stride + c.cb_off;
            for (int jb = 0; jb < qc; jb += 32) {
                const int dcol = (jb + lx < qc) ? __ldg(rel + jb + lx) - c0 : -1;
                unsigned mask = __ballot_sync(0xffffffffu, dcol >= 0 && dcol < wj && (dcol & (NW - 1)) == wy);
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    T *dst = Cs + (size_t)ldc * __shfl_sync(0xffffffffu, dcol, b);
                    const T *src = CBc + (size_t)qc * (jb + b);
                    for (int ib = lx; ib < qc; ib += 128) {
                        int ri[4];
                        T v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const bool in = ib + 32 * u < qc;
                            ri[u] = in ? __ldg(rel + ib + 32 * u) : -1;
                            v[u] = in ? src[ib + 32 * u] : S::zero();
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) if (ri[u] >= 0) { T *e = dst + ri[u]; *e = S::add(*e, v[u]); }
                    }
                }
            }
            __syncwarp();
        }
        __syncthreads();
        // ---- (b) interchanges of all earlier pivots (rare) ----
        const int kdone = c0 < p ? c0 : p;
        if (*s_nswap > 0 && kdone > 0) {
            for (int j = tid; j < wj; j += NT) {
                T *cj = Cs + (size_t)ldc * j;
                for (int k = 0; k < kdone; k++) { const int ip = ipiv[k]; if (ip != k) { const T t = cj[k]; cj[k] = cj[ip]; cj[ip] = t; } }
            }
            __syncthreads();
        }
        LLP(0);
        // ---- (c) left-looking: the earlier pivot panels ----
        // earlier panels are taken in PAIRS: after the first panel's U rows only the second panel's own rows are updated (a small
        // product), then both panels update everything below them in ONE pass with K = 2W — half as many accumulator round trips through
        // shared memory and exposed first-fragment loads per DMMA flop (cycle profile r2s: the DMMA phases ran at ~35 % of the tensor peak
        // with 8-10 k-steps per accumulator tile)
        for (int k0 = 0; k0 < kdone;) {
            const int wk0 = (p - k0) < W ? (p - k0) : W;
            const int k1 = k0 + wk0;
            const bool pair = k1 < kdone;
            const int wk1 = pair ? ((p - k1) < W ? (p - k1) : W) : 0;
            for (int h = 0; h < (pair ? 2 : 1); h++) {
                const int kb = h ? k1 : k0, wb = h ? wk1 : wk0;
                if (h) {
                    FrontMma<CPLX, NW>::update(Cs + k1, ldc, wk1, wj, LUf + k1 + (size_t)m * k0, m, Cs + k0, ldc, wk0, lx, wy);
                    __syncthreads();
                    LLP(3);
                }
                for (int e = tid; e < wb * wb; e += NT) {
                    const int k = e / wb, r = e - k * wb;
                    Ls[e] = r > k ? LUf[(kb + r) + (size_t)m * (kb + k)] : S::zero();
                }
                __syncthreads();
                LLP(1);
            {   // U_KJ = L_KK^-1 C_K: 8 lanes per column; the 8 x 8 diagonal steps redundantly, the rows below them split over the lanes
                    const int r8 = tid & 7;
                    for (int cb = 0; cb < wj; cb += NT / 8) {
                    const int col = cb + (tid >> 3);
                    const bool act = col < wj;
                    T *cj = Cs + kb + (size_t)ldc * (act ? col : 0);
                    for (int s0 = 0; s0 < wb; s0 += 8) {
                        const int ns = (wb - s0) < 8 ? (wb - s0) : 8;
                        T v[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) v[k] = k < ns ? cj[s0 + k] : S::zero();
#pragma unroll
                        for (int k = 0; k < 8; k++)
#pragma unroll
                            for (int r = k + 1; r < 8; r++)
                                if (r < ns) v[r] = S::fnma(v[r], Ls[(s0 + r) + wb * (s0 + k)], v[k]);
                        __syncwarp();                 // every lane of the group has read the 8 rows
                        if (act && r8 == 0)
#pragma unroll
                            for (int k = 0; k < 8; k++) if (k < ns) cj[s0 + k] = v[k];
                        for (int r = s0 + 8 + r8; r < wb; r += 8) {
                            T acc = cj[r];
#pragma unroll
                            for (int k = 0; k < 8; k++) if (k < ns) acc = S::fnma(acc, Ls[r + wb * (s0 + k)], v[k]);
                            if (act) cj[r] = acc;
                        }
                        __syncwarp();
                    }
                    }
                }
                __syncthreads();
                LLP(2);
            }
            const int kend = k1 + wk1;
            if (m - kend > 0)
                FrontMma<CPLX, NW>::update(Cs + kend, ldc, m - kend, wj, LUf + kend + (size_t)m * k0, m, Cs + k0, ldc, wk0 + wk1, lx, wy);
            __syncthreads();
            LLP(3);
            k0 = kend;
        }
        // ---- (d) the block's own pivot columns, sub-panels of SPW columns, one row per thread in registers ----
        const int npiv = (c0 < p) ? ((p - c0) < wj ? (p - c0) : wj) : 0;
        for (int s0 = 0; s0 < npiv; s0 += SPW) {
            const int ns = (npiv - s0) < SPW ? (npiv - s0) : SPW;
            const int kk0 = c0 + s0;                  // front row / column of the sub-panel's first pivot
            const int row = kk0 + tid;                // this thread's row (m - kk0 <= 32 NW; RPT == 2: its second row is NT below)
            // a[0] is always the CURRENT pivot column's entry of this thread's row, a[1..] the columns right of it: after every pivot the
            // row shifts left by one, finished values (multipliers l, the pivot row's U entries) go to shared memory at once.  The pivot
            // loops are therefore ROLLED: unrolled 8x (two paths) they were ~40 KB of straight-line code that every front runs once, on a
            // single warp — instruction-fetch bound (cycle profile r2s: ~1 800 cycles per pivot for a ~300-cycle dependent chain).
            T a[SPW];
#pragma unroll
            for (int k = 0; k < SPW; k++) a[k] = (k < ns && row < m) ? Cs[row + (size_t)ldc * (s0 + k)] : S::zero();
            const int row2 = row + NT;                // RPT == 2: the thread's second row
            T a2[RPT == 2 ? SPW : 1];
            if (RPT == 2) {
#pragma unroll
                for (int k = 0; k < SPW; k++) a2[RPT == 2 ? k : 0] = (k < ns && row2 < m) ? Cs[row2 + (size_t)ldc * (s0 + k)] : S::zero();
            }
            // a second row is updated like any row below the pivot: l = a2[0] / pivot, shift-update with the pivot row
            auto second_row = [&](int k, T rp, const T *prow) {
                if (RPT == 2 && row2 < m) {
                    const T l = S::mul(a2[0], rp);
                    Cs[row2 + (size_t)ldc * (s0 + k)] = l;
                    grow = max(grow, growth_key(l));
#pragma unroll
                    for (int j = 1; j < SPW; j++) a2[RPT == 2 ? j - 1 : 0] = S::fnma(a2[RPT == 2 ? j : 0], l, prow[j]);
                    a2[RPT == 2 ? SPW - 1 : 0] = S::zero();
                }
            };
            if (p - kk0 <= 32) {
                // every candidate row of this sub-panel belongs to warp 0 (always the case for mid-size fronts, p <= 32): that warp runs the
                // whole pivot sequence with warp primitives and publishes the pivot rows; the other warps apply them after ONE barrier
                T *pub = slot_rows;                  // [SPW][SPW] pivot rows (shifted: pub[k][j] = U(kk, kk + j)), then [SPW] reciprocal pivots
                if (wy == 0) {
#pragma unroll 1
                    for (int k = 0; k < ns; k++) {
                        const int kk = kk0 + k;
                        const double mg = FrontAbs<CPLX>::mag(a[0]);
                        const bool cand = lx > k && row < p && mg >= 0.0;
                        // usual outcome, one REDUX: no candidate's magnitude (high word) exceeds THRESH x the static pivot's
                        const uint32_t mx = __reduce_max_sync(0xffffffffu, cand ? (uint32_t)__double2hiint(mg) : 0u);
                        const double dk = __shfl_sync(0xffffffffu, mg, k);
                        bool keep = isfinite(dk) && dk > 0.0 && mx <= (uint32_t)__double2hiint(dk) + (FrontAbs<CPLX>::LOG2T << 20);
                        if (!keep) {
                            // (magnitude, lowest row) in one 64-bit key: 10 index bits replace the low mantissa bits
                            long long key = cand ? ((__double_as_longlong(mg) & ~0x3ffLL) | (long long)(1023 - lx)) : -1LL;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) { const long long ok = __shfl_xor_sync(0xffffffffu, key, o); key = ok > key ? ok : key; }
                            const double gb = key >= 0 ? __longlong_as_double(key & ~0x3ffLL) : -1.0;
                            keep = (key < 0) || (isfinite(dk) && dk > 0.0 && !(gb > FrontAbs<CPLX>::THRESH * dk));
                            if (!keep) {
                                // rows kk <-> ip: the two lanes exchange their register rows, the warp swaps the rest of the two rows (the
                                // block's other columns incl. this sub-panel's finished ones, the finished L columns left of the block)
                                const int gi = 1023 - (int)(key & 0x3ffLL), ip = kk0 + gi;
                                const int src = lx == k ? gi : lx == gi ? k : lx;
#pragma unroll
                                for (int j = 0; j < SPW; j++) a[j] = front_shfl(a[j], src);
                                for (int j = lx; j < wj; j += 32)
                                    if (j < s0 + k || j >= s0 + ns) { T *x = Cs + kk + (size_t)ldc * j, *y = Cs + ip + (size_t)ldc * j; const T t = *x; *x = *y; *y = t; }
                                for (int j = lx; j < c0; j += 32) { T *x = LUf + kk + (size_t)m * j, *y = LUf + ip + (size_t)m * j; const T t = *x; *x = *y; *y = t; }
                                if (lx == 0) { ipiv[kk] = ip; *s_nswap += 1; }
                            }
                        }
                        if (keep && lx == 0) ipiv[kk] = kk;
                        if (lx == k) {                // the pivot row: published for everybody, and final (its U entries)
#pragma unroll
                            for (int j = 0; j < SPW; j++) {
                                pub[k * SPW + j] = a[j];
                                if (k + j < ns) Cs[kk + (size_t)ldc * (s0 + k + j)] = a[j];
                            }
                        }
                        __syncwarp();
                        bool b1;
                        const T rp = S::rcpf(pub[k * SPW], b1);
                        bad |= b1;
                        if (lx == k) pub[SPW * SPW + k] = rp;
                        if (lx > k && row < m) {
                            const T l = S::mul(a[0], rp);
                            Cs[row + (size_t)ldc * (s0 + k)] = l;
                            grow = max(grow, growth_key(l));
#pragma unroll
                            for (int j = 1; j < SPW; j++) a[j - 1] = S::fnma(a[j], l, pub[k * SPW + j]);
                            a[SPW - 1] = S::zero();
                        }
                    }
                }
                __syncthreads();
                if (wy > 0 && row < m) {
#pragma unroll 1
                    for (int k = 0; k < ns; k++) {
                        const T l = S::mul(a[0], pub[SPW * SPW + k]);
                        Cs[row + (size_t)ldc * (s0 + k)] = l;
                        grow = max(grow, growth_key(l));
#pragma unroll
                        for (int j = 1; j < SPW; j++) a[j - 1] = S::fnma(a[j], l, pub[k * SPW + j]);
                        a[SPW - 1] = S::zero();
                    }
                }
                if (RPT == 2) {
#pragma unroll 1
                    for (int k = 0; k < ns; k++) second_row(k, pub[SPW * SPW + k], pub + k * SPW);
                }
            } else {
#pragma unroll 1
                for (int k = 0; k < ns; k++) {
                    const int kk = kk0 + k, par = k & 1;
                    // candidates: fully-summed rows below the static pivot position
                    double best = (row > kk && row < p) ? FrontAbs<CPLX>::mag(a[0]) : -1.0;
                    if (!(best >= 0.0)) best = -1.0;          // nan never wins
                    int bi = tid;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                    }
                    if (tid == bi && best >= 0.0) {
#pragma unroll
                        for (int j = 0; j < SPW; j++) slot_rows[(par * NW + wy) * SPW + j] = a[j];
                    }
                    if (lx == 0) { slot_mag[par * NW + wy] = best; slot_idx[par * NW + wy] = bi; }
                    if (tid == k) {
#pragma unroll
                        for (int j = 0; j < SPW; j++) rowk[par * SPW + j] = a[j];
                        bool bb;
                        rp_slot[par] = S::rcpf(a[0], bb);     // the usual outcome (static pivot kept): one division per pivot, not one per thread
                    }
                    __syncthreads();
                    double gb = -1.0;
                    int gi = k, gw = 0;
#pragma unroll
                    for (int w = 0; w < NW; w++) { const double v = slot_mag[par * NW + w]; if (v > gb) { gb = v; gi = slot_idx[par * NW + w]; gw = w; } }
                    const double dk = FrontAbs<CPLX>::mag(rowk[par * SPW]);
                    const bool keep = (gb < 0.0) || (isfinite(dk) && dk > 0.0 && !(gb > FrontAbs<CPLX>::THRESH * dk));
                    const T *prow = keep ? rowk + par * SPW : slot_rows + (par * NW + gw) * SPW;
                    if (!keep) {
                        // rows kk <-> ip = kk0 + gi: the two owners exchange their register rows, helper threads the rest of the two rows
                        const int ip = kk0 + gi;
                        if (tid == gi) {
#pragma unroll
                            for (int j = 0; j < SPW; j++) a[j] = rowk[par * SPW + j];
                        }
                        if (tid == k) {
#pragma unroll
                            for (int j = 0; j < SPW; j++) a[j] = prow[j];
                        }
                        for (int j = tid; j < wj; j += NT)
                            if (j < s0 + k || j >= s0 + ns) { T *x = Cs + kk + (size_t)ldc * j, *y = Cs + ip + (size_t)ldc * j; const T t = *x; *x = *y; *y = t; }
                        for (int j = tid; j < c0; j += NT) { T *x = LUf + kk + (size_t)m * j, *y = LUf + ip + (size_t)m * j; const T t = *x; *x = *y; *y = t; }
                        if (tid == 0) { ipiv[kk] = ip; *s_nswap += 1; }
                    } else if (tid == 0) ipiv[kk] = kk;
                    T rp;
                    if (keep) { rp = rp_slot[par]; bad |= !(dk > 0.0) || !isfinite(dk); }
                    else { bool b1; rp = S::rcpf(prow[0], b1); bad |= b1; }
                    if (tid == k) {                           // the pivot row is final: its U entries
#pragma unroll
                        for (int j = 0; j < SPW; j++) if (k + j < ns) Cs[kk + (size_t)ldc * (s0 + k + j)] = a[j];
                    }
                    second_row(k, rp, prow);
                    if (tid > k && row < m) {
                        const T l = S::mul(a[0], rp);
                        Cs[row + (size_t)ldc * (s0 + k)] = l;
                        grow = max(grow, growth_key(l));
#pragma unroll
                        for (int j = 1; j < SPW; j++) a[j - 1] = S::fnma(a[j], l, prow[j]);
                        a[SPW - 1] = S::zero();
                    }
                }
            }
            LLP(4);
            __syncthreads();
            const int rest = wj - (s0 + ns);          // columns of the block right of the sub-panel
            if (tid < rest) {                         // their U rows: unit-lower solve with the sub-panel's diagonal block
                T *cj = Cs + kk0 + (size_t)ldc * (s0 + ns + tid);
                T v[SPW];
#pragma unroll
                for (int k = 0; k < SPW; k++) v[k] = k < ns ? cj[k] : S::zero();
#pragma unroll
                for (int k = 0; k < SPW; k++)
#pragma unroll
                    for (int r = k + 1; r < SPW; r++)
                        if (r < ns) v[r] = S::fnma(v[r], Cs[(kk0 + r) + (size_t)ldc * (s0 + k)], v[k]);
#pragma unroll
                for (int k = 0; k < SPW; k++) if (k < ns) cj[k] = v[k];
            }
            __syncthreads();
            LLP(5);
            if (rest > 0 && m - (kk0 + ns) > 0)
                FrontMma<CPLX, NW>::update(Cs + (kk0 + ns) + (size_t)ldc * (s0 + ns), ldc, m - (kk0 + ns), rest, Cs + (kk0 + ns) + (size_t)ldc * s0, ldc,
                                       Cs + kk0 + (size_t)ldc * (s0 + ns), ldc, ns, lx, wy);
            __syncthreads();
            LLP(6);
        }
        // ---- (e) write the block out: pivot columns -> L | U11 (reciprocal pivot on the diagonal), the others -> U12 and the contribution block ----
        for (int j = wy; j < wj; j += NW) {
            const int c = c0 + j;
            const T *cj = Cs + (size_t)ldc * j;
            if (c < p) {
                for (int i = lx; i < m; i += 32) {
                    T v = cj[i];
                    if (i == c) { bool b1; v = S::rcpf(v, b1); }
                    LUf[i + (size_t)m * c] = v;
                }
            } else {
                for (int k = lx; k < p; k += 32) U12[k + (size_t)p * (c - p)] = cj[k];
                for (int i = lx; i < q; i += 32) CBf[i + (size_t)q * (c - p)] = cj[p + i];
            }
        }
        __syncthreads();
        LLP(7);
    }
#ifdef KRB_LL_PROFILE
    if (tid == 0)
        for (int i = 0; i < 8; i++) atomicAdd(krb_ll_prof + (NW == 16 ? 0 : NW == 8 ? 8 : 16) + i, (unsigned long long)llp[i]);
#endif
    if (bad && tid == 0 && P.status) atomicOr(P.status + lane, KRB_LANE_ZERO_PIVOT);
    if (grow >= KRB_GROWTH_KEY && grow < KRB_NONFINITE_KEY && P.status) atomicOr(P.status + lane, KRB_LANE_INACCURATE);
}

// ---- (b'') small fronts (m <= 32): one WARP per (front, lane), no block-wide barrier ----------------------------------------------
// The two lowest tree levels of a mesh hold ~6 600 fronts of 16..31 rows per lane (half of the factor entries, a tenth of the flops).
// With a 128-thread CTA per front they cost 3 barriers per pivot and ~15 000 warp instructions per front for ~500 useful ones
// (ncu r1: 46 % issue-active, 6 % of the DRAM peak).  Here lane i of a warp owns row i of the front (shared memory, odd leading
// dimension); a pivot step is one warp arg-max, a scaled column and a rank-1 update of the rows below with the pivot row broadcast,
// separated by __syncwarp only.  blockDim = (32, W): W independent (front, lane) pairs per CTA.
template <bool CPLX>
__global__ void __launch_bounds__(256) krb_front_factor_warp(const FrontParams P) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    extern __shared__ __align__(16) unsigned char front_smem[];
    const long long lane = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    if (lane >= P.lanes) return;
    T *F = reinterpret_cast<T *>(front_smem) + (size_t)threadIdx.y * P.smem_elems;
    const int f = P.list[blockIdx.x];
    const FrontDesc d = P.fronts[f];
    const int p = d.p, q = d.q, m = p + q, ld = m | 1;
    const int lx = threadIdx.x;
    bool bad = false;
    uint32_t grow = 0;
    int *ipiv = P.IPIV + lane * P.n + d.piv0;
    // ---- assemble ----
    for (int e = lx; e < ld * m; e += 32) F[e] = S::zero();
    __syncwarp();
    {
        const T *A = reinterpret_cast<const T *>(P.A) + lane * P.a_stride;
        for (int t = d.asm_begin + lx; t < d.asm_end; t += 32) {
            const int pos = __ldg(P.asm_pos + t), c = pos / m, r = pos - c * m;
            F[r + ld * c] = A[__ldg(P.asm_ent + t)];
        }
    }
    __syncwarp();
    for (int ci = d.child_begin; ci < d.child_end; ci++) {
        const FrontChild c = P.childrec[ci];
        const int qc = c.q;
        const T *CBc = reinterpret_cast<const T *>(P.CB) + lane * P.cb_stride + c.cb_off;
        const int *rel = P.rel + c.idx_off;
        for (int e = lx; e < qc * qc; e += 32) {
            const int j = e / qc, i = e - j * qc;
            T *dst = F + __ldg(rel + i) + ld * __ldg(rel + j);
            *dst = S::add(*dst, CBc[e]);
        }
        __syncwarp();
    }
    // ---- eliminate the p pivot columns ----
    const bool mine = lx < m;
    for (int k = 0; k < p; k++) {
        T v = mine ? F[lx + ld * k] : S::zero();
        int ip = k;
        if (k + 1 < p) {
            // (magnitude, lowest row) in one 64-bit key: 10 index bits replace the low mantissa bits
            const double mg = FrontAbs<CPLX>::mag(v);
            long long key = (lx > k && lx < p && mg >= 0.0) ? ((__double_as_longlong(mg) & ~0x3ffLL) | (long long)(1023 - lx)) : -1LL;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const long long ok = __shfl_xor_sync(0xffffffffu, key, o); key = ok > key ? ok : key; }
            const double dk = __shfl_sync(0xffffffffu, mg, k);
            const double best = key >= 0 ? __longlong_as_double(key & ~0x3ffLL) : -1.0;
            const bool keep = (key < 0) || (isfinite(dk) && dk > 0.0 && !(best > FrontAbs<CPLX>::THRESH * dk));
            if (!keep) {
                ip = 1023 - (int)(key & 0x3ffLL);
                if (mine) { T *x = F + k + ld * lx, *y = F + ip + ld * lx; const T t = *x; *x = *y; *y = t; }   // lane j swaps column j
                __syncwarp();
                v = mine ? F[lx + ld * k] : S::zero();
            }
        }
        if (lx == 0) ipiv[k] = ip;
        bool b1;
        const T rp = S::rcpf(front_shfl(v, k), b1);
        bad |= b1;
        T l = S::zero();
        const bool below = lx > k && mine;
        if (below) { l = S::mul(v, rp); F[lx + ld * k] = l; grow = max(grow, growth_key(l)); }
        if (lx == k) F[k + ld * k] = rp;                    // the factor store keeps reciprocal pivots
        int j = k + 1;
        for (; j + 4 <= m; j += 4) {
            const T u0 = F[k + ld * j], u1 = F[k + ld * (j + 1)], u2 = F[k + ld * (j + 2)], u3 = F[k + ld * (j + 3)];
            if (below) {
                T *c0 = F + lx + ld * j;
                const T a0 = c0[0], a1 = c0[ld], a2 = c0[2 * ld], a3 = c0[3 * ld];
                c0[0] = S::fnma(a0, l, u0); c0[ld] = S::fnma(a1, l, u1); c0[2 * ld] = S::fnma(a2, l, u2); c0[3 * ld] = S::fnma(a3, l, u3);
            }
        }
        for (; j < m; j++) {
            const T u0 = F[k + ld * j];
            if (below) { T *c0 = F + lx + ld * j; *c0 = S::fnma(*c0, l, u0); }
        }
        __syncwarp();
    }
    // ---- write out: L | U11 (m x p), U12 (p x q), contribution block (q x q) ----
    T *LUf = reinterpret_cast<T *>(P.LU) + lane * P.lu_stride + d.lu_off;
    for (int k = 0; k < p; k++) if (mine) LUf[lx + m * k] = F[lx + ld * k];
    T *U12 = LUf + m * p;
    T *CBf = reinterpret_cast<T *>(P.CB) + lane * P.cb_stride + d.cb_off;
    for (int j = 0; j < q; j++) {
        if (lx < p) U12[lx + p * j] = F[lx + ld * (p + j)];
        if (lx < q) CBf[lx + q * j] = F[p + lx + ld * (p + j)];
    }
    if (bad && lx == 0 && P.status) atomicOr(P.status + lane, KRB_LANE_ZERO_PIVOT);
    if (grow >= KRB_GROWTH_KEY && grow < KRB_NONFINITE_KEY && P.status) atomicOr(P.status + lane, KRB_LANE_INACCURATE);
}

// ---- iterative refinement: r = b - A x on the stamped values, in pivot order; x += correction ------------------------
struct FrontRefineParams {
    const int *res_ptr, *res_ent, *res_col;   // CSR of the permuted matrix: position k -> (CSC entry id, unknown)
    const void *A;
    long long a_stride;
    void *BP;                                  // in: b (pivot order), out: residual
    const void *X;                             // current solution, X[lane * x_lane_stride + x_row_off + unknown]
    long long x_lane_stride, x_row_off;
    void *DX;                                  // correction, DX[lane * n + unknown]
    int n;
    unsigned long long *err;                   // verify: per-lane max componentwise backward error (bits of a non-negative double), or nullptr
    int *status;
};
template <bool CPLX>
__global__ void krb_front_residual(const FrontRefineParams p) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const long long lane = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.n) return;
    const T *A = reinterpret_cast<const T *>(p.A) + lane * p.a_stride;
    const T *X = reinterpret_cast<const T *>(p.X) + lane * p.x_lane_stride + p.x_row_off;
    T *bp = reinterpret_cast<T *>(p.BP) + lane * p.n + k;
    T acc = *bp;
    if (p.err == nullptr) {
        for (int t = p.res_ptr[k]; t < p.res_ptr[k + 1]; t++) acc = S::fnma(acc, A[p.res_ent[t]], X[p.res_col[t]]);
        *bp = acc;
        return;
    }
    // verification pass: componentwise backward error of row k (Oettli-Prager),  |b - A x|_k / (|A| |x| + |b|)_k
    double den = FrontAbs<CPLX>::abs1(acc);
    for (int t = p.res_ptr[k]; t < p.res_ptr[k + 1]; t++) {
        const T a = A[p.res_ent[t]], x = X[p.res_col[t]];
        acc = S::fnma(acc, a, x);
        den = fma(FrontAbs<CPLX>::abs1(a), FrontAbs<CPLX>::abs1(x), den);
    }
    const double r = FrontAbs<CPLX>::abs1(acc);
    const double w = den > 0.0 ? r / den : (r > 0.0 ? 1.0 : 0.0);
    atomicMax(p.err + lane, (unsigned long long)__double_as_longlong(w == w ? w : 1.0));
}
// sets / clears KRB_LANE_INACCURATE from the verified backward error: the a-posteriori answer replaces the a-priori growth flag
__global__ void krb_front_verify(const unsigned long long *err, int *status, long long lanes, double tol) {
    const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    const double w = __longlong_as_double((long long)err[lane]);
    if (w <= tol) status[lane] &= ~KRB_LANE_INACCURATE;
    else status[lane] |= KRB_LANE_INACCURATE;
}
template <bool CPLX>
__global__ void krb_front_axpy(const FrontRefineParams p) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const long long lane = blockIdx.y;
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= p.n) return;
    T *x = const_cast<T *>(reinterpret_cast<const T *>(p.X)) + lane * p.x_lane_stride + p.x_row_off + u;
    *x = S::add(*x, reinterpret_cast<const T *>(p.DX)[lane * p.n + u]);
}

// ---- (c) triangular solves ------------------------------------------------------------------------------------------
// Two launch shapes: CTA per (front, lane) for large fronts, and WARP mode (blockDim = (32, W)) where every warp owns its own
// (front, lane) pair and synchronises with __syncwarp only — the lowest tree levels hold thousands of tiny fronts per lane.
// Both walk the pivot block in panels of 32 columns: the 32 x 32 diagonal block is solved by one warp with shuffles (all its
// L/U entries are loaded up front), the rows outside it are updated by one thread per row, so every factor entry is read
// exactly once, with unit stride, and never behind a dependent chain.
// non-blocking L2 prefetch of the first lines of a block (at most 32 lines per thread pass: small fronts completely, large ones
// their first panels)
__device__ __forceinline__ void front_prefetch(const void *base, size_t bytes, int tid, int nt) {
    const char *b = reinterpret_cast<const char *>(base);
    const size_t lines = (bytes + 127) / 128;
    const size_t cap = lines < (size_t)nt * 2 ? lines : (size_t)nt * 2;
    for (size_t i = tid; i < cap; i += nt) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + i * 128));
}
// asynchronous global -> shared copy of `count` elements (8 or 16 bytes each); completion: front_stage_wait + front_sync
template <typename T>
__device__ __forceinline__ void front_stage(T *dst, const T *src, int count, int tid, int nt) {
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
    if (sizeof(T) == 8 && ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)d0) & 15u) == 0) {
        // real values, both sides 16-byte aligned: two elements per copy (half the LSU instructions of a leaf front's staging)
        const int pairs = count >> 1;
        for (int i = tid; i < pairs; i += nt) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (uint32_t)i * 16u), "l"(src + 2 * i) : "memory");
        if ((count & 1) && tid == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + (uint32_t)(count - 1) * 8u), "l"(src + count - 1) : "memory");
        return;
    }
    for (int i = tid; i < count; i += nt) {
        if (sizeof(T) == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (uint32_t)i * 16u), "l"(src + i) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + (uint32_t)i * 8u), "l"(src + i) : "memory");
    }
}
__device__ __forceinline__ void front_stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <bool WARP> __device__ __forceinline__ void front_sync() { if (WARP) __syncwarp(); else __syncthreads(); }
template <bool WARP> __device__ __forceinline__ int front_any(int pred) { return WARP ? __any_sync(0xffffffffu, pred) : __syncthreads_or(pred); }

// forward: y_piv = L11^-1 P (b_piv + children), contribution = rest - L21 y_piv
template <bool CPLX, bool WARP>
__global__ void krb_front_forward(const FrontParams P) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    extern __shared__ __align__(16) unsigned char front_smem[];
    const long long lane = WARP ? (long long)blockIdx.y * blockDim.y + threadIdx.y : (long long)blockIdx.y;
    if (WARP && lane >= P.lanes) return;
    T *w = reinterpret_cast<T *>(front_smem) + (WARP ? (size_t)threadIdx.y * P.smem_elems : 0);
    const int f = P.list[blockIdx.x];
    const FrontDesc d = P.fronts[f];
    const int p = d.p, q = d.q, m = p + q;
    const int tid = threadIdx.x, NT = blockDim.x;
    T *BP = reinterpret_cast<T *>(P.BP) + lane * P.n;
    T *CBV = reinterpret_cast<T *>(P.CBV) + lane * P.cbv_stride;
    const T *L = reinterpret_cast<const T *>(P.LU) + lane * P.lu_stride + d.lu_off;
    if (P.stage_off > 0) {                                      // the whole m x p block: one round trip, overlapped with the gather below
        front_stage(w + P.stage_off, L, m * p, tid, NT);
        L = w + P.stage_off;
    } else front_prefetch(L, (size_t)m * p * sizeof(T), tid, NT);      // the factor block streams in while the children are gathered
    for (int i = tid; i < m; i += NT) w[i] = i < p ? BP[d.piv0 + i] : S::zero();
    front_sync<WARP>();
    for (int ci = d.child_begin; ci < d.child_end; ci++) {
        const FrontChild c = P.childrec[ci];
        const int *rel = P.rel + c.idx_off;
        for (int i = tid; i < c.q; i += NT) { T *dst = w + rel[i]; *dst = S::add(*dst, CBV[c.cbv_off + i]); }
        front_sync<WARP>();
    }
    {   // the lane's row interchanges inside the pivot block (usually none)
        const int *ipiv = P.IPIV + lane * P.n + d.piv0;
        int moved = 0;
        for (int k = tid; k < p; k += NT) moved |= (ipiv[k] != k);
        if (front_any<WARP>(moved)) {
            if (tid == 0)
                for (int k = 0; k < p; k++) { const int ip = ipiv[k]; if (ip != k) { const T t = w[k]; w[k] = w[ip]; w[ip] = t; } }
            front_sync<WARP>();
        }
    }
    if (P.stage_off > 0) { front_stage_wait(); front_sync<WARP>(); }
    for (int kb = 0; kb < p; kb += 32) {
        const int nbk = (p - kb) < 32 ? (p - kb) : 32;
        if (tid < 32) {
            // unit-lower diagonal block, column by column: lane t owns row kb + t and reads ITS element of column k (one conflict-free
            // column access, fetched one column ahead), y_k is broadcast by shuffle.  (ncu r3: the row-oriented form with 8 predicated
            // loads per group spent a third of a leaf front's ~630 instructions here.)
            const T *Lr = L + (kb + tid) + (size_t)m * kb;
            T val = tid < nbk ? w[kb + tid] : S::zero();
            T ln = tid < nbk ? Lr[0] : S::zero();
#pragma unroll 4
            for (int k = 0; k < nbk; k++) {
                const T l = ln;
                if (k + 1 < nbk) ln = tid < nbk ? Lr[(size_t)m * (k + 1)] : S::zero();
                const T yk = front_shfl(val, k);
                if (tid > k) val = S::fnma(val, l, yk);
            }
            if (tid < nbk) w[kb + tid] = val;
        }
        front_sync<WARP>();
        for (int i = kb + nbk + tid; i < m; i += NT) {
            T acc = w[i];
            const T *Lr = L + i + (size_t)m * kb;
            int k0 = 0;
            for (; k0 + 8 <= nbk; k0 += 8) {   // 8 loads in flight before the dependent chain
                T lv[8];
#pragma unroll
                for (int k = 0; k < 8; k++) lv[k] = Lr[(size_t)m * (k0 + k)];
#pragma unroll
                for (int k = 0; k < 8; k++) acc = S::fnma(acc, lv[k], w[kb + k0 + k]);
            }
            for (; k0 < nbk; k0++) acc = S::fnma(acc, Lr[(size_t)m * k0], w[kb + k0]);
            w[i] = acc;
        }
        front_sync<WARP>();
    }
    for (int i = tid; i < q; i += NT) CBV[d.cbv_off + i] = w[p + i];
    for (int k = tid; k < p; k += NT) BP[d.piv0 + k] = w[k];
}

// backward: x_piv = U11^-1 (y_piv - U12 x_update), delivered point-major in unknown order
template <bool CPLX, bool WARP>
__global__ void krb_front_backward(const FrontParams P) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    extern __shared__ __align__(16) unsigned char front_smem[];
    const long long lane = WARP ? (long long)blockIdx.y * blockDim.y + threadIdx.y : (long long)blockIdx.y;
    if (WARP && lane >= P.lanes) return;
    T *w = reinterpret_cast<T *>(front_smem) + (WARP ? (size_t)threadIdx.y * P.smem_elems : 0);   // [0,p): rhs -> x, [p,m): x at the update positions
    const int f = P.list[blockIdx.x];
    const FrontDesc d = P.fronts[f];
    const int p = d.p, q = d.q, m = p + q;
    T *red = w + m;                                // [NT] partial sums
    const int tid = threadIdx.x, NT = blockDim.x;
    const T *BP = reinterpret_cast<const T *>(P.BP) + lane * P.n;
    T *X = reinterpret_cast<T *>(P.X) + lane * P.x_lane_stride + P.x_row_off;
    const T *L = reinterpret_cast<const T *>(P.LU) + lane * P.lu_stride + d.lu_off;
    if (P.stage_off > 0) {
        // what the backward pass reads: the top p rows of the m x p block (U11 with the reciprocal pivots) and U12 (p x q, contiguous
        // behind the block); the L21 rows below are not touched (half of a leaf front's block)
        T *Ls = w + P.stage_off;
        if (q > 0 && p < m) {
            // column by column, thread r takes row r of each (no integer division per element: ncu r3, leaf levels of a 592-lane solve —
            // the e / p, e % p form of this loop was 15 % of the kernel's executed instructions)
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(Ls);
            if (p <= NT) {
                if (tid < p)
                    for (int k = 0; k < p; k++) {
                        const uint32_t off = (uint32_t)(tid + m * k);
                        if (sizeof(T) == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + off * 16u), "l"(L + off) : "memory");
                        else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + off * 8u), "l"(L + off) : "memory");
                    }
            } else
            for (int e = tid; e < p * p; e += NT) {
                const int k = e / p, r = e - k * p;
                const uint32_t off = (uint32_t)(r + m * k);
                if (sizeof(T) == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + off * 16u), "l"(L + off) : "memory");
                else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + off * 8u), "l"(L + off) : "memory");
            }
            front_stage(Ls + m * p, L + (size_t)m * p, p * q, tid, NT);
        } else front_stage(Ls, L, m * p + p * q, tid, NT);
        L = Ls;
    } else front_prefetch(L, ((size_t)m * p + (size_t)p * q) * sizeof(T), tid, NT);
    for (int i = tid; i < m; i += NT) w[i] = i < p ? BP[d.piv0 + i] : X[P.idx_unknown[d.idx_off + (i - p)]];
    if (P.stage_off > 0) front_stage_wait();
    front_sync<WARP>();
    const T *U12 = L + (size_t)m * p;
    if (q > 0) {
        // w[k] -= sum_j U12[k, j] x_j : G thread groups split the columns, partial sums reduced in a fixed order
        const int stride = p < NT ? p : NT;
        const int G = NT / p > 0 ? NT / p : 1;
        for (int k0 = 0; k0 < p; k0 += NT) {
            const int k = k0 + (tid % stride), g = tid / stride;
            T acc = S::zero();
            const bool act = k < p && g < G;
            if (act) for (int j = g; j < q; j += G) acc = S::fnma(acc, U12[k + (size_t)p * j], w[p + j]);
            red[tid] = acc;
            front_sync<WARP>();
            if (act && g == 0) {
                T s = w[k];
                for (int gg = 0; gg < G; gg++) s = S::add(s, red[(k - k0) + gg * stride]);
                w[k] = s;
            }
            front_sync<WARP>();
        }
    }
    for (int kb = ((p - 1) / 32) * 32; kb >= 0; kb -= 32) {
        const int nbk = (p - kb) < 32 ? (p - kb) : 32;
        if (tid < 32) {
            // upper-triangular diagonal block (reciprocal pivots on the diagonal), column by column from the last: lane t owns row kb + t
            const T *Ur = L + (kb + tid) + (size_t)m * kb;
            T val = tid < nbk ? w[kb + tid] : S::zero();
            T un = tid < nbk ? Ur[(size_t)m * (nbk - 1)] : S::zero();
#pragma unroll 4
            for (int k = nbk - 1; k >= 0; k--) {
                const T u = un;
                if (k > 0) un = tid < nbk ? Ur[(size_t)m * (k - 1)] : S::zero();
                if (tid == k) val = S::mul(val, u);                    // x_k = rhs_k / pivot_k
                const T xk = front_shfl(val, k);
                if (tid < k) val = S::fnma(val, u, xk);
            }
            if (tid < nbk) w[kb + tid] = val;
        }
        front_sync<WARP>();
        for (int i = tid; i < kb; i += NT) {
            T acc = w[i];
            const T *Ur = L + i + (size_t)m * kb;
            int k0 = 0;
            for (; k0 + 8 <= nbk; k0 += 8) {
                T uv[8];
#pragma unroll
                for (int k = 0; k < 8; k++) uv[k] = Ur[(size_t)m * (k0 + k)];
#pragma unroll
                for (int k = 0; k < 8; k++) acc = S::fnma(acc, uv[k], w[kb + k0 + k]);
            }
            for (; k0 < nbk; k0++) acc = S::fnma(acc, Ur[(size_t)m * k0], w[kb + k0]);
            w[i] = acc;
        }
        front_sync<WARP>();
    }
    for (int k = tid; k < p; k += NT) X[P.col_order[d.piv0 + k]] = w[k];
}

}  // namespace krb
