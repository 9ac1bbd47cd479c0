// sm_100a kernel for the transient stepping of SMALL linear circuits (n <= 8): BASELINE config 4 is dual_rc_ladder, n = 4,
// 10^4 steps x 16 384 Monte-Carlo lanes.  With so few lanes (512 warps on 592 warp schedulers) the run time is the length of
// the dependent chain per step, so the per-step work must be native code, not interpreted lane-program ops:
//
//   once per lane : stamp the dense companion matrix (stampable.rs:442-468, 780-792, 972-980 ...), LU with the plan's static
//                   pivots; turn the right-hand-side stamps (:470-505, :806-815, :1043-1048) into dense coefficient rows
//                   b = H1 + Hx x_prev + Hs state + Hw sources(t)
//   every step    : propagator mode (default): A is constant over the steps, so x_n = A^-1 b = G z with z = (1, x_prev, state,
//                   sources) and G = A^-1 [H1 | Hx | Hs | Hw] solved once per lane with the LU factors: a step is N independent
//                   dot products (two accumulators each) - no forward / backward substitution in the loop.  G is permuted
//                   into UNKNOWN order on its way to registers, so that a solved point leaves as ONE 256-bit (N = 4) or
//                   128-bit (N = 2) store per lane: lanes are rows * n values apart, a scalar store of a warp touches 32
//                   sectors, and that store path was what bounded the kernel (config 4: 4.7 -> 2.5 ms from the store alone).
//                   The source table is staged through shared memory in chunks of 32 steps.  Measured on B200 (tools/
//                   microbench/fp64_lat.cu): DFMA dependent latency 8 cycles, one warp issues a DFMA every 2 cycles.
//                   KRB_DENSE_TRAN_LU=1: forward / backward substitution with the LU factors every step (the work
//                   structure of transient.rs:43-112), scalar stores; the GPU tests hold both against the oracle.
//                   Then trapezoidal capacitor currents; without direct delivery one coalesced store per unknown into the
//                   lane-interleaved result planes X[row][unknown][lane] (the layout krb_pack_point_major delivers from).
//
// Coefficients live in per-thread shared-memory columns C[idx][tid] (conflict-free, static offsets after unrolling).
#pragma once
#include <cuda_runtime.h>

#include "scalar.cuh"
#include <stdint.h>

#include "plan.hpp"

namespace krb {

struct DenseElem { double value; int in_plane, pp, pm, pc, wave_col, state_idx; };   // pp/pm/pc: pivot positions of V(plus), V(minus), I(ident) or -1
struct DenseTerm { int k, l, kind, elem; };                                         // matrix: position (k, l); rhs: row k (l unused); kind = SrcKind | neg << 8

struct DenseTranParams {
    const DenseElem *elems;
    const DenseTerm *mat_terms, *rhs_terms;
    int n_mat, n_rhs;
    const int *state_elems;      // [ns] element ordinals carrying a trapezoidal state
    const int *col_order;        // position -> unknown
    int n, ns, nw, width;        // width = n + ns result columns per row
    const double *IN;            // [n_inputs][L]
    const double *Bstep;         // [rows][2][nw] source table (iteration-0 entries are used: linear circuit)
    double *X;                   // [rows][width][L]; with OUT set only row 0 (the operating point) is read from it
    double *OUT;                 // optional: point-major delivery x[lane][row][unknown] straight from the kernel
    long long out_lanes;         // lanes that own storage in OUT
    long long L;
    int rows;
    double h_comp;
    int *status;
    int propagate;               // 1: steps apply the per-lane propagator G = A^-1 [H1 | Hx | Hs | Hw] (N independent FMA rows);
                                 // 0: forward / backward substitution with the LU factors every step (KRB_DENSE_TRAN_LU=1)
};

// N unknowns, NSP / NWP = padded counts of trapezoidal state columns and source-table columns (zero coefficients beyond
// ns / nw), all compile-time so that the step loop is straight-line code.  For N <= 4 every coefficient is copied from the
// shared-memory columns into registers before the loop: a step is then ~60 instructions of FMA chain.
template <int N, int NSP, int NWP, int TPB>
__global__ void __launch_bounds__(TPB) krb_dense_tran(const DenseTranParams p) {
    extern __shared__ double dense_smem[];
    const int tid = threadIdx.x;
    const long long lane = (long long)blockIdx.x * TPB + tid;
    const long long L = p.L;
    const int ns = p.ns, nw = p.nw;
    // per-thread columns: A[N*N] | Hx[N*N] | H1[N] | Hs[N*NSP] | Hw[N*NWP] | Gx[NSP*N]
    constexpr int O_A = 0, O_HX = N * N, O_H1 = 2 * N * N, O_HS = O_H1 + N, O_HW = O_HS + N * NSP, O_GX = O_HW + N * NWP, TOTAL = O_GX + NSP * N;
    double *C = dense_smem + tid;
    for (int i = 0; i < TOTAL; i++) C[(size_t)i * TPB] = 0.0;
    auto value_of = [&](const DenseElem &el) { return el.in_plane >= 0 ? p.IN[(size_t)el.in_plane * L + lane] : el.value; };

    // ---- companion matrix in pivot order, summed in stamp order (== sum_triplets, solver.rs:59-70) ----
    for (int t = 0; t < p.n_mat; t++) {
        const DenseTerm tm = p.mat_terms[t];
        const DenseElem el = p.elems[tm.elem];
        const double v = value_of(el);
        double c = 0.0;
        switch (tm.kind & 0xff) {
        case S_ONE: c = 1.0; break;
        case S_G: c = __ddiv_rn(1.0, v); break;
        case S_VAL: c = v; break;
        case S_VOH: c = __ddiv_rn(v, p.h_comp); break;
        default: break;
        }
        if (tm.kind & 0x100) c = -c;
        C[(size_t)(O_A + tm.k * N + tm.l) * TPB] += c;
    }
    // ---- right-hand side as coefficient rows: b = H1 + Hx x_prev + Hs state + Hw sources(t) ----
    for (int t = 0; t < p.n_rhs; t++) {
        const DenseTerm tm = p.rhs_terms[t];
        const DenseElem el = p.elems[tm.elem];
        const double v = value_of(el);
        const double sg = (tm.kind & 0x100) ? -1.0 : 1.0;
        const int k = tm.k;
        switch (tm.kind & 0xff) {
        case S_VAL: C[(size_t)(O_H1 + k) * TPB] += sg * v; break;
        case S_WAVE: C[(size_t)(O_HW + k * NWP + el.wave_col) * TPB] += sg; break;
        case S_CAP_HIST: {                                   // (C/h) (v+_prev - v-_prev)
            const double c = sg * __ddiv_rn(v, p.h_comp);
            if (el.pp >= 0) C[(size_t)(O_HX + k * N + el.pp) * TPB] += c;
            if (el.pm >= 0) C[(size_t)(O_HX + k * N + el.pm) * TPB] -= c;
            break;
        }
        case S_IND_HIST: if (el.pc >= 0) C[(size_t)(O_HX + k * N + el.pc) * TPB] += sg * __ddiv_rn(v, p.h_comp); break;   // (L/h) i_prev
        case S_CAP_IPREV: C[(size_t)(O_HS + k * NSP + el.state_idx) * TPB] += sg; break;
        case S_IND_VPREV:
            if (el.pp >= 0) C[(size_t)(O_HX + k * N + el.pp) * TPB] += sg;
            if (el.pm >= 0) C[(size_t)(O_HX + k * N + el.pm) * TPB] -= sg;
            break;
        default: break;
        }
    }
    // trapezoidal capacitor currents: i_n = (2C/h) ((v+ - v-)_n - (v+ - v-)_{n-1}) - i_{n-1}
    for (int j = 0; j < ns; j++) {
        const DenseElem el = p.elems[p.state_elems[j]];
        const double g = __ddiv_rn(value_of(el), p.h_comp);
        if (el.pp >= 0) C[(size_t)(O_GX + j * N + el.pp) * TPB] += g;
        if (el.pm >= 0) C[(size_t)(O_GX + j * N + el.pm) * TPB] -= g;
    }
    // ---- LU with the static pivots (unit lower, reciprocal pivots on the diagonal) ----
    bool bad = false, nonfinite = false;
    uint32_t grow = 0;   // pivot growth: largest multiplier |l| = |a_ik / pivot| (scalar.cuh: growth_key, KRB_LANE_INACCURATE)
#pragma unroll
    for (int k = 0; k < N; k++) {
        const double d = C[(size_t)(O_A + k * N + k) * TPB];
        bad |= !(fabs(d) > 0.0) || !isfinite(d);
        const double r = 1.0 / d;
        C[(size_t)(O_A + k * N + k) * TPB] = r;
#pragma unroll
        for (int i = k + 1; i < N; i++) {
            const double l = C[(size_t)(O_A + i * N + k) * TPB] * r;
            grow = max(grow, growth_key(l));
            C[(size_t)(O_A + i * N + k) * TPB] = l;
#pragma unroll
            for (int j = k + 1; j < N; j++)
                C[(size_t)(O_A + i * N + j) * TPB] = fma(-l, C[(size_t)(O_A + k * N + j) * TPB], C[(size_t)(O_A + i * N + j) * TPB]);
        }
    }
    // ---- propagator: with A constant over the steps, x_n = A^-1 b = G z, z = (1, x_prev, state, sources).  The columns of
    // [H1 | Hx | Hs | Hw] are solved once, in place; a step is then N independent dot products (two accumulators each) instead
    // of a forward and a backward substitution whose ~3N dependent FP64 operations set the run time at one warp per scheduler.
    const bool prop = p.propagate != 0;
    if (prop) {
        auto solve_col = [&](int base, int stride) {
            double v[N];
#pragma unroll
            for (int k = 0; k < N; k++) v[k] = C[(size_t)(base + k * stride) * TPB];
#pragma unroll
            for (int k = 1; k < N; k++)
#pragma unroll
                for (int l = 0; l < k; l++) v[k] = fma(-C[(size_t)(O_A + k * N + l) * TPB], v[l], v[k]);
#pragma unroll
            for (int k = N - 1; k >= 0; k--) {
                double acc = v[k];
#pragma unroll
                for (int l = k + 1; l < N; l++) acc = fma(-C[(size_t)(O_A + k * N + l) * TPB], v[l], acc);
                v[k] = acc * C[(size_t)(O_A + k * N + k) * TPB];
            }
#pragma unroll
            for (int k = 0; k < N; k++) C[(size_t)(base + k * stride) * TPB] = v[k];
        };
        solve_col(O_H1, 1);
#pragma unroll
        for (int l = 0; l < N; l++) solve_col(O_HX + l, N);
#pragma unroll
        for (int j = 0; j < NSP; j++) solve_col(O_HS + j, NSP);
#pragma unroll
        for (int j = 0; j < NWP; j++) solve_col(O_HW + j, NWP);
    }
    constexpr bool REGS = (N <= 4);
    const bool direct = p.OUT != nullptr;
    const bool owns_out = !direct || lane < p.out_lanes;                                     // padding lanes compute but own no output
    // Point-major delivery with the propagator (N = 2, 4): the rows and columns of G are permuted into UNKNOWN order while they
    // are copied to registers, so that a solved point is one 128 / 256-bit store per lane instead of N scalar ones: lanes are
    // rows * n values apart, every scalar store of a warp touches 32 sectors, and that store path - not the arithmetic - is
    // what bounds this kernel (config 4: 4 x 8 B stores per lane and step -> one 32 B store).
    constexpr bool VEC = REGS && (N == 2 || N == 4);
    const bool ordered = VEC && prop && direct && ((reinterpret_cast<uintptr_t>(p.OUT) & (N * 8 - 1)) == 0);
    double rc[REGS ? TOTAL : 1];
    if (REGS) {
        if (ordered) {
            int pos[N];                                                                      // pivot position of unknown u
#pragma unroll
            for (int l = 0; l < N; l++) {
                const int u = p.col_order[l];
#pragma unroll
                for (int v = 0; v < N; v++) if (v == u) pos[v] = l;
            }
#pragma unroll
            for (int u = 0; u < N; u++) {
                const int l = pos[u];
                rc[REGS ? O_H1 + u : 0] = C[(size_t)(O_H1 + l) * TPB];
#pragma unroll
                for (int v = 0; v < N; v++) rc[REGS ? O_HX + u * N + v : 0] = C[(size_t)(O_HX + l * N + pos[v]) * TPB];
#pragma unroll
                for (int j = 0; j < NSP; j++) rc[REGS ? O_HS + u * NSP + j : 0] = C[(size_t)(O_HS + l * NSP + j) * TPB];
#pragma unroll
                for (int j = 0; j < NWP; j++) rc[REGS ? O_HW + u * NWP + j : 0] = C[(size_t)(O_HW + l * NWP + j) * TPB];
            }
#pragma unroll
            for (int j = 0; j < NSP; j++)
#pragma unroll
                for (int v = 0; v < N; v++) rc[REGS ? O_GX + j * N + v : 0] = C[(size_t)(O_GX + j * N + pos[v]) * TPB];
        } else {
#pragma unroll
            for (int i = 0; i < TOTAL; i++) rc[REGS ? i : 0] = C[(size_t)i * TPB];
        }
    }
#define KRB_DC(i) (REGS ? rc[REGS ? (i) : 0] : C[(size_t)(i) * TPB])
    double x[N], st[NSP > 0 ? NSP : 1];
    double *xp[N];                                                                           // store cursors, one per unknown
    double *pv = direct ? p.OUT + (size_t)lane * p.rows * p.n : nullptr;                     // ordered: cursor of the whole point
    auto store_point = [&](double *dst, const double *v) {
        if (N == 4) asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(v[0]), "d"(v[N > 1 ? 1 : 0]), "d"(v[N > 2 ? 2 : 0]), "d"(v[N > 3 ? 3 : 0]) : "memory");
        else asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(v[0]), "d"(v[N > 1 ? 1 : 0]) : "memory");
    };
#pragma unroll
    for (int l = 0; l < N; l++) {
        const int u = ordered ? l : p.col_order[l];
        x[l] = p.X[(size_t)u * L + lane];                                                    // row 0: the operating point
        xp[l] = direct ? p.OUT + (size_t)lane * p.rows * p.n + u : p.X + (size_t)u * L + lane;
        if (direct && !ordered && owns_out) *xp[l] = x[l];
    }
    if (ordered && owns_out) store_point(pv, x);
#pragma unroll
    for (int j = 0; j < NSP; j++) st[j] = j < ns ? p.X[(size_t)(p.n + j) * L + lane] : 0.0;   // i_C(0)
    const size_t row_stride = direct ? (size_t)p.n : (size_t)p.width * L;
    // The source table is the same for every lane: the warp (TPB == 32: one warp per block) stages it through shared memory in
    // chunks of 32 steps, the next chunk loaded into registers while the current one is consumed (one L2 latency per 32 steps
    // and off the critical path, instead of one exposed global load per step).
    static_assert(TPB == 32, "source-table staging assumes one warp per block");
    double *ws = dense_smem + (size_t)TOTAL * TPB;             // [2][32][NWP]
    auto fetch_chunk = [&](int c, double *dst) {               // thread tid fetches the sources of step c * 32 + tid
        const int r = c * 32 + tid;
#pragma unroll
        for (int j = 0; j < NWP; j++) dst[j] = (j < nw && r < p.rows) ? __ldg(p.Bstep + (size_t)r * 2 * nw + j) : 0.0;
    };
    {
        double w0[NWP];
        fetch_chunk(0, w0);
#pragma unroll
        for (int j = 0; j < NWP; j++) ws[tid * NWP + j] = w0[j];
        __syncwarp();
    }
    double wnext[NWP];
    fetch_chunk(1, wnext);

    for (int row = 1; row < p.rows; row++) {
        if ((row & 31) == 0) {                                 // publish the chunk fetched during the previous 32 steps, fetch the next
            double *dstc = ws + (size_t)((row >> 5) & 1) * 32 * NWP;
#pragma unroll
            for (int j = 0; j < NWP; j++) dstc[tid * NWP + j] = wnext[j];
            __syncwarp();
            fetch_chunk((row >> 5) + 1, wnext);
        }
        // (reading the NEXT step's sources here, one step ahead, and unrolling two steps were both measured slower)
        double wv[NWP];
#pragma unroll
        for (int j = 0; j < NWP; j++) wv[j] = ws[(size_t)((row >> 5) & 1) * 32 * NWP + (row & 31) * NWP + j];
        double xn[N];
        if (prop) {
#pragma unroll
            for (int k = 0; k < N; k++) {
                // two accumulators; x_prev and the state first (both in registers when the step starts), the source terms
                // last: their shared-memory load was issued at the top of the step and completes behind the other terms
                double a0 = KRB_DC(O_H1 + k), a1 = 0.0;
#pragma unroll
                for (int l = 0; l < N; l++) {
                    if (l & 1) a1 = fma(KRB_DC(O_HX + k * N + l), x[l], a1); else a0 = fma(KRB_DC(O_HX + k * N + l), x[l], a0);
                }
#pragma unroll
                for (int j = 0; j < NSP; j++) {
                    if (j & 1) a1 = fma(KRB_DC(O_HS + k * NSP + j), st[j], a1); else a0 = fma(KRB_DC(O_HS + k * NSP + j), st[j], a0);
                }
#pragma unroll
                for (int j = 0; j < NWP; j++) {
                    if (j & 1) a0 = fma(KRB_DC(O_HW + k * NWP + j), wv[j], a0); else a1 = fma(KRB_DC(O_HW + k * NWP + j), wv[j], a1);
                }
                xn[k] = a0 + a1;
            }
        } else {
            double b[N];
#pragma unroll
            for (int k = 0; k < N; k++) {
                double acc = KRB_DC(O_H1 + k);
#pragma unroll
                for (int l = 0; l < N; l++) acc = fma(KRB_DC(O_HX + k * N + l), x[l], acc);
#pragma unroll
                for (int j = 0; j < NSP; j++) acc = fma(KRB_DC(O_HS + k * NSP + j), st[j], acc);
#pragma unroll
                for (int j = 0; j < NWP; j++) acc = fma(KRB_DC(O_HW + k * NWP + j), wv[j], acc);
                b[k] = acc;
            }
#pragma unroll
            for (int k = 1; k < N; k++)
#pragma unroll
                for (int l = 0; l < k; l++) b[k] = fma(-KRB_DC(O_A + k * N + l), b[l], b[k]);
#pragma unroll
            for (int k = N - 1; k >= 0; k--) {
                double acc = b[k];
#pragma unroll
                for (int l = k + 1; l < N; l++) acc = fma(-KRB_DC(O_A + k * N + l), xn[l], acc);
                xn[k] = acc * KRB_DC(O_A + k * N + k);
            }
        }
#pragma unroll
        for (int j = 0; j < NSP; j++) {
            double a0 = -st[j], a1 = 0.0;
#pragma unroll
            for (int l = 0; l < N; l++) {
                if (prop && (l & 1)) a1 = fma(KRB_DC(O_GX + j * N + l), xn[l] - x[l], a1); else a0 = fma(KRB_DC(O_GX + j * N + l), xn[l] - x[l], a0);
            }
            st[j] = prop ? a0 + a1 : a0;
        }
        if (ordered) { pv += N; if (owns_out) store_point(pv, xn); }
#pragma unroll
        for (int l = 0; l < N; l++) {
            if (!ordered) { xp[l] += row_stride; if (owns_out) *xp[l] = xn[l]; }
            x[l] = xn[l];
        }
        // a linear recurrence never comes back from inf / NaN (inf - inf, 0 * inf = NaN): testing every 32nd step and the last
        // one finds every lane that overflowed at any step
        if ((row & 31) == 31 || row + 1 == p.rows) {
            double chk = 0.0;
#pragma unroll
            for (int l = 0; l < N; l++) chk += xn[l] * 0.0;    // NaN iff any unknown is inf / NaN
            nonfinite |= (chk != 0.0);
        }
    }
#undef KRB_DC
    const bool inacc = grow >= KRB_GROWTH_KEY && grow < KRB_NONFINITE_KEY;
    if (bad || nonfinite || inacc) p.status[lane] |= (bad ? KRB_LANE_ZERO_PIVOT : 0) | (nonfinite ? KRB_LANE_NONFINITE : 0) | (inacc ? KRB_LANE_INACCURATE : 0);
}

}  // namespace krb
