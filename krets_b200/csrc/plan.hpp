// Host-side plan objects for the batched lane solver (one-time symbolic analysis, north-star "Rust host" role,
// written in C++ because this image has no Rust toolchain).
#pragma once
#include <cstdint>
#include <complex>
#include <string>
#include <vector>

#include "../../include/krets_b200.h"
#include "vm_ops.h"

namespace krb {

struct Circuit {
    std::vector<krb_element> el;
    int n = 0;
    std::vector<uint8_t> is_current;
    bool nonlinear = false;   // elements.rs:143-148
};

// ---- symbolic stamps -----------------------------------------------------------------------------------
// Where a stamped value comes from (stampable.rs).  `neg` is an exact sign flip.
enum SrcKind : uint8_t {
    S_ONE,        // 1.0 (group-2 coupling entries)
    S_G,          // 1.0 / value                           resistor   :232
    S_VAL,        // value                                 I / V DC right-hand side  :567, :1017
    S_VOH,        // value / h                             C, L backward-Euler companions  :449, :781
    S_JWV,        // (0, 2*pi*f*value)                     C, L AC    :379-382, :715
    S_ACAMP,      // (ac_amplitude, 0)                     V AC RHS   :1031
    S_WAVE,       // transient_value_at(time of guess)     V transient RHS  :1043-1048
    S_DIODE_G,    // linearised diode conductance          :597, :622
    S_DIODE_IEQ,  // diode equivalent current              :651
    S_CAP_HIST,   // (value/h) * (v+_prev - v-_prev)       :480-493  (== -i_eq)
    S_IND_HIST,   // (value/h) * i_prev                    :806-812
    // trapezoidal companions (north-star extension; not in the reference). value/h is taken with h/2, i.e. 2*value/h.
    S_CAP_IPREV,  // capacitor current of the previous time point (per-lane state column)
    S_IND_VPREV,  // v+_prev - v-_prev across an inductor
    // level-1 NMOS companion (extension, KRB_ELEM_FLAG_LEVEL1; equations nmosfet.rs:51-102 at the linearisation point)
    S_MOS_GDS,    // g_ds
    S_MOS_GM,     // g_m
    S_MOS_GDSGM,  // g_ds + g_m
    S_MOS_IEQ     // i_d - g_ds * v_ds - g_m * v_gs          (stampable.rs:889)
};
struct Src { SrcKind kind; bool neg; int elem; };
struct MatContrib { int row, col; Src s; };
struct RhsContrib { int row; Src s; };
struct Stamps {
    std::vector<MatContrib> mat;   // in the reference's push order (element order, then push order within the element)
    std::vector<RhsContrib> rhs;
};
// returns KRB_OK / KRB_ERR_UNSUPPORTED
int build_stamps(const Circuit &c, int mode, int integrator, Stamps &out);

// unique (row,col) set in CSC order + the contributions of each entry in stamp order (== sum_triplets order)
struct Pattern {
    int n = 0, nnz = 0;
    std::vector<int> colptr, rowidx;
    std::vector<std::vector<int>> contribs;   // per entry: indices into Stamps::mat
    std::vector<std::vector<int>> rhs_contribs;   // per row: indices into Stamps::rhs
};
void build_pattern(int n, const Stamps &s, Pattern &p);

// nominal |value| of every entry (for the weighted static pivot matching)
void nominal_magnitudes(const Circuit &c, const Stamps &s, const Pattern &p, int mode, double h, double fnom, std::vector<double> &mag);
// the signed complex values the magnitudes are taken of
void nominal_values(const Circuit &c, const Stamps &s, const Pattern &p, int mode, double h, double fnom, std::vector<std::complex<double>> &val);

// ---- static pivot plan ----------------------------------------------------------------------------------
// maximum-product transversal: match_row[col] = row used as the pivot equation of unknown `col`; false if
// structurally singular
// diag_bonus (natural-log units): a diagonal entry is treated as if it were exp(diag_bonus) times larger.  The front executor
// pivots dynamically inside its fronts, so it prefers the structurally symmetric (diagonal) matching, which keeps the fill of
// the symmetrised pattern low; the one-thread-per-lane executors use 0 (pure maximum product).
bool weighted_matching(const Pattern &p, const std::vector<double> &mag, std::vector<int> &match_row, double diag_bonus = 0.0);
// minimum-degree ordering of the symmetrised, row-permuted pattern; returns order[k] = unknown eliminated k-th
void minimum_degree(const Pattern &p, const std::vector<int> &match_row, std::vector<int> &order);
// identity row matching; leader_of[u] >= 0: unknown u (structurally zero diagonal) is eliminated right after unknown leader_of[u]
void minimum_degree_paired(const Pattern &p, const std::vector<int> &leader_of, std::vector<int> &order);
void rcm_order(const Pattern &p, const std::vector<int> &match_row, std::vector<int> &order);
bool band_ok(const Pattern &p, const std::vector<int> &match_row, const std::vector<int> &order);

struct Symbolic {
    int n = 0;
    std::vector<int> row_order, col_order;        // pivot k: equation row_order[k], unknown col_order[k]
    // permuted matrix C[k][l] = A[row_order[k]][col_order[l]]; per pivot row: L part (l<k), U part (l>k), both sorted
    std::vector<std::vector<int>> Lpat, Upat;
    // A entry id for (k, l) positions present in A: parallel to Lpat/Upat (-1 for fill) and the diagonal
    std::vector<std::vector<int>> Lent, Uent;
    std::vector<int> Dent;
    int64_t nnz_lu = 0, flops = 0;
};
void symbolic_lu(const Pattern &p, const std::vector<int> &match_row, const std::vector<int> &order, Symbolic &s);

// ---- compiled lane program -----------------------------------------------------------------------------
enum PoolKind : uint8_t { P_CONST, P_G, P_VAL, P_VOH, P_ACAMP, P_DIO_ISNVT, P_DIO_NVT, P_DIO_IS, P_DIO_VCRIT, P_MOS_VTH, P_MOS_BETA, P_MOS_LAMBDA };
struct PoolEntry { PoolKind kind; bool neg; int elem; double c; };

struct Program {
    std::vector<uint32_t> ops;   // 2 words per op, KOP_END terminated
    int64_t n_ops = 0;
};

struct DeviceState;
struct FrontDevice;

// "chain" executor tables (chain_kernel.cuh): the permuted matrix is tridiagonal, so each pivot row is described by a
// short run of stamp terms for its sub-diagonal, diagonal, super-diagonal and right-hand side.
struct ChainTerm { uint32_t meta; uint32_t plane; double K; };   // 16 B, one GENERIC-row term (chain_kernel.cuh)
struct ChainPlan {
    bool valid = false;
    std::vector<uint32_t> rowinfo;       // per pivot row (+1 sentinel): bit0 FIXED, bit1 HAS_U2, bits 8.. = #terms of a GENERIC row
    std::vector<uint64_t> stream;        // 16-byte units (2 words each): FIXED records / GENERIC terms, padded with 4 zero units
    std::vector<int> unknown_of_row;     // = Symbolic::col_order
    int n_u2 = 0;                        // rows with a second super-diagonal entry (extra planes 2n .. 2n+n_u2)
    int n_fixed = 0;                     // rows encoded as FIXED records
    std::vector<int> u2_rows;            // rows owning a u2 plane, ascending
    std::vector<uint32_t> runs;          // (count << 1) | is_fixed: consecutive rows of the same record type
    bool identity_order = false;         // unknown_of_row[k] == k: results can be delivered point-major straight from the kernel
    // ---- block-partitioned kernel (chain_block_kernel.cuh): one CTA of T threads per lane, M consecutive pivot rows per thread ----
    struct Block {
        bool valid = false;
        int T = 0, M = 0;                // n <= T * M; rows >= n are identity padding
        std::vector<uint32_t> rows;      // [M][T][2]: word 0 = FIXED ? (dictionary index << 1) | 1 : (stream unit offset << 1);
                                         //            word 1 = (HAS_U2 ? 2 : 0) | #terms << 2 | slot << 16
        std::vector<uint64_t> dict;      // unique FIXED records, 4 (real) / 8 (complex) words each, same layout as in `stream`
        int n_dict = 0;
        std::vector<int> gen_rows;       // GENERIC pivot rows, ascending (slot order)
        bool shared_dict = false;        // n_dict + #GENERIC rows <= 256: the kernel keeps every row's stamped values in a shared-memory table
        double nominal_diff = 0.0;       // max scaled |x_block - x_sequential| at the nominal inputs (host check, must be <= 5e-10)
        std::string why_not;             // reason the block kernel is not applicable (diagnostics)
    } block;
};

// "front" executor tables (front_kernels.cuh): supernodal multifrontal LU on the symmetrised pattern of the permuted
// matrix, for circuits whose factor is far too large for one thread per lane (meshes).  Everything is
// in *position space*: position k is pivot k of the final elimination order (equation row_order[k], unknown col_order[k]).
struct FrontPlan {
    bool valid = false;
    int n = 0, n_fronts = 0, n_levels = 0, max_m = 0, max_p = 0;
    std::vector<int> row_order, col_order;
    std::vector<int> piv0;                 // [n_fronts + 1] front f owns pivots [piv0[f], piv0[f+1])
    std::vector<int> idx_ptr, idx;         // update index list of each front (positions, ascending); q = idx_ptr[f+1]-idx_ptr[f]
    std::vector<int> parent;               // parent front or -1
    std::vector<int> child_ptr, child;     // children, ascending
    std::vector<int> rel_ptr, rel;         // per front (as a child): local index inside the parent's front of each update position
    std::vector<int> level_ptr, level_fronts;   // fronts grouped by level (leaves first), largest front first inside a level
    std::vector<int> asm_ptr, asm_ent, asm_pos; // stamped entries assembled into each front: CSC entry id, local r + m * c
    std::vector<int64_t> lu_off;           // factor block of front f in the per-lane factor store: m x p column-major (L incl. the
                                           // pivot block, reciprocal pivots on its diagonal), then the p x q block U12 column-major
    std::vector<int64_t> cb_off;           // q x q contribution block in the per-lane workspace
    std::vector<int64_t> cbv_off;          // q contribution vector (forward substitution)
    int64_t lu_size = 0, cb_size = 0, cbv_size = 0;
    int64_t nnz_lu = 0;                    // stored factor entries incl. explicit zeros of relaxed supernodes (= lu_size)
    int64_t flops = 0;                     // factor + one solve, real flops of the real-valued problem
};
void build_front_plan(const Pattern &p, const std::vector<int> &match_row, const std::vector<int> &order, FrontPlan &fp);

// "dense transient" executor tables (dense_tran_kernel.cuh): small linear circuits (n <= 8) stepped natively
struct DensePlan {
    bool valid = false;
    struct Elem { double value; int in_plane, pp, pm, pc, wave_col, state_idx; };
    struct Term { int k, l, kind, elem; };
    std::vector<Elem> elems;
    std::vector<Term> mat_terms, rhs_terms;
};

struct Plan {
    Circuit circuit;
    DeviceState *dev = nullptr;        // device-side buffers, created lazily by the runtime
    int mode = 0;
    krb_plan_opts opts{};
    std::vector<int> varied;            // element ordinals with per-lane values
    bool is_complex = false;
    Stamps stamps;
    Pattern pattern;
    Symbolic sym;
    // program
    Program sec[3];                     // 0: main (one-shot / per-NR-iteration), 1: factor-once, 2: per-step
    bool has_step_sections = false;     // linear transient: sec[1] once + sec[2] per step
    bool repivoted = false;             // the transversal's pivot sequence cancelled at the nominal values and was re-pivoted numerically (compiler.cpp)
    std::vector<PoolEntry> pool;        // B table layout (static part)
    std::vector<int> wave_elems;        // per-step table columns (LDBS): V sources with a transient RHS
    int n_slots = 0, n_planes = 0, n_inputs = 0;
    int input_sweep = -1, input_freq = -1;   // input plane holding the DC sweep value / the AC frequency
    std::vector<int> input_of_varied;        // input plane of varied[k]
    int64_t flops_per_point = 0, hbm_bytes_per_point = 0, alg_bytes_per_point = 0;
    double h_comp = 0.0;                // companion-model step: h (backward Euler) or h/2 (trapezoidal: g = 2C/h)
    std::vector<int> state_elems;       // trapezoidal: capacitors whose current is carried in result-row column n + j
    ChainPlan chain;                    // filled when the structure allows the native chain executor
    bool use_chain = false;             // executor selected for krb_solve_* (chain.valid and not forced to the lane VM)
    int64_t vm_hbm_bytes_per_point = 0;
    DensePlan dense;                    // linear transient plans of small circuits: native stepping kernel
    FrontPlan front;                    // filled when the front executor is selected
    bool use_front = false;
    FrontDevice *fdev = nullptr;        // device tables of the front executor (front_exec.cu)
    std::vector<int> front_input_of_elem;   // element -> per-lane input plane or -1
};

// full pipeline; returns KRB_* status and fills `err`
int analyze(const Circuit &c, int mode, const krb_plan_opts &opts, Plan &plan, std::string &err);

// evaluate the static B table for the circuit's current element values (host, per call)
void evaluate_pool(const Plan &plan, const Circuit &c, double h, std::vector<double> &out);
// reference waveform evaluation (voltage_source.rs:52-93, 244-252)
double transient_value_at(const krb_element &e, double time);

}  // namespace krb
