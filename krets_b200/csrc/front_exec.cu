// Runtime of the front executor (front_kernels.cuh): device tables of a FrontPlan, per-level launch lists, lane chunking
// by device memory, and the analysis drivers (one-shot DC/AC solves, transient stepping with factors kept in HBM).
// No CPU fallback: without a CUDA device every entry point returns KRB_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "front_kernels.cuh"
#include "plan.hpp"
#include "runtime.hpp"

namespace krb {

#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);                    \
            return (_e == cudaErrorMemoryAllocation) ? KRB_ERR_OUT_OF_MEMORY : KRB_ERR_CUDA;                   \
        }                                                                                                      \
    } while (0)

namespace {
struct FBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return KRB_OK;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        CUDA_TRY(cudaMalloc(&p, std::max<size_t>(need, 16)));
        bytes = need;
        return KRB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};
struct FrontLaunch { int first, count; dim3 block; size_t smem; bool warp; int slot_elems; int stage_fwd = 0, stage_bwd = 0; int ll = 0, rpt = 1; };   // ll: warps of krb_front_factor_ll, 0 = another kernel
}  // namespace

struct FrontDevice {
    bool uploaded = false;
    int device = -1;          // device of the tables and buffers (a plan stays on the device of its first solve)
    FrontDesc *fronts = nullptr;
    int *list = nullptr, *idx_unknown = nullptr, *rel = nullptr, *child = nullptr, *asm_ent = nullptr, *asm_pos = nullptr, *col_order = nullptr;
    FrontElem *elems = nullptr;
    FrontTerm *mat_terms = nullptr, *rhs_terms = nullptr;
    int *mat_ptr = nullptr, *rhs_ptr = nullptr, *state_elems = nullptr;
    FrontChild *childrec = nullptr;
    int *res_ptr = nullptr, *res_ent = nullptr, *res_col = nullptr;
    int refine_steps = 0;
    bool verify = false;      // krb_plan_opts.refine_steps > 0: backward-error check after the refinement sets / clears KRB_LANE_INACCURATE
    std::vector<std::vector<FrontLaunch>> factor_launches, solve_launches;   // per level
    long long fw_size = 0;
    int smem_cap = 0;
    FBuf IN, EV, A, LU, LU0, CB, FW, CBV, BP, ST, XS, status, Bstep, IPIV, DX, ERR, SEL, SELIDX;
    std::vector<double> bstep_host;
    int bstep_rows = -1;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    size_t events_used = 0;
};

static double front_sum_event_ms(void *obj) {
    FrontDevice *d = (FrontDevice *)obj;
    if (!d || d->events_used == 0) return -1.0;
    double total = 0.0;
    for (size_t i = 0; i < d->events_used; i++) {
        if (cudaEventSynchronize(d->events[i].second) != cudaSuccess) return -1.0;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, d->events[i].first, d->events[i].second) != cudaSuccess) return -1.0;
        total += ms;
    }
    return total;
}

void front_device_free(FrontDevice *d) {
    if (!d) return;
    retire_last_timer(d);
    void *ptrs[] = {d->fronts, d->list, d->idx_unknown, d->rel, d->child, d->asm_ent, d->asm_pos, d->col_order, d->elems,
                    d->mat_terms, d->rhs_terms, d->mat_ptr, d->rhs_ptr, d->state_elems, d->res_ptr, d->res_ent, d->res_col, d->childrec};
    for (void *p : ptrs) if (p) cudaFree(p);
    FBuf *bufs[] = {&d->IN, &d->EV, &d->A, &d->LU, &d->CB, &d->FW, &d->CBV, &d->BP, &d->ST, &d->XS, &d->status, &d->Bstep, &d->IPIV, &d->DX, &d->LU0, &d->ERR, &d->SEL, &d->SELIDX};
    for (FBuf *b : bufs) b->release();
    for (auto &e : d->events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    delete d;
}

template <typename V>
static int up(const std::vector<V> &v, V **out) {
    CUDA_TRY(cudaMalloc((void **)out, std::max<size_t>(v.size(), 1) * sizeof(V)));
    if (!v.empty()) CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
    return KRB_OK;
}

static int upload_front(Plan &plan) {
    if (!plan.fdev) plan.fdev = new FrontDevice();
    FrontDevice &D = *plan.fdev;
    if (D.uploaded) return KRB_OK;
    const FrontPlan &fp = plan.front;
    const size_t ts = plan.is_complex ? 16 : 8;
    int dev = 0, optin = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    D.smem_cap = std::min(optin, 224 * 1024);
    int m_smem = 1;
    while ((size_t)(m_smem + 1) * ((m_smem + 1) | 1) * ts <= (size_t)D.smem_cap) m_smem++;
    if (const char *e = getenv("KRB_FRONT_SMEM_M")) m_smem = std::max(1, std::min(m_smem, atoi(e)));
    const int nf = fp.n_fronts;
    const bool use_ll = !(getenv("KRB_FRONT_LL") && getenv("KRB_FRONT_LL")[0] == '0');
    const bool use_warpf = !(getenv("KRB_FRONT_WARPF") && getenv("KRB_FRONT_WARPF")[0] == '0');
    int ll_max = 512;         // largest front of the left-looking kernel (one row per thread); KRB_FRONT_LL_MAX_M lowers it (tests: mixed routing)
    if (const char *e = getenv("KRB_FRONT_LL_MAX_M")) ll_max = std::max(32, std::min(512, atoi(e)));
    // fronts the left-looking kernel takes: one row per thread up to 512 rows; real plans up to 1024 rows with two rows per thread when
    // the pivot block has at most 512 (the second rows are never pivot candidates)
    const bool cplx_plan = plan.is_complex;
    auto ll_ok = [&](int m, int p) { return use_ll && (m <= ll_max || (ll_max == 512 && !cplx_plan && m <= 1024 && p <= 512)); };
    std::vector<FrontDesc> desc(nf);
    D.fw_size = 0;
    auto msize = [&](int f) { return (fp.piv0[f + 1] - fp.piv0[f]) + (fp.idx_ptr[f + 1] - fp.idx_ptr[f]); };
    for (int f = 0; f < nf; f++) {
        FrontDesc &d = desc[f];
        d.p = fp.piv0[f + 1] - fp.piv0[f];
        d.q = fp.idx_ptr[f + 1] - fp.idx_ptr[f];
        d.piv0 = fp.piv0[f];
        d.idx_off = fp.idx_ptr[f];
        d.child_begin = fp.child_ptr[f]; d.child_end = fp.child_ptr[f + 1];
        d.asm_begin = fp.asm_ptr[f]; d.asm_end = fp.asm_ptr[f + 1];
        d.lu_off = fp.lu_off[f]; d.cb_off = fp.cb_off[f];
        d.cbv_off = (int)fp.cbv_off[f]; d.pad = 0;
        const long long m = d.p + d.q;
        // fronts beyond shared memory: left-looking by resident column blocks (krb_front_factor_ll, no workspace) up to 512 rows,
        // above that the blocked right-looking kernel on a per-lane workspace; KRB_FRONT_LL=0 (A/B) sends all of them there
        if (m > m_smem && !ll_ok((int)m, d.p)) { d.fw_off = D.fw_size; D.fw_size += ((m * m + 1) / 2) * 2; }   // (with use_ll only fronts beyond the kernel's row limits)
        else d.fw_off = -1;
    }
    // launch lists: each level's fronts are sorted by size (largest first); cut them into size classes
    D.factor_launches.assign(fp.n_levels, {});
    D.solve_launches.assign(fp.n_levels, {});
    for (int l = 0; l < fp.n_levels; l++) {
        const int b = fp.level_ptr[l], e = fp.level_ptr[l + 1];
        // size classes: -1 left-looking (beyond shared memory), 0 blocked on the workspace, 1..4 CTA per front in shared memory,
        // 5..7 one warp per front (m <= 32; KRB_FRONT_WARPF=0 (A/B) keeps them on classes 3 / 4)
        auto fclass = [&](int f) {
            const int m = msize(f), pf = fp.piv0[f + 1] - fp.piv0[f];
            if (use_warpf && m <= 32 && m <= m_smem) return m > 24 ? 5 : m > 16 ? 6 : 7;
            if (ll_ok(m, pf)) return m > 512 ? -4 : m > 256 ? -1 : m > 128 ? -2 : -3;   // left-looking kernel on 16 (two rows per thread) / 16 / 8 / 4 warps
            if (m > m_smem) return 0;
            return m > 96 ? 1 : m > 48 ? 2 : m > 16 ? 3 : 4;
        };
        auto sclass = [&](int m) { return m > 192 ? 0 : m > 64 ? 1 : 2; };
        for (int i = b; i < e;) {
            const int m0 = msize(fp.level_fronts[i]);
            const int c = fclass(fp.level_fronts[i]);
            int j = i;
            while (j < e && fclass(fp.level_fronts[j]) == c) j++;
            FrontLaunch L;
            L.first = i; L.count = j - i;
            L.block = c <= 1 ? dim3(32, 16) : c == 2 ? dim3(16, 16) : c == 3 ? dim3(16, 8) : dim3(8, 8);
            L.smem = c == 0 ? std::min<size_t>((size_t)D.smem_cap, ((size_t)2 * m0 * 32 + m0) * ts) : (size_t)m0 * (m0 | 1) * ts;
            L.warp = false; L.slot_elems = 0;
            if (c < 0) {
                // shared memory by residency: 4 warps -> 56 KB (four CTAs per SM), 8 -> 112 KB (two), 16 -> everything; doubled while the
                // column blocks would be narrower than 8
                L.ll = (c == -1 || c == -4) ? 16 : c == -2 ? 8 : 4;
                L.rpt = c == -4 ? 2 : 1;
                size_t sm_b = (c == -1 || c == -4) ? (size_t)D.smem_cap : c == -2 ? (size_t)112 * 1024 : (size_t)56 * 1024;
                while (sm_b < (size_t)D.smem_cap && front_ll_width(m0, front_ll_ldc(m0), (long long)(sm_b / ts) - FRONT_LL_AUX) < 8) sm_b *= 2;
                L.smem = std::min(sm_b, (size_t)D.smem_cap);
                L.block = dim3(32, L.ll);
                for (int k = i; k < j; k++) {      // per-front column-block width for this launch's shared memory
                    const int f = fp.level_fronts[k], mf = msize(f);
                    desc[f].pad = front_ll_width(mf, front_ll_ldc(mf), (long long)(L.smem / ts) - FRONT_LL_AUX);
                }
            }
            if (c >= 5) {   // warp per (front, lane): 8 pairs per CTA while four CTAs still fit an SM, else 4
                L.warp = true;
                L.slot_elems = m0 * (m0 | 1);
                const int wpb = (size_t)L.slot_elems * ts * 8 <= (size_t)56 * 1024 ? 8 : 4;
                L.block = dim3(32, wpb);
                L.smem = (size_t)L.slot_elems * ts * wpb;
            }
            D.factor_launches[l].push_back(L);
            i = j;
        }
        for (int i = b; i < e;) {
            const int m0 = msize(fp.level_fronts[i]);
            const int c = sclass(m0);
            int j = i;
            while (j < e && sclass(msize(fp.level_fronts[j])) == c) j++;
            FrontLaunch L;
            L.first = i; L.count = j - i;
            const int nt = c == 0 ? 256 : c == 1 ? 128 : 32;
            L.warp = (c == 2);
            L.slot_elems = m0 + nt;
            L.block = L.warp ? dim3(32, 16) : dim3(nt);      // WARP mode: 16 independent (front, lane) pairs per CTA, 64 warps per SM
            L.smem = (size_t)L.slot_elems * ts * (L.warp ? 16 : 1);
            // staging of the factor block in shared memory (front_kernels.cuh: front_stage): largest block of the launch, when 16
            // (WARP mode: at least 8) slots of it still fit a CTA; KRB_FRONT_NO_STAGE=1 switches it off (A/B)
            if (getenv("KRB_FRONT_NO_STAGE") == nullptr) {
                long long sf = 0, sb = 0;
                for (int k = i; k < j; k++) {
                    const int f = fp.level_fronts[k];
                    const long long pp = fp.piv0[f + 1] - fp.piv0[f], qq = fp.idx_ptr[f + 1] - fp.idx_ptr[f];
                    sf = std::max(sf, (pp + qq) * pp);
                    sb = std::max(sb, (pp + qq) * pp + pp * qq);
                }
                const size_t budget = L.warp ? (size_t)96 * 1024 : (size_t)72 * 1024;     // per CTA
                auto fits = [&](long long e, int slots) { return ((size_t)L.slot_elems + (size_t)e) * ts * slots <= budget; };
                if (L.warp) {
                    // fewer warps per CTA when the blocks are large: 16, 8 or 4 slots
                    int slots = 16;
                    while (slots > 4 && !fits(sb, slots)) slots /= 2;
                    if (fits(sb, slots)) { L.block = dim3(32, slots); L.stage_fwd = (int)sf; L.stage_bwd = (int)sb; }
                } else if (fits(sb, 1)) { L.stage_fwd = (int)sf; L.stage_bwd = (int)sb; }
            }
            D.solve_launches[l].push_back(L);
            i = j;
        }
    }
    int rc;
    if ((rc = up(desc, &D.fronts))) return rc;
    if ((rc = up(fp.level_fronts, &D.list))) return rc;
    std::vector<int> idx_unknown(fp.idx.size());
    for (size_t i = 0; i < fp.idx.size(); i++) idx_unknown[i] = fp.col_order[fp.idx[i]];
    if ((rc = up(idx_unknown, &D.idx_unknown))) return rc;
    if ((rc = up(fp.rel, &D.rel))) return rc;
    if ((rc = up(fp.child, &D.child))) return rc;
    {
        std::vector<FrontChild> cr(fp.child.size());
        for (size_t i = 0; i < fp.child.size(); i++) {
            const int c = fp.child[i];
            cr[i].q = fp.idx_ptr[c + 1] - fp.idx_ptr[c]; cr[i].idx_off = fp.idx_ptr[c]; cr[i].cbv_off = (int)fp.cbv_off[c]; cr[i].pad = 0;
            cr[i].cb_off = fp.cb_off[c]; cr[i].pad2 = 0;
        }
        if ((rc = up(cr, &D.childrec))) return rc;
    }
    if ((rc = up(fp.asm_ent, &D.asm_ent))) return rc;
    if ((rc = up(fp.asm_pos, &D.asm_pos))) return rc;
    if ((rc = up(fp.col_order, &D.col_order))) return rc;
    // element table and stamp terms
    const Circuit &c = plan.circuit;
    std::vector<FrontElem> elems(c.el.size());
    for (size_t e = 0; e < c.el.size(); e++) {
        FrontElem &fe = elems[e];
        fe.value = c.el[e].value; fe.ac_amp = c.el[e].ac_amplitude;
        fe.ip = c.el[e].ip; fe.im = c.el[e].im; fe.ic = c.el[e].ic;
        fe.in_plane = plan.front_input_of_elem.empty() ? -1 : plan.front_input_of_elem[e];
        fe.wave_col = -1; fe.state_idx = -1;
    }
    for (size_t k = 0; k < plan.wave_elems.size(); k++) elems[plan.wave_elems[k]].wave_col = (int)k;
    for (size_t j = 0; j < plan.state_elems.size(); j++) elems[plan.state_elems[j]].state_idx = (int)j;
    if ((rc = up(elems, &D.elems))) return rc;
    std::vector<FrontTerm> mt, rt;
    std::vector<int> mp(plan.pattern.nnz + 1, 0), rp(plan.circuit.n + 1, 0);
    for (int q = 0; q < plan.pattern.nnz; q++) {
        for (int ci : plan.pattern.contribs[q]) { const Src &s = plan.stamps.mat[ci].s; mt.push_back({s.elem, (int)s.kind | (s.neg ? 0x100 : 0)}); }
        mp[q + 1] = (int)mt.size();
    }
    for (int k = 0; k < plan.circuit.n; k++) {
        for (int ri : plan.pattern.rhs_contribs[fp.row_order[k]]) { const Src &s = plan.stamps.rhs[ri].s; rt.push_back({s.elem, (int)s.kind | (s.neg ? 0x100 : 0)}); }
        rp[k + 1] = (int)rt.size();
    }
    if ((rc = up(mt, &D.mat_terms))) return rc;
    if ((rc = up(rt, &D.rhs_terms))) return rc;
    if ((rc = up(mp, &D.mat_ptr))) return rc;
    if ((rc = up(rp, &D.rhs_ptr))) return rc;
    if ((rc = up(plan.state_elems, &D.state_elems))) return rc;
    {   // rows of the permuted matrix for the refinement residual
        const Pattern &pt = plan.pattern;
        std::vector<int> pos_of_row(pt.n), rptr(pt.n + 1, 0), rent(pt.nnz), rcol(pt.nnz);
        for (int k = 0; k < pt.n; k++) pos_of_row[fp.row_order[k]] = k;
        for (int e = 0; e < pt.nnz; e++) rptr[pos_of_row[pt.rowidx[e]] + 1]++;
        for (int k = 0; k < pt.n; k++) rptr[k + 1] += rptr[k];
        std::vector<int> fill(rptr.begin(), rptr.end() - 1);
        for (int col = 0; col < pt.n; col++)
            for (int e = pt.colptr[col]; e < pt.colptr[col + 1]; e++) { const int t = fill[pos_of_row[pt.rowidx[e]]]++; rent[t] = e; rcol[t] = col; }
        if ((rc = up(rptr, &D.res_ptr))) return rc;
        if ((rc = up(rent, &D.res_ent))) return rc;
        if ((rc = up(rcol, &D.res_col))) return rc;
    }
    // AC plans pick static pivots at one nominal frequency but serve lanes decades away from it: two refinement steps on the
    // stamped matrix restore full accuracy there (each costs one residual + one pair of triangular solves).  Transient and DC
    // plans are solved at the point they were planned for.
    D.refine_steps = plan.mode == KRB_MODE_AC ? 2 : 0;
    if (plan.opts.refine_steps > 0) { D.refine_steps = plan.opts.refine_steps; D.verify = true; }   // the host mirror's pivoting fallback
    if (const char *e = getenv("KRB_FRONT_REFINE")) D.refine_steps = std::max(0, atoi(e));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor_ll<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor_ll<false, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor_ll<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor_ll<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor_ll<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor_ll<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_factor_ll<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_forward<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_forward<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_backward<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_backward<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_forward<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_forward<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_backward<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    CUDA_TRY(cudaFuncSetAttribute(krb_front_backward<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D.smem_cap));
    D.uploaded = true;
    return KRB_OK;
}

namespace {
struct Bufs { void *A, *LU, *CB, *FW, *CBV, *BP; int *status, *ipiv; };
}

static FrontParams make_params(const Plan &plan, const Bufs &b, void *X, long long x_lane_stride, long long x_row_off) {
    const FrontDevice &D = *plan.fdev;
    const FrontPlan &fp = plan.front;
    FrontParams P{};
    P.fronts = D.fronts; P.list = D.list; P.idx_unknown = D.idx_unknown; P.rel = D.rel; P.child = D.child; P.childrec = D.childrec;
    P.asm_ent = D.asm_ent; P.asm_pos = D.asm_pos; P.col_order = D.col_order;
    P.A = b.A; P.LU = b.LU; P.CB = b.CB; P.FW = b.FW; P.CBV = b.CBV; P.BP = b.BP;
    P.a_stride = plan.pattern.nnz; P.lu_stride = fp.lu_size; P.cb_stride = fp.cb_size; P.fw_stride = D.fw_size; P.cbv_stride = fp.cbv_size;
    P.n = fp.n; P.status = b.status; P.IPIV = b.ipiv;
    P.use_dmma = getenv("KRB_FRONT_NO_DMMA") == nullptr ? 1 : 0;
    P.X = X; P.x_lane_stride = x_lane_stride; P.x_row_off = x_row_off;
    return P;
}

static int front_factor(const Plan &plan, FrontParams P, long long lanes, cudaStream_t st) {
    const FrontDevice &D = *plan.fdev;
    const size_t ts = plan.is_complex ? 16 : 8;
    for (size_t l = 0; l < D.factor_launches.size(); l++)
        for (const FrontLaunch &L : D.factor_launches[l]) {
            FrontParams Q = P;
            Q.list = D.list + L.first;
            Q.smem_elems = (int)(L.smem / ts);
            Q.lanes = lanes;
            dim3 grid((unsigned)L.count, (unsigned)lanes);
            const bool small = L.block.x * L.block.y <= 128;
            if (L.ll) {
                if (plan.is_complex) {
                    if (L.ll == 16) krb_front_factor_ll<true, 16><<<grid, L.block, L.smem, st>>>(Q);
                    else if (L.ll == 8) krb_front_factor_ll<true, 8><<<grid, L.block, L.smem, st>>>(Q);
                    else krb_front_factor_ll<true, 4><<<grid, L.block, L.smem, st>>>(Q);
                } else {
                    if (L.ll == 16 && L.rpt == 2) krb_front_factor_ll<false, 16, 2><<<grid, L.block, L.smem, st>>>(Q);
                    else if (L.ll == 16) krb_front_factor_ll<false, 16><<<grid, L.block, L.smem, st>>>(Q);
                    else if (L.ll == 8) krb_front_factor_ll<false, 8><<<grid, L.block, L.smem, st>>>(Q);
                    else krb_front_factor_ll<false, 4><<<grid, L.block, L.smem, st>>>(Q);
                }
            } else if (L.warp) {
                Q.smem_elems = L.slot_elems;
                dim3 gw((unsigned)L.count, (unsigned)((lanes + L.block.y - 1) / L.block.y));
                if (plan.is_complex) krb_front_factor_warp<true><<<gw, L.block, L.smem, st>>>(Q);
                else krb_front_factor_warp<false><<<gw, L.block, L.smem, st>>>(Q);
            } else if (plan.is_complex) { if (small) krb_front_factor<true, 2><<<grid, L.block, L.smem, st>>>(Q); else krb_front_factor<true, 4><<<grid, L.block, L.smem, st>>>(Q); }
            else { if (small) krb_front_factor<false, 2><<<grid, L.block, L.smem, st>>>(Q); else krb_front_factor<false, 4><<<grid, L.block, L.smem, st>>>(Q); }
            add_launches(1);
        }
    CUDA_TRY(cudaGetLastError());
    return KRB_OK;
}

template <bool FWD>
static void launch_solve(const Plan &plan, const FrontLaunch &L, FrontParams Q, long long lanes, cudaStream_t st) {
    const size_t ts = plan.is_complex ? 16 : 8;
    Q.lanes = lanes;
    const int stage = FWD ? L.stage_fwd : L.stage_bwd;
    Q.stage_off = stage > 0 ? ((L.slot_elems + 1) & ~1) : 0;             // 16-byte aligned for the complex copies
    const int slot = stage > 0 ? ((Q.stage_off + stage + 1) & ~1) : L.slot_elems;   // even: every warp's slot stays 16-byte aligned
    if (L.warp) {
        Q.smem_elems = slot;
        const size_t smem = (size_t)slot * ts * L.block.y;
        dim3 grid((unsigned)L.count, (unsigned)((lanes + L.block.y - 1) / L.block.y));
        if (plan.is_complex) { if (FWD) krb_front_forward<true, true><<<grid, L.block, smem, st>>>(Q); else krb_front_backward<true, true><<<grid, L.block, smem, st>>>(Q); }
        else { if (FWD) krb_front_forward<false, true><<<grid, L.block, smem, st>>>(Q); else krb_front_backward<false, true><<<grid, L.block, smem, st>>>(Q); }
    } else {
        const size_t smem = (size_t)slot * ts;
        Q.smem_elems = slot;
        dim3 grid((unsigned)L.count, (unsigned)lanes);
        if (plan.is_complex) { if (FWD) krb_front_forward<true, false><<<grid, L.block, smem, st>>>(Q); else krb_front_backward<true, false><<<grid, L.block, smem, st>>>(Q); }
        else { if (FWD) krb_front_forward<false, false><<<grid, L.block, smem, st>>>(Q); else krb_front_backward<false, false><<<grid, L.block, smem, st>>>(Q); }
    }
    add_launches(1);
}

static int front_solve(const Plan &plan, FrontParams P, long long lanes, cudaStream_t st) {
    const FrontDevice &D = *plan.fdev;
    const int nl = (int)D.solve_launches.size();
    for (int l = 0; l < nl; l++)
        for (const FrontLaunch &L : D.solve_launches[l]) {
            FrontParams Q = P;
            Q.list = D.list + L.first;
            launch_solve<true>(plan, L, Q, lanes, st);
        }
    for (int l = nl - 1; l >= 0; l--)
        for (const FrontLaunch &L : D.solve_launches[l]) {
            FrontParams Q = P;
            Q.list = D.list + L.first;
            launch_solve<false>(plan, L, Q, lanes, st);
        }
    CUDA_TRY(cudaGetLastError());
    return KRB_OK;
}

int run_front(Plan *plan_first, Plan *plan_steps, const RunArgs &a) {
    if (device_count() <= 0) { set_error("no CUDA device: krets_b200 has no CPU fallback"); return KRB_ERR_NO_DEVICE; }
    Plan &P0 = *plan_first;
    Plan *PS = plan_steps;
    if (PS && !PS->use_front) { set_error("transient: plan_dc uses the front executor but plan_tran does not (analyse both with KRB_EXEC_STAGED)"); return KRB_ERR_UNSUPPORTED; }
    if (a.layout != KRB_LAYOUT_POINT_MAJOR) { set_error("the front executor delivers point-major results only"); return KRB_ERR_UNSUPPORTED; }
    if (a.op_solution && P0.circuit.nonlinear) { set_error("the front executor handles linear circuits only"); return KRB_ERR_UNSUPPORTED; }
    int rc;
    if ((rc = upload_front(P0))) return rc;
    if (PS && (rc = upload_front(*PS))) return rc;
    if (PS && PS->n_inputs != P0.n_inputs) { set_error("plan_dc and plan_tran must be analysed with the same varied elements"); return KRB_ERR_INVALID_ARGUMENT; }
    FrontDevice &D = *P0.fdev;
    {
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        if (D.device < 0) D.device = dev;
        if (PS && PS->fdev->device < 0) PS->fdev->device = dev;
        if (D.device != dev || (PS && PS->fdev->device != dev)) {
            set_error("this plan was first used on another CUDA device; analyse one plan per device");
            return KRB_ERR_INVALID_ARGUMENT;
        }
    }
    cudaStream_t st = (cudaStream_t)a.stream;
    const bool cplx = P0.is_complex;
    const size_t ts = cplx ? 16 : 8;
    const int n = P0.circuit.n, rows = a.rows;
    const int n_elems = (int)P0.circuit.el.size();
    const int n_state = PS ? (int)PS->state_elems.size() : 0;
    if (n == 0 || a.lanes <= 0) return KRB_OK;
    const bool out_dev = (a.memory == KRB_MEM_DEVICE);
    const int nsel = a.n_select;               // extension (krb_io.select): full rows are solved into XS, a gather compacts them
    const bool sel = nsel > 0;
    const bool staged = !out_dev || sel;

    if (PS) {   // per-step source table, as in the lane-VM runtime (the reference evaluates sources one step late)
        FrontDevice &DS = *PS->fdev;
        if (DS.bstep_rows != rows) {
            const int nw = (int)PS->wave_elems.size();
            const double h = PS->opts.time_step;
            DS.bstep_host.assign((size_t)rows * 2 * std::max(nw, 1), 0.0);
            for (int r = 1; r < rows; r++)
                for (int k = 0; k < nw; k++) {
                    const krb_element &el = PS->circuit.el[PS->wave_elems[k]];
                    DS.bstep_host[((size_t)r * 2 + 0) * nw + k] = transient_value_at(el, (double)(PS->opts.integrator == KRB_INTEGRATOR_BE_KRETS ? r - 1 : r) * h);
                    DS.bstep_host[((size_t)r * 2 + 1) * nw + k] = transient_value_at(el, (double)r * h);
                }
            if ((rc = DS.Bstep.ensure(DS.bstep_host.size() * 8))) return rc;
            CUDA_TRY(cudaMemcpy(DS.Bstep.p, DS.bstep_host.data(), DS.bstep_host.size() * 8, cudaMemcpyHostToDevice));
            DS.bstep_rows = rows;
        }
    }

    // ---- chunking by device memory ----
    // persistent per lane of a chunk: inputs, element values, stamped values A, the factor store of the plan that is solved
    // repeatedly (LU), vectors; workspace per lane of a factor window: contribution blocks, out-of-shared-memory fronts, and for a
    // transient the operating-point plan's own factor store (needed for row 0 only).
    Plan &PM = PS ? *PS : P0;                      // plan whose factors stay resident
    const size_t a_e = (size_t)std::max<long long>(P0.pattern.nnz, PS ? PS->pattern.nnz : 0);
    const size_t cbv_e = (size_t)std::max<long long>(P0.front.cbv_size, PS ? PS->front.cbv_size : 0);
    const size_t lu_e = (size_t)PM.front.lu_size;
    const size_t lu0_e = PS ? (size_t)P0.front.lu_size : 0;
    const size_t cb_e = (size_t)std::max<long long>(P0.front.cb_size, PS ? PS->front.cb_size : 0);
    const size_t fw_e = (size_t)std::max<long long>(P0.fdev->fw_size, PS ? PS->fdev->fw_size : 0);
    const bool refine = P0.fdev->refine_steps > 0 || (PS && PS->fdev->refine_steps > 0);
    const bool verify = P0.fdev->verify || (PS && PS->fdev->verify);
    const size_t persist = (lu_e + a_e + cbv_e + (size_t)n * (refine ? 2 : 1)) * ts + (size_t)n_elems * 8 + (size_t)P0.n_inputs * 8 + (size_t)n_state * 8 +
                           (size_t)n * 4 + (staged ? (size_t)rows * n * ts : 0) + (sel && !out_dev ? (size_t)rows * nsel * ts : 0) + 64;
    const size_t work = (cb_e + fw_e + lu0_e) * ts + 64;
    long long chunk = a.chunk_lanes > 0 ? a.chunk_lanes : 2048;
    if (const char *e = getenv("KRB_FRONT_CHUNK")) chunk = std::max(1, atoi(e));   // tests: force several chunks
    chunk = std::min<long long>(chunk, a.lanes);
    chunk = std::min<long long>(chunk, 65535);
    long long window = 256;
    {   // factor windows of whole waves: the top tree levels launch one CTA per (front, lane) at one CTA per SM
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) window = 2 * sms;
    }
    if (const char *e = getenv("KRB_FRONT_WINDOW")) window = std::max(1, atoi(e));
    window = std::min(window, chunk);
    {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        size_t reusable = 0;
        FBuf *all[] = {&D.IN, &D.EV, &D.A, &D.LU, &D.LU0, &D.CB, &D.FW, &D.CBV, &D.BP, &D.ST, &D.XS, &D.IPIV, &D.DX};
        for (FBuf *bf : all) reusable += bf->bytes;
        const double avail = (double)(free_b + reusable) * 0.88;
        if (avail < (double)(persist + work)) { set_error("not enough device memory for one lane of this plan"); return KRB_ERR_OUT_OF_MEMORY; }
        window = std::min<long long>(window, std::max<long long>(1, (long long)(avail * 0.5 / (double)work)));
        const long long max_chunk = (long long)((avail - (double)window * (double)work) / (double)persist);
        chunk = std::max<long long>(1, std::min(chunk, max_chunk));
        window = std::min(window, chunk);
    }
    if ((rc = D.IN.ensure(std::max<size_t>((size_t)P0.n_inputs, 1) * chunk * 8))) return rc;
    if ((rc = D.EV.ensure(std::max<size_t>((size_t)n_elems, 1) * chunk * 8))) return rc;
    if ((rc = D.A.ensure(a_e * chunk * ts))) return rc;
    if ((rc = D.LU.ensure(lu_e * chunk * ts))) return rc;
    if (PS && (rc = D.LU0.ensure(lu0_e * window * ts))) return rc;
    if ((rc = D.CB.ensure(std::max<size_t>(cb_e, 1) * window * ts))) return rc;
    if ((rc = D.FW.ensure(std::max<size_t>(fw_e, 1) * window * ts))) return rc;
    if ((rc = D.CBV.ensure(std::max<size_t>(cbv_e, 1) * chunk * ts))) return rc;
    if ((rc = D.BP.ensure((size_t)n * chunk * ts))) return rc;
    if ((rc = D.ST.ensure(std::max<size_t>((size_t)n_state, 1) * chunk * 8))) return rc;
    if ((rc = D.status.ensure((size_t)chunk * 4))) return rc;
    if ((rc = D.IPIV.ensure((size_t)n * chunk * 4))) return rc;
    if (refine && (rc = D.DX.ensure((size_t)n * chunk * ts))) return rc;
    if (verify && (rc = D.ERR.ensure((size_t)chunk * 8))) return rc;
    if (staged && (rc = D.XS.ensure((size_t)rows * n * chunk * ts))) return rc;
    if (sel && !out_dev && (rc = D.SEL.ensure((size_t)rows * nsel * chunk * ts))) return rc;
    if (sel) {
        if ((rc = D.SELIDX.ensure((size_t)nsel * 4))) return rc;
        CUDA_TRY(cudaMemcpyAsync(D.SELIDX.p, a.select, (size_t)nsel * 4, cudaMemcpyHostToDevice, st));
    }

    D.events_used = 0;
    set_last_timer(front_sum_event_ms, &D);
    const long long xls = (long long)rows * n;   // lane stride of the result rows
    for (long long lane0 = 0; lane0 < a.lanes; lane0 += chunk) {
        const long long lanes = std::min<long long>(chunk, a.lanes - lane0);
        double *IN = (double *)D.IN.p;
        {   // input planes IN[k][lanes]: the sweep value / frequency plane, then (normally) one contiguous run of Monte-Carlo planes
            const cudaMemcpyKind kind = out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
            const int nv = (int)P0.input_of_varied.size();
            if (nv > 0 && !a.varied_values) { set_error("missing per-lane values for a varied element"); return KRB_ERR_INVALID_ARGUMENT; }
            if (P0.input_sweep >= 0) CUDA_TRY(cudaMemcpyAsync(IN + (size_t)P0.input_sweep * lanes, a.lane_param + lane0, (size_t)lanes * 8, kind, st));
            if (P0.input_freq >= 0) CUDA_TRY(cudaMemcpyAsync(IN + (size_t)P0.input_freq * lanes, a.lane_param + lane0, (size_t)lanes * 8, kind, st));
            bool run = nv > 0;
            for (int q = 1; q < nv && run; q++) run = P0.input_of_varied[q] == P0.input_of_varied[0] + q;
            for (int q = 0; q < nv && run; q++) run = P0.input_of_varied[q] != P0.input_sweep;   // the swept source's plane is the sweep value (as runtime.cu)
            if (run)
                CUDA_TRY(cudaMemcpy2DAsync(IN + (size_t)P0.input_of_varied[0] * lanes, (size_t)lanes * 8, a.varied_values + lane0, (size_t)a.lanes * 8,
                                           (size_t)lanes * 8, (size_t)nv, kind, st));
            else
                for (int q = 0; q < nv; q++) {
                    const int k = P0.input_of_varied[q];
                    if (k == P0.input_sweep) continue;   // the swept source's plane is the sweep value
                    CUDA_TRY(cudaMemcpyAsync(IN + (size_t)k * lanes, a.varied_values + (size_t)q * a.lanes + lane0, (size_t)lanes * 8, kind, st));
                }
        }
        CUDA_TRY(cudaMemsetAsync(D.status.p, 0, (size_t)lanes * 4, st));
        if (n_state) CUDA_TRY(cudaMemsetAsync(D.ST.p, 0, (size_t)n_state * lanes * 8, st));   // i_C(0) = 0
        char *X = staged ? (char *)D.XS.p : (char *)a.x_out + (size_t)lane0 * xls * ts;

        if (D.events_used == D.events.size()) {
            cudaEvent_t x, y;
            CUDA_TRY(cudaEventCreate(&x));
            CUDA_TRY(cudaEventCreate(&y));
            D.events.push_back({x, y});
        }
        cudaEvent_t e0 = D.events[D.events_used].first, e1 = D.events[D.events_used].second;
        D.events_used++;
        CUDA_TRY(cudaEventRecord(e0, st));

        // a lane window [s0, s0 + sl) of the chunk: persistent stores are addressed from s0, workspace from 0
        auto bufs_for = [&](const Plan &pl, void *lu_base, bool lu_windowed, long long s0) {
            Bufs b{};
            b.A = (char *)D.A.p + (size_t)s0 * pl.pattern.nnz * ts;
            b.LU = (char *)lu_base + (lu_windowed ? 0 : (size_t)s0 * pl.front.lu_size * ts);
            b.CB = D.CB.p; b.FW = D.FW.p;
            b.CBV = (char *)D.CBV.p + (size_t)s0 * pl.front.cbv_size * ts;
            b.BP = (char *)D.BP.p + (size_t)s0 * n * ts;
            b.status = (int *)D.status.p + s0;
            b.ipiv = (int *)D.IPIV.p + (size_t)s0 * n;
            return b;
        };
        auto stamp_params = [&](const Plan &pl, long long s0) {
            const FrontDevice &DD = *pl.fdev;
            FrontStampParams s{};
            s.elems = DD.elems; s.n_elems = n_elems; s.IN = IN + s0; s.L = lanes; s.EV = (double *)D.EV.p + (size_t)s0 * n_elems;
            s.input_freq = pl.input_freq; s.h_comp = pl.h_comp; s.x_lane_stride = xls;
            s.ST = (double *)D.ST.p + (size_t)s0 * n_state; s.n_state = n_state;
            return s;
        };
        {   // per-lane element values
            FrontStampParams s = stamp_params(P0, 0);
            krb_front_expand<<<dim3((unsigned)((n_elems + 255) / 256), (unsigned)lanes), 256, 0, st>>>(s);
            add_launches(1);
        }
        auto stamp_and_factor = [&](const Plan &pl, void *lu_base, bool lu_windowed, long long s0, long long sl) -> int {
            NvtxRange nvtx("krb front: stamp + multifrontal factor (one lane window)");
            FrontStampParams s = stamp_params(pl, s0);
            s.terms = pl.fdev->mat_terms; s.term_ptr = pl.fdev->mat_ptr; s.count = pl.pattern.nnz;
            s.OUT = (char *)D.A.p + (size_t)s0 * pl.pattern.nnz * ts;
            dim3 grid((unsigned)((s.count + 255) / 256), (unsigned)sl);
            if (pl.is_complex) krb_front_stamp<true><<<grid, 256, 0, st>>>(s);
            else krb_front_stamp<false><<<grid, 256, 0, st>>>(s);
            add_launches(1);
            return front_factor(pl, make_params(pl, bufs_for(pl, lu_base, lu_windowed, s0), X + (size_t)s0 * xls * ts, xls, 0), sl, st);
        };
        auto rhs_and_solve = [&](const Plan &pl, void *lu_base, bool lu_windowed, long long s0, long long sl, int row) -> int {
            NvtxRange nvtx("krb front: rhs + level-scheduled solves (+ refinement)");
            char *Xw = X + (size_t)s0 * xls * ts;
            const Bufs bufs = bufs_for(pl, lu_base, lu_windowed, s0);
            FrontStampParams s = stamp_params(pl, s0);
            s.terms = pl.fdev->rhs_terms; s.term_ptr = pl.fdev->rhs_ptr; s.count = n; s.OUT = bufs.BP;
            s.Bstep = (const double *)pl.fdev->Bstep.p; s.n_wave = (int)pl.wave_elems.size(); s.row = row;
            s.Xprev = row > 0 ? (const double *)Xw + (size_t)(row - 1) * n : nullptr;
            dim3 grid((unsigned)((n + 255) / 256), (unsigned)sl);
            if (pl.is_complex) krb_front_rhs<true><<<grid, 256, 0, st>>>(s);
            else krb_front_rhs<false><<<grid, 256, 0, st>>>(s);
            add_launches(1);
            int rc2 = front_solve(pl, make_params(pl, bufs, Xw, xls, (long long)row * n), sl, st);
            for (int it = 0; it < pl.fdev->refine_steps && !rc2; it++) {
                if (pl.is_complex) krb_front_rhs<true><<<grid, 256, 0, st>>>(s);
                else krb_front_rhs<false><<<grid, 256, 0, st>>>(s);
                FrontRefineParams r{};
                r.res_ptr = pl.fdev->res_ptr; r.res_ent = pl.fdev->res_ent; r.res_col = pl.fdev->res_col;
                r.A = bufs.A; r.a_stride = pl.pattern.nnz; r.BP = bufs.BP; r.X = Xw; r.x_lane_stride = xls; r.x_row_off = (long long)row * n;
                r.DX = (char *)D.DX.p + (size_t)s0 * n * ts; r.n = n;
                if (pl.is_complex) krb_front_residual<true><<<grid, 256, 0, st>>>(r);
                else krb_front_residual<false><<<grid, 256, 0, st>>>(r);
                rc2 = front_solve(pl, make_params(pl, bufs, r.DX, n, 0), sl, st);
                if (pl.is_complex) krb_front_axpy<true><<<grid, 256, 0, st>>>(r);
                else krb_front_axpy<false><<<grid, 256, 0, st>>>(r);
                add_launches(3);
            }
            if (pl.fdev->verify && !rc2) {
                // a-posteriori check of the delivered row: componentwise backward error <= 2^-40, else KRB_LANE_INACCURATE
                unsigned long long *err = (unsigned long long *)D.ERR.p + s0;
                CUDA_TRY(cudaMemsetAsync(err, 0, (size_t)sl * 8, st));
                if (pl.is_complex) krb_front_rhs<true><<<grid, 256, 0, st>>>(s);
                else krb_front_rhs<false><<<grid, 256, 0, st>>>(s);
                FrontRefineParams r{};
                r.res_ptr = pl.fdev->res_ptr; r.res_ent = pl.fdev->res_ent; r.res_col = pl.fdev->res_col;
                r.A = bufs.A; r.a_stride = pl.pattern.nnz; r.BP = bufs.BP; r.X = Xw; r.x_lane_stride = xls; r.x_row_off = (long long)row * n;
                r.DX = nullptr; r.n = n; r.err = err; r.status = bufs.status;
                if (pl.is_complex) krb_front_residual<true><<<grid, 256, 0, st>>>(r);
                else krb_front_residual<false><<<grid, 256, 0, st>>>(r);
                krb_front_verify<<<(unsigned)((sl + 255) / 256), 256, 0, st>>>(err, bufs.status, sl, 9.094947017729282e-13);
                add_launches(3);
            }
            return rc2;
        };
        // KRB_FRONT_DEBUG=1: per-phase device times of the first chunk on stderr (synchronises; profiling aid only)
        const bool dbg = getenv("KRB_FRONT_DEBUG") != nullptr && lane0 == 0;
        std::vector<std::pair<const char *, cudaEvent_t>> marks;
        auto mark = [&](const char *name) {
            if (!dbg) return;
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, st);
            marks.push_back({name, e});
        };
        mark("start");
        if (!PS) {
            // one-shot solves (DC sweep / OP lanes / AC): factor window by window, each window solved while its factors are hot
            for (long long s0 = 0; s0 < lanes; s0 += window) {
                const long long sl = std::min(window, lanes - s0);
                if ((rc = stamp_and_factor(P0, D.LU.p, false, s0, sl))) return rc;
                if ((rc = rhs_and_solve(P0, D.LU.p, false, s0, sl, 0))) return rc;
            }
            mark("stamp+factor+solve");
        } else {
            // transient: operating point (row 0) through the DC plan window by window, then the step plan's factors are built
            // window by window into the resident store, then every step solves the whole chunk at once
            for (long long s0 = 0; s0 < lanes; s0 += window) {
                const long long sl = std::min(window, lanes - s0);
                if ((rc = stamp_and_factor(P0, D.LU0.p, true, s0, sl))) return rc;
                if ((rc = rhs_and_solve(P0, D.LU0.p, true, s0, sl, 0))) return rc;
            }
            mark("operating point (stamp+factor+solve)");
            if (rows > 1) {
                for (long long s0 = 0; s0 < lanes; s0 += window) {
                    const long long sl = std::min(window, lanes - s0);
                    if ((rc = stamp_and_factor(*PS, D.LU.p, false, s0, sl))) return rc;
                }
                mark("stamp+factor (step plan)");
                for (int r = 1; r < rows; r++) {
                    if ((rc = rhs_and_solve(*PS, D.LU.p, false, 0, lanes, r))) return rc;
                    if (r <= 2) mark("rhs+solve step");
                    if (n_state) {
                        FrontTrapParams t{};
                        t.elems = PS->fdev->elems; t.state_elems = PS->fdev->state_elems; t.n_state = n_state; t.n_elems = n_elems;
                        t.EV = (const double *)D.EV.p; t.h_comp = PS->h_comp;
                        t.Xnew = (const double *)X + (size_t)r * n; t.Xprev = (const double *)X + (size_t)(r - 1) * n; t.x_lane_stride = xls;
                        t.ST = (double *)D.ST.p;
                        krb_front_trap_state<<<dim3((unsigned)((n_state + 255) / 256), (unsigned)lanes), 256, 0, st>>>(t);
                        add_launches(1);
                    }
                }
            }
        }
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaGetLastError());
#ifdef KRB_LL_PROFILE
        if (getenv("KRB_LL_PROFILE")) {
            cudaStreamSynchronize(st);
            unsigned long long h[24] = {0};
            cudaMemcpyFromSymbol(h, krb_ll_prof, sizeof(h));
            const char *nm[8] = {"assemble+swaps", "stage L_KK", "U_KJ solve", "DMMA update", "pivot columns", "sub-panel U rows", "in-block DMMA", "write-out"};
            for (int g = 0; g < 3; g++) {
                unsigned long long tot = 0;
                for (int i = 0; i < 8; i++) tot += h[8 * g + i];
                for (int i = 0; i < 8; i++)
                    fprintf(stderr, "[krb ll %2d warps] %-18s %6.2f %%  %12llu cycles\n", g == 0 ? 16 : g == 1 ? 8 : 4, nm[i], 100.0 * (double)h[8 * g + i] / (double)(tot ? tot : 1), h[8 * g + i]);
            }
            unsigned long long z[24] = {0};
            cudaMemcpyToSymbol(krb_ll_prof, z, sizeof(z));
        }
#endif
        if (dbg) {
            mark("end");
            cudaStreamSynchronize(st);
            for (size_t i = 1; i < marks.size(); i++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
                fprintf(stderr, "[krb front] lanes %lld  %-28s %9.3f ms\n", lanes, marks[i].first, ms);
            }
            for (auto &mk : marks) cudaEventDestroy(mk.second);
        }
        // ---- delivery ----
        NvtxRange nvtx_deliver("krb front: deliver");
        const long long xls_out = (long long)rows * (sel ? nsel : n);
        if (sel) {
            void *dst = out_dev ? (void *)((char *)a.x_out + (size_t)lane0 * xls_out * ts) : D.SEL.p;
            const long long points = lanes * rows, work = points * nsel;
            if (cplx) krb_select_columns<double2><<<(unsigned)((work + 255) / 256), 256, 0, st>>>((const double2 *)D.XS.p, (double2 *)dst, (const int *)D.SELIDX.p, n, nsel, points);
            else krb_select_columns<double><<<(unsigned)((work + 255) / 256), 256, 0, st>>>((const double *)D.XS.p, (double *)dst, (const int *)D.SELIDX.p, n, nsel, points);
            add_launches(1);
            CUDA_TRY(cudaGetLastError());
        }
        if (!out_dev) {
            CUDA_TRY(cudaMemcpyAsync((char *)a.x_out + (size_t)lane0 * xls_out * ts, sel ? D.SEL.p : D.XS.p, (size_t)lanes * xls_out * ts, cudaMemcpyDeviceToHost, st));
            if (a.lane_status) CUDA_TRY(cudaMemcpyAsync(a.lane_status + lane0, D.status.p, (size_t)lanes * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        } else if (a.lane_status) {
            CUDA_TRY(cudaMemcpyAsync(a.lane_status + lane0, D.status.p, (size_t)lanes * 4, cudaMemcpyDeviceToDevice, st));
        }
        if (a.iterations) {
            // linear circuits: one Newton iteration (op.rs:33-82 with an empty first guess converges on the second pass only for
            // nonlinear elements; the reference returns after iteration 1 for linear circuits)
            std::vector<int> ones((size_t)lanes, 1);
            CUDA_TRY(cudaMemcpyAsync(a.iterations + lane0, ones.data(), (size_t)lanes * 4, out_dev ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
    }
    return KRB_OK;
}

}  // namespace krb
