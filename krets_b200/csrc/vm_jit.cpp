// Circuit-specialised lane kernels (see vm_jit.hpp): source generation, NVRTC compilation, module loading and launch.
// NVRTC and the driver API are bound with dlopen at first use, so the library itself links against neither; a box without
// NVRTC keeps the interpreter (runtime.cu reports why once on stderr).
#include "vm_jit.hpp"

#include <cuda.h>
#include <dlfcn.h>
#include <unistd.h>
#include <nvrtc.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace krb {

static const char *const PRELUDE =
#include "vm_jit_prelude.inc"
    ;

static const int64_t KRB_VM_JIT_SMALL_OPS = 2500;   // about a second of NVRTC: always specialised

int64_t vm_jit_max_ops() {
    static const int64_t v = [] {
        const char *e = getenv("KRB_VM_JIT_MAX_OPS");
        return e ? (int64_t)atoll(e) : (int64_t)20000;
    }();
    return v;
}
bool vm_jit_enabled() {
    const char *e = getenv("KRB_VM_JIT");
    return !(e && e[0] == '0');
}
bool vm_jit_eligible(const Plan &plan) {
    if (plan.use_chain || plan.use_front) return false;
    const int64_t total = plan.sec[0].n_ops + plan.sec[1].n_ops + plan.sec[2].n_ops;
    return total > 0 && total <= vm_jit_max_ops();
}

bool vm_jit_worthwhile(const Plan &plan, long long points) {
    const int64_t total = plan.sec[0].n_ops + plan.sec[1].n_ops + plan.sec[2].n_ops;
    static const long long min_points = [] { const char *e = getenv("KRB_VM_JIT_MIN_POINTS"); return e ? atoll(e) : 32768ll; }();
    return total <= KRB_VM_JIT_SMALL_OPS || points >= min_points;
}

// ---- source generation -----------------------------------------------------------------------------------------------------
static void put(std::string &o, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
static void put(std::string &o, const char *fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    o += buf;
}

// ---- decoded lane-program operations ----
struct DecOp { uint32_t opc, d, a, b, imm; uint64_t bits; };
static std::vector<DecOp> decode(const Program &prog) {
    std::vector<DecOp> v;
    const std::vector<uint32_t> &w = prog.ops;
    const size_t n = w.size() / 2;
    for (size_t i = 0; i < n; i++) {
        const uint32_t w0 = w[2 * i];
        DecOp o{w0 & 0xffu, (w0 >> 8) & 0xffu, (w0 >> 16) & 0xffu, w0 >> 24, w[2 * i + 1], 0};
        if (o.opc == KOP_END) break;
        if (o.opc == KOP_NOP || o.opc == KOP_WAIT) continue;   // loads are synchronous here: the compiler schedules them
        if (o.opc == KOP_LDK) {                                  // the constant sits in the next op slot (both words)
            if (i + 1 >= n) break;
            o.bits = (uint64_t)w[2 * i + 2] | ((uint64_t)w[2 * i + 3] << 32);
            i++;
        }
        v.push_back(o);
    }
    return v;
}
// scratch cells an operation reads (n_r of them in r[]) and the one it writes (-1: none)
static void op_slots(const DecOp &o, int r[3], int &n_r, int &w) {
    n_r = 0; w = -1;
    switch (o.opc) {
    case KOP_LDB: case KOP_LDBS: case KOP_LDI: case KOP_LDG: case KOP_LDX: case KOP_ZERO: case KOP_LDK: w = (int)o.d; break;
    case KOP_STG: case KOP_XOUT: r[n_r++] = (int)o.a; break;
    case KOP_MOV: case KOP_NEG: case KOP_RCP: case KOP_INV: case KOP_MULJ: case KOP_DIOG: case KOP_DIOI: r[n_r++] = (int)o.a; w = (int)o.d; break;
    case KOP_ADD: case KOP_SUB: case KOP_MUL: case KOP_MULG: case KOP_DIVR: case KOP_MOSGM: case KOP_MOSGDS: case KOP_MOSID:
        r[n_r++] = (int)o.a; r[n_r++] = (int)o.b; w = (int)o.d; break;
    case KOP_FNMA: r[n_r++] = (int)o.d; r[n_r++] = (int)o.a; r[n_r++] = (int)o.b; w = (int)o.d; break;
    default: break;
    }
}

// one operation = one statement (same order, same arithmetic as vm_kernel.cuh); scratch cell k is written as pre<k>post
static void emit_stmt(const DecOp &op, const char *pre, const char *post, std::string &o) {
    char D[24], A[24], Bn[24];
    snprintf(D, sizeof D, "%s%u%s", pre, op.d, post);
    snprintf(A, sizeof A, "%s%u%s", pre, op.a, post);
    snprintf(Bn, sizeof Bn, "%s%u%s", pre, op.b, post);
    const uint32_t imm = op.imm;
    switch (op.opc) {
    case KOP_LDB: put(o, "%s = k_real(c.Bc[%u]);\n", D, imm); break;
    case KOP_LDBS: put(o, "%s = k_real(c.Bs[%u]);\n", D, imm); break;
    case KOP_LDI: put(o, "%s = k_real(c.IN[(size_t)%uu * c.L]);\n", D, imm); break;
    case KOP_LDG: put(o, "%s = c.G[(size_t)%uu * c.L];\n", D, imm); break;
    case KOP_STG: put(o, "c.G[(size_t)%uu * c.L] = %s;\n", imm, A); break;
    case KOP_LDX: put(o, "%s = k_ldx(c, %uu, %uu);\n", D, imm, op.b); break;
    case KOP_XOUT: put(o, "k_xout(c, %s, %uu, %uu, status, notconv);\n", A, imm, op.b); break;
    case KOP_ZERO: put(o, "%s = k_zero();\n", D); break;
    case KOP_MOV: put(o, "%s = %s;\n", D, A); break;
    case KOP_NEG: put(o, "%s = k_neg(%s);\n", D, A); break;
    case KOP_ADD: put(o, "%s = k_add(%s, %s);\n", D, A, Bn); break;
    case KOP_SUB: put(o, "%s = k_sub(%s, %s);\n", D, A, Bn); break;
    case KOP_MUL: put(o, "%s = k_mul(%s, %s);\n", D, A, Bn); break;
    case KOP_FNMA: put(o, "%s = k_fnma(%s, %s, %s);\n", D, D, A, Bn); break;
    case KOP_RCP: put(o, "%s = k_rcp(%s, status);\n", D, A); break;
    case KOP_INV: put(o, "%s = k_inv(%s);\n", D, A); break;
    case KOP_DIVR: put(o, "%s = k_divr(%s, %s);\n", D, A, Bn); break;
    case KOP_MULJ: put(o, "%s = k_mulj(%s);\n", D, A); break;
    case KOP_DIOG: put(o, "%s = k_diog(%s, c.Bc + %u);\n", D, A, imm); break;
    case KOP_DIOI: put(o, "%s = k_dioi(%s, c.Bc + %u);\n", D, A, imm); break;
    case KOP_LDK: put(o, "%s = k_const(0x%016llxull);\n", D, (unsigned long long)op.bits); break;
    case KOP_MOSGM:
    case KOP_MOSGDS:
    case KOP_MOSID: put(o, "%s = k_mos(%d, %s, %s, c.Bc + %u);\n", D, (int)op.opc - (int)KOP_MOSGM, A, Bn, imm); break;
    case KOP_MULG: put(o, "%s = k_mulg(%s, %s, grow);\n", D, A, Bn); break;
    default: break;
    }
}

static int max_slot(const std::vector<DecOp> &ops) {
    int m = 0, r[3], n_r, w;
    for (const DecOp &op : ops) {
        op_slots(op, r, n_r, w);
        for (int k = 0; k < n_r; k++) m = std::max(m, r[k]);
        m = std::max(m, w);
    }
    return m;
}

// Long programs: NVVM's time grows faster than linearly with the length of one basic block, so a section of more than
// KRB_JIT_INLINE_OPS operations is cut into chunks of KRB_JIT_CHUNK_OPS, each a __noinline__ device function.  Inside a chunk the
// scratch cells are local variables (registers); at its boundaries only the cells some chunk reads before it writes them
// (the live-in set of any chunk of any section: sound across Newton iterations, time steps and pinned cells) travel through
// the per-thread array s[].
static const int KRB_JIT_INLINE_OPS = 1200;
static const int KRB_JIT_CHUNK_OPS = [] { const char *e = getenv("KRB_VM_JIT_CHUNK_OPS"); const int v = e ? atoi(e) : 320; return v > 8 ? v : 320; }();
struct Chunk { size_t begin, end; std::vector<int> live_in, written; };
static std::vector<Chunk> make_chunks(const std::vector<DecOp> &ops) {
    std::vector<Chunk> out;
    for (size_t b = 0; b < ops.size(); b += KRB_JIT_CHUNK_OPS) {
        Chunk ch;
        ch.begin = b; ch.end = std::min(ops.size(), b + (size_t)KRB_JIT_CHUNK_OPS);
        char wr[KRB_MAX_SLOTS] = {}, in[KRB_MAX_SLOTS] = {};
        for (size_t i = ch.begin; i < ch.end; i++) {
            int r[3], n_r, w;
            op_slots(ops[i], r, n_r, w);
            for (int k = 0; k < n_r; k++) if (!wr[r[k]]) in[r[k]] = 1;
            if (w >= 0) wr[w] = 1;
        }
        for (int s = 0; s < KRB_MAX_SLOTS; s++) { if (in[s]) ch.live_in.push_back(s); if (wr[s]) ch.written.push_back(s); }
        out.push_back(ch);
    }
    return out;
}

std::string vm_jit_source(const Plan &plan, size_t vm_params_size, bool host_mode) {
    std::string o;
    o.reserve(64 * 1024);
    std::vector<DecOp> ops[3];
    int ns = 0;
    size_t longest = 0;
    for (int k = 0; k < 3; k++) { ops[k] = decode(plan.sec[k]); ns = std::max(ns, max_slot(ops[k])); longest = std::max(longest, ops[k].size()); }
    const bool chunked = longest > (size_t)KRB_JIT_INLINE_OPS;
    if (host_mode) o += "#define KRB_JIT_HOST 1\n#include \"jit_host_shim.h\"\n";
    put(o, "#define KRB_JIT_CPLX %d\n", plan.is_complex ? 1 : 0);
    put(o, "#define KRB_LANE_ZERO_PIVOT %d\n#define KRB_LANE_NOT_CONVERGED %d\n#define KRB_LANE_NONFINITE %d\n#define KRB_LANE_INACCURATE %d\n",
        (int)KRB_LANE_ZERO_PIVOT, (int)KRB_LANE_NOT_CONVERGED, (int)KRB_LANE_NONFINITE, (int)KRB_LANE_INACCURATE);
    put(o, "#define KRB_GROWTH_LIMIT_LOG2 %d\n", (int)KRB_GROWTH_LIMIT_LOG2);
    put(o, "#define KRB_JIT_NSLOTS %d\n", ns + 1);
    o += PRELUDE;
    put(o, "static_assert(sizeof(VmParams) == %zu, \"VmParams layout differs from the library's\");\n", vm_params_size);
    std::vector<Chunk> chunks[3];
    if (chunked) {
        char any_in[KRB_MAX_SLOTS] = {};
        for (int k = 0; k < 3; k++) {
            chunks[k] = make_chunks(ops[k]);
            for (const Chunk &ch : chunks[k]) for (int sl : ch.live_in) any_in[sl] = 1;
        }
        for (int k = 0; k < 3; k++)
            for (size_t q = 0; q < chunks[k].size(); q++) {
                const Chunk &ch = chunks[k][q];
                put(o, "static __device__ __noinline__ void krb_s%d_c%zu(T *__restrict__ s, const KCtx &c, int &status, bool &notconv, uint32_t &grow) {\n", k, q);
                char is_in[KRB_MAX_SLOTS] = {};
                for (int sl : ch.live_in) { is_in[sl] = 1; put(o, "T v%d = s[%d];\n", sl, sl); }
                for (int sl : ch.written) if (!is_in[sl]) put(o, "T v%d;\n", sl);
                for (size_t i = ch.begin; i < ch.end; i++) emit_stmt(ops[k][i], "v", "", o);
                for (int sl : ch.written) if (any_in[sl]) put(o, "s[%d] = v%d;\n", sl, sl);
                o += "}\n";
            }
    }
    o += R"KRBJIT(
KRB_JIT_ENTRY {
#ifndef KRB_JIT_HOST
    const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#endif
    const size_t L = (size_t)p.L;
    KCtx c;
    c.IN = p.IN + lane; c.G = (T *)p.G + lane; c.Bc = p.B; c.L = L;
    c.reltol = p.reltol; c.iabstol = p.iabstol; c.vabstol = p.vabstol; c.nonlinear = p.nonlinear != 0;
    int status = 0, it_used = 0;
    uint32_t grow = 0;
    T s[KRB_JIT_NSLOTS];   // the interpreter's scratch cells: constant indices only (registers), or the hand-over array between chunks
    for (int i = 0; i < KRB_JIT_NSLOTS; i++) s[i] = k_zero();
    for (int row = p.row_begin; row < p.row_end; ++row) {
        T *Xcur = (T *)p.X + (size_t)row * (size_t)p.row_width * L + lane;
        const double *Xprev;
        if (row == p.row_begin && p.Xprev0) Xprev = p.Xprev0 + lane;
        else Xprev = (const double *)((T *)p.X + (size_t)(row > 0 ? row - 1 : 0) * (size_t)p.row_width * L) + lane;
        c.Xcur = Xcur; c.Xprev = Xprev; c.lane_active = true;
        const int n_pass = (p.sec1 && row == p.row_begin) ? 2 : 1;
        for (int pass = 0; pass < n_pass; ++pass) {
            const int prog = p.sec1 ? ((n_pass == 2 && pass == 0) ? 1 : 2) : 0;
            const bool solve_pass = !(n_pass == 2 && pass == 0);
            for (int iter = 0; iter < p.max_iter; ++iter) {
                c.guess_empty = p.guess_empty_at_iter0 && iter == 0;
                c.Xguess = (iter == 0 && !p.guess_empty_at_iter0) ? Xprev : (const double *)Xcur;
                c.Bs = p.Bstep + ((size_t)row * 2 + (iter > 0 ? 1 : 0)) * (size_t)p.n_wave;
                bool notconv = false;
)KRBJIT";
    for (int k = 0; k < 3; k++) {
        put(o, k == 0 ? "                if (prog == 0) {\n" : k == 1 ? "                } else if (prog == 1) {\n" : "                } else {\n");
        if (chunked) for (size_t q = 0; q < chunks[k].size(); q++) put(o, "krb_s%d_c%zu(s, c, status, notconv, grow);\n", k, q);
        else for (const DecOp &op : ops[k]) emit_stmt(op, "s[", "]", o);
    }
    o += R"KRBJIT(                }
                // ---- NR bookkeeping (op.rs:65-81, dc.rs:99-110, transient.rs:97-109), as vm_kernel.cuh ----
                if (!solve_pass || !p.nonlinear) break;
                bool pseudo_ok = true;
                if (p.pseudo_time && iter == 0) {
                    const double tn = (double)row * p.h, tp = (double)(row - 1) * p.h;
                    pseudo_ok = fabs(tn - tp) <= p.reltol * fmax(fabs(tn), fabs(tp)) + p.vabstol;
                }
                if (c.lane_active) it_used = iter + 1;
                const bool conv = !c.guess_empty && !notconv && pseudo_ok;
                if (c.lane_active && conv) c.lane_active = false;
                if (c.lane_active && iter == p.max_iter - 1) status |= KRB_LANE_NOT_CONVERGED;
                if (KRB_JIT_ALL(!c.lane_active)) break;
            }
        }
    }
    if (!p.nonlinear && grow >= KRB_GROWTH_KEY_V && grow < KRB_NONFINITE_KEY_V) status |= KRB_LANE_INACCURATE;
    if (p.status) p.status[lane] |= status;
    if (p.iters) p.iters[lane] = it_used;
}
)KRBJIT";
    return o;
}

// ---- NVRTC -----------------------------------------------------------------------------------------------------------------
namespace {
struct Nvrtc {
    void *h = nullptr;
    nvrtcResult (*create)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*compile)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*cubin_size)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*cubin)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*log_size)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*log)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*destroy)(nvrtcProgram *) = nullptr;
    const char *(*errstr)(nvrtcResult) = nullptr;
    std::string why;
    bool ok = false;
};
struct Driver {
    void *h = nullptr;
    CUresult (*ctx_current)(CUcontext *) = nullptr;
    CUresult (*load)(CUmodule *, const void *) = nullptr;
    CUresult (*get_fn)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*launch)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **) = nullptr;
    CUresult (*unload)(CUmodule) = nullptr;
    CUresult (*errstr)(CUresult, const char **) = nullptr;
    std::string why;
    bool ok = false;
};
template <class F> bool bind(void *h, const char *name, F &fn, std::string &why) {
    fn = reinterpret_cast<F>(dlsym(h, name));
    if (!fn) { why = std::string("symbol ") + name + " not found"; return false; }
    return true;
}
Nvrtc &nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char *nm : names) if ((n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
        if (!n.h) { n.why = "libnvrtc.so.12 not found"; return; }
        n.ok = bind(n.h, "nvrtcCreateProgram", n.create, n.why) && bind(n.h, "nvrtcCompileProgram", n.compile, n.why) &&
               bind(n.h, "nvrtcGetCUBINSize", n.cubin_size, n.why) && bind(n.h, "nvrtcGetCUBIN", n.cubin, n.why) &&
               bind(n.h, "nvrtcGetProgramLogSize", n.log_size, n.why) && bind(n.h, "nvrtcGetProgramLog", n.log, n.why) &&
               bind(n.h, "nvrtcDestroyProgram", n.destroy, n.why) && bind(n.h, "nvrtcGetErrorString", n.errstr, n.why);
    });
    return n;
}
Driver &driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libcuda.so.1", "libcuda.so"};
        for (const char *nm : names) if ((d.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
        if (!d.h) { d.why = "libcuda.so.1 not found"; return; }
        d.ok = bind(d.h, "cuCtxGetCurrent", d.ctx_current, d.why) && bind(d.h, "cuModuleLoadData", d.load, d.why) &&
               bind(d.h, "cuModuleGetFunction", d.get_fn, d.why) && bind(d.h, "cuLaunchKernel", d.launch, d.why) &&
               bind(d.h, "cuModuleUnload", d.unload, d.why) && bind(d.h, "cuGetErrorString", d.errstr, d.why);
    });
    return d;
}
std::string cu_err(Driver &d, CUresult r) {
    const char *s = nullptr;
    if (d.errstr) d.errstr(r, &s);
    return s ? s : ("CUresult " + std::to_string((int)r));
}
}  // namespace

bool vm_jit_compile(const std::string &src, std::string &cubin, std::string &err) {
    Nvrtc &n = nvrtc();
    if (!n.ok) { err = "NVRTC unavailable: " + n.why; return false; }
    nvrtcProgram prog = nullptr;
    nvrtcResult r = n.create(&prog, src.c_str(), "krb_jit_vm.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) { err = std::string("nvrtcCreateProgram: ") + n.errstr(r); return false; }
    // -lineinfo only for short programs (it triples the cubin and nobody reads the source view of a 10 000-line kernel)
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    r = n.compile(prog, src.size() < 400000 ? 3 : 2, opts);
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        n.log_size(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.log(prog, &log[0]);
        if (log.size() > 2000) log.resize(2000);
        err = std::string("nvrtcCompileProgram: ") + n.errstr(r) + "\n" + log;
        n.destroy(&prog);
        return false;
    }
    size_t cs = 0;
    r = n.cubin_size(prog, &cs);
    if (r != NVRTC_SUCCESS || cs == 0) { err = "nvrtcGetCUBINSize failed"; n.destroy(&prog); return false; }
    cubin.assign(cs, '\0');
    r = n.cubin(prog, &cubin[0]);
    n.destroy(&prog);
    if (r != NVRTC_SUCCESS) { err = "nvrtcGetCUBIN failed"; return false; }
    return true;
}

struct JitKernel {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
};

// ---- on-disk cubin cache ---------------------------------------------------------------------------------------------------
// KRB_VM_JIT_CACHE=<directory>: compiled kernels are kept there as <fnv1a64(source + NVRTC version)>.cubin, so that a later process
// (the next sweep of the same circuit) skips NVRTC.  Off unless the variable is set: the library never writes outside /tmp or the
// caller's directories on its own.
static std::string cache_path(const std::string &src) {
    const char *dir = getenv("KRB_VM_JIT_CACHE");
    if (!dir || !dir[0]) return std::string();
    uint64_t h = 1469598103934665603ull;
    for (unsigned char ch : src) { h ^= ch; h *= 1099511628211ull; }
    for (const char *v = "nvrtc-12/sm_100a/v1"; *v; ++v) { h ^= (unsigned char)*v; h *= 1099511628211ull; }
    char name[40];
    snprintf(name, sizeof name, "/%016llx.cubin", (unsigned long long)h);
    return std::string(dir) + name;
}
static bool cache_read(const std::string &path, std::string &cubin) {
    if (path.empty()) return false;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[65536];
    size_t k;
    cubin.clear();
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) cubin.append(buf, k);
    fclose(f);
    return cubin.size() > 64;
}
static void cache_write(const std::string &path, const std::string &cubin) {
    if (path.empty()) return;
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    fclose(f);
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());   // atomic publish: concurrent ranks write the same bytes
}

JitKernel *vm_jit_build(const Plan &plan, size_t vm_params_size, std::string &err) {
    Driver &d = driver();
    if (!d.ok) { err = "CUDA driver API unavailable: " + d.why; return nullptr; }
    CUcontext ctx = nullptr;
    if (d.ctx_current(&ctx) != CUDA_SUCCESS || !ctx) { err = "no current CUDA context"; return nullptr; }
    const std::string src = vm_jit_source(plan, vm_params_size, false);
    const std::string path = cache_path(src);
    std::string cubin;
    JitKernel *k = new JitKernel();
    CUresult r = CUDA_ERROR_UNKNOWN;
    if (cache_read(path, cubin)) r = d.load(&k->mod, cubin.data());          // a stale or truncated file just falls through to NVRTC
    if (r != CUDA_SUCCESS) {
        if (!vm_jit_compile(src, cubin, err)) { delete k; return nullptr; }
        r = d.load(&k->mod, cubin.data());
        if (r == CUDA_SUCCESS) cache_write(path, cubin);
    }
    if (r != CUDA_SUCCESS) { err = "cuModuleLoadData: " + cu_err(d, r); delete k; return nullptr; }
    r = d.get_fn(&k->fn, k->mod, "krb_jit_vm");
    if (r != CUDA_SUCCESS) { err = "cuModuleGetFunction: " + cu_err(d, r); d.unload(k->mod); delete k; return nullptr; }
    return k;
}

bool vm_jit_launch(JitKernel *k, const void *params, unsigned grid, unsigned block, void *stream, std::string &err) {
    Driver &d = driver();
    void *args[1] = {const_cast<void *>(params)};
    const CUresult r = d.launch(k->fn, grid, 1, 1, block, 1, 1, 0, (CUstream)stream, args, nullptr);
    if (r != CUDA_SUCCESS) { err = "cuLaunchKernel (specialised lane kernel): " + cu_err(d, r); return false; }
    return true;
}

void vm_jit_free(JitKernel *k) {
    if (!k) return;
    Driver &d = driver();
    if (d.ok && k->mod) d.unload(k->mod);
    delete k;
}

}  // namespace krb
