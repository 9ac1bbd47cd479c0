/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see krets_oracle.h).  Plain C restatement of the krets CPU hot path.
 * All citations are relative to /root/reference.
 */
#define _GNU_SOURCE
#include "krets_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * scalar types + the two LU instantiations
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double re, im; } ko_c64;

/* num_complex::Complex<f64> arithmetic (faer::c64): naive formulas, no inf/nan special-casing */
static inline ko_c64 c_sub_mul(ko_c64 x, ko_c64 a, ko_c64 b) {
    ko_c64 r = { x.re - (a.re * b.re - a.im * b.im), x.im - (a.re * b.im + a.im * b.re) };
    return r;
}
static inline ko_c64 c_div(ko_c64 a, ko_c64 b) {
    double ns = b.re * b.re + b.im * b.im;
    ko_c64 r = { (a.re * b.re + a.im * b.im) / ns, (a.im * b.re - a.re * b.im) / ns };
    return r;
}
static const ko_c64 KO_C_ZERO = { 0.0, 0.0 };
static const ko_c64 KO_C_ONE = { 1.0, 0.0 };

#define KO_T double
#define KO_SFX _r
#define KO_ZERO 0.0
#define KO_ONE 1.0
#define KO_ABS2(v) ((v) * (v))
#define KO_SUB_MUL(x, a, b) ((x) -= (a) * (b))
#define KO_DIV(a, b) ((a) / (b))
#include "ko_lu_template.h"
#undef KO_T
#undef KO_SFX
#undef KO_ZERO
#undef KO_ONE
#undef KO_ABS2
#undef KO_SUB_MUL
#undef KO_DIV

#define KO_T ko_c64
#define KO_SFX _c
#define KO_ZERO KO_C_ZERO
#define KO_ONE KO_C_ONE
#define KO_ABS2(v) ((v).re * (v).re + (v).im * (v).im)
#define KO_SUB_MUL(x, a, b) ((x) = c_sub_mul((x), (a), (b)))
#define KO_DIV(a, b) c_div((a), (b))
#include "ko_lu_template.h"
#undef KO_T
#undef KO_SFX
#undef KO_ZERO
#undef KO_ONE
#undef KO_ABS2
#undef KO_SUB_MUL
#undef KO_DIV

/* ------------------------------------------------------------------------------------------------
 * circuit IR  (crates/krets-parser/src/circuit.rs:7-20, elements.rs:14-24)
 * ---------------------------------------------------------------------------------------------- */
#define KO_NAME 64

typedef struct {
    int kind;
    char name[KO_NAME];      /* text after the type letter(s) */
    char ident[KO_NAME + 2]; /* resistor.rs:21-23 etc.: "R"+name, ... */
    char plus[KO_NAME], minus[KO_NAME], third[KO_NAME]; /* NMOS: drain(plus) gate(minus) source(third) */
    double value;            /* R/C/L value, V dc_value, I value */
    int g2;                  /* capacitor.rs:31,38: trailing G2 token */
    double ac_amplitude;
    int has_pulse; double pulse[7];  /* voltage_source.rs:15-31 */
    int has_sin; double sinp[6];     /* voltage_source.rs:35-48 */
    char model_name[KO_NAME];
    double is, n_emis;       /* models/diode.rs:14-23 */
    int ip, im, ic;          /* cached index_map lookups; -1 = key absent */
    int is3;                 /* NMOS: index of V(<source>) (extension below) */
    double vth, beta, lambda;/* NMOS level-1 parameters: models/nmosfet.rs:24-45 (beta = KP * (W / L)) */
} ko_elem;

typedef struct { char name[KO_NAME]; int type; /* 0 D, 1 NMOS, 2 PMOS */ double is, n_emis;
                 double w, l, vto, kp, lambda; /* models/nmosfet.rs:24-36 */ } ko_model;

struct ko_circuit {
    ko_elem *e; int ne, ecap;
    char **uname; int n, ucap;     /* index_map inverted: unknown index -> key */
    int *is_current;               /* key starts with 'I' (solver.rs:96) */
    ko_model *models; int nmodels;
    /* tiny string hash for index_map */
    int *hslot; int hcap;
};

static uint64_t str_hash(const char *s) {
    uint64_t h = 1469598103934665603ULL;
    while (*s) { h ^= (unsigned char)*s++; h *= 1099511628211ULL; }
    return h;
}
static void map_rehash(ko_circuit *c, int cap) {
    free(c->hslot);
    c->hcap = cap;
    c->hslot = (int *)malloc(sizeof(int) * cap);
    for (int i = 0; i < cap; i++) c->hslot[i] = -1;
    for (int i = 0; i < c->n; i++) {
        uint64_t h = str_hash(c->uname[i]) & (cap - 1);
        while (c->hslot[h] >= 0) h = (h + 1) & (cap - 1);
        c->hslot[h] = i;
    }
}
int ko_index_of(const ko_circuit *c, const char *name) {
    if (!c->hcap) return -1;
    uint64_t h = str_hash(name) & (c->hcap - 1);
    while (c->hslot[h] >= 0) {
        if (strcmp(c->uname[c->hslot[h]], name) == 0) return c->hslot[h];
        h = (h + 1) & (c->hcap - 1);
    }
    return -1;
}
static int map_insert(ko_circuit *c, const char *key) {
    if (c->n == c->ucap) {
        c->ucap = c->ucap ? 2 * c->ucap : 64;
        c->uname = (char **)realloc(c->uname, sizeof(char *) * c->ucap);
    }
    c->uname[c->n] = strdup(key);
    c->n++;
    if (2 * c->n > c->hcap) map_rehash(c, c->hcap ? 2 * c->hcap : 128);
    else {
        uint64_t h = str_hash(key) & (c->hcap - 1);
        while (c->hslot[h] >= 0) h = (h + 1) & (c->hcap - 1);
        c->hslot[h] = c->n - 1;
    }
    return c->n - 1;
}

void ko_default_config(ko_config *cfg) {   /* config.rs:26-36 */
    cfg->relative_tolerance = 0.001;
    cfg->current_absolute_tolerance = 1e-12;
    cfg->voltage_absolute_tolerance = 1e-6;
    cfg->maximum_iterations = 300;
}

/* ------------------------------------------------------------------------------------------------
 * lexing helpers (nom combinators used by the reference, restated)
 * ---------------------------------------------------------------------------------------------- */
static int is_sp(char ch) { return ch == ' ' || ch == '\t'; }            /* nom space0/space1 */
static const char *space0(const char *p) { while (is_sp(*p)) p++; return p; }
static const char *space1(const char *p) { if (!is_sp(*p)) return NULL; return space0(p); }
/* utils.rs:66-69 alphanumeric_or_underscore1 */
static const char *ident1(const char *p, char *out) {
    int k = 0;
    while ((isalnum((unsigned char)*p) && (unsigned char)*p < 128) || *p == '_') { if (k < KO_NAME - 1) out[k++] = *p; p++; }
    out[k] = 0;
    return k ? p : NULL;
}
static int tag_nc(const char *p, const char *t) {      /* tag_no_case; returns length or 0 */
    int k = 0;
    while (t[k]) { if (tolower((unsigned char)p[k]) != tolower((unsigned char)t[k])) return 0; k++; }
    return k;
}
/* Rust <f64 as FromStr> grammar on an upper-cased string: sign? (INF|INFINITY|NAN | digits[.digits*][E sign? digits] | .digits...) */
static int rust_parse_f64(const char *s, double *out) {
    const char *p = s;
    if (*p == '+' || *p == '-') p++;
    if (!strcmp(p, "INF") || !strcmp(p, "INFINITY")) { *out = (s[0] == '-') ? -INFINITY : INFINITY; return 1; }
    if (!strcmp(p, "NAN")) { *out = NAN; return 1; }
    int nd = 0;
    while (isdigit((unsigned char)*p)) { p++; nd++; }
    if (*p == '.') { p++; while (isdigit((unsigned char)*p)) { p++; nd++; } }
    if (!nd) return 0;
    if (*p == 'E' || *p == 'e') {
        p++;
        if (*p == '+' || *p == '-') p++;
        if (!isdigit((unsigned char)*p)) return 0;
        while (isdigit((unsigned char)*p)) p++;
    }
    if (*p) return 0;
    *out = strtod(s, NULL);
    return 1;
}
/* utils.rs:32-64 parse_value */
static int parse_value_tok(const char *tok, int len, double *out) {
    char up[128];
    if (len <= 0 || len >= (int)sizeof(up)) return 0;
    for (int i = 0; i < len; i++) up[i] = (char)toupper((unsigned char)tok[i]);
    up[len] = 0;
    double mult = 1.0;
    if (len >= 3 && !strcmp(up + len - 3, "MEG")) { up[len - 3] = 0; mult = 1e6; }
    else {
        switch (up[len - 1]) {
            case 'F': mult = 1e-15; up[len - 1] = 0; break;
            case 'P': mult = 1e-12; up[len - 1] = 0; break;
            case 'N': mult = 1e-9; up[len - 1] = 0; break;
            case 'U': mult = 1e-6; up[len - 1] = 0; break;
            case 'M': mult = 1e-3; up[len - 1] = 0; break;
            case 'K': mult = 1e3; up[len - 1] = 0; break;
            case 'G': mult = 1e9; up[len - 1] = 0; break;
            case 'T': mult = 1e12; up[len - 1] = 0; break;
            default: break;
        }
    }
    double base;
    if (!rust_parse_f64(up, &base)) return 0;
    *out = base * mult;
    return 1;
}
/* utils.rs:72-78 value_parser: token = is_not(" \t\r\n()") */
static const char *value_parser(const char *p, double *out) {
    const char *q = p;
    while (*q && *q != ' ' && *q != '\t' && *q != '\r' && *q != '\n' && *q != '(' && *q != ')') q++;
    if (q == p) return NULL;
    if (!parse_value_tok(p, (int)(q - p), out)) return NULL;
    return q;
}
/* nom::number::complete::double, float-prefix part (voltage_source.rs:194,101,106) */
static const char *nom_double(const char *p, double *out) {
    const char *s = p;
    if (*p == '+' || *p == '-') p++;
    int nd = 0;
    const char *q = p;
    while (isdigit((unsigned char)*q)) { q++; nd++; }
    if (*q == '.') { const char *r = q + 1; int fd = 0; while (isdigit((unsigned char)*r)) { r++; fd++; } if (nd || fd) { q = r; nd += fd; } }
    if (!nd) return NULL;
    if (*q == 'e' || *q == 'E') {
        const char *r = q + 1;
        if (*r == '+' || *r == '-') r++;
        if (!isdigit((unsigned char)*r)) return NULL;   /* nom: cut(digit1) -> hard failure */
        while (isdigit((unsigned char)*r)) r++;
        q = r;
    }
    char buf[128];
    int len = (int)(q - s);
    if (len >= (int)sizeof(buf)) return NULL;
    memcpy(buf, s, len); buf[len] = 0;
    *out = strtod(buf, NULL);
    return q;
}

/* ------------------------------------------------------------------------------------------------
 * element line parsers
 * ---------------------------------------------------------------------------------------------- */
/* shared head: <letter(s)><name> <plus> <minus> */
static const char *parse_two_node_head(const char *p, int taglen, ko_elem *e) {
    p += taglen;
    if (!(p = ident1(p, e->name))) return NULL;
    if (!(p = space1(p))) return NULL;
    if (!(p = ident1(p, e->plus))) return NULL;
    if (!(p = space1(p))) return NULL;
    if (!(p = ident1(p, e->minus))) return NULL;
    return p;
}
/* resistor.rs:36-52, capacitor.rs:25-42, inductor.rs:23-38, current_source.rs:23-38 */
static int parse_rcli(const char *line, ko_elem *e, int kind, char letter) {
    const char *p = parse_two_node_head(line, 1, e);
    if (!p) return 0;
    if (!(p = space1(p))) return 0;
    if (!(p = value_parser(p, &e->value))) return 0;
    e->kind = kind;
    snprintf(e->ident, sizeof(e->ident), "%c%s", letter, e->name);
    if (kind == KO_C) {
        const char *q = space1(p);
        if (q && tag_nc(q, "G2")) e->g2 = 1;
    }
    return 1;
}
static const char *parse_paren_values(const char *p, double *v, int count) {
    p = space0(p);
    if (*p != '(') return NULL;
    p++;
    p = space0(p);
    for (int i = 0; i < count; i++) {
        if (i) { if (!(p = space1(p))) return NULL; }
        if (!(p = value_parser(p, &v[i]))) return NULL;
    }
    p = space0(p);
    if (*p != ')') return NULL;
    return p + 1;
}
/* voltage_source.rs:188-235 */
static int parse_vsource(const char *line, ko_elem *e) {
    const char *p = parse_two_node_head(line, 1, e);
    if (!p) return 0;
    e->kind = KO_V;
    snprintf(e->ident, sizeof(e->ident), "V%s", e->name);
    double dc = 0.0, v;
    const char *q = space1(p);
    if (q) { const char *r = nom_double(q, &v); if (r) { dc = v; p = r; } }
    for (;;) {                                             /* many0(preceded(space1, alt(dc, ac, pulse, sin))) */
        q = space1(p);
        if (!q) break;
        int k;
        const char *r;
        if ((k = tag_nc(q, "dc")) && (r = space1(q + k)) && (r = nom_double(r, &v))) { dc = v; p = r; continue; }
        if ((k = tag_nc(q, "ac")) && (r = space1(q + k)) && (r = nom_double(r, &v))) { e->ac_amplitude = v; p = r; continue; }
        if ((k = tag_nc(q, "pulse")) && (r = parse_paren_values(q + k, e->pulse, 7))) { e->has_pulse = 1; p = r; continue; }
        if ((k = tag_nc(q, "sin")) && (r = parse_paren_values(q + k, e->sinp, 6))) { e->has_sin = 1; p = r; continue; }
        break;
    }
    e->value = dc;
    return 1;
}
/* diode.rs:74-92 */
static int parse_diode(const char *line, ko_elem *e) {
    const char *p = parse_two_node_head(line, 1, e);
    if (!p) return 0;
    e->kind = KO_D;
    snprintf(e->ident, sizeof(e->ident), "D%s", e->name);
    strcpy(e->model_name, "default");
    const char *q = space1(p);
    char tmp[KO_NAME];
    if (q && ident1(q, tmp)) strcpy(e->model_name, tmp);
    e->is = 1e-12; e->n_emis = 1.0;
    return 1;
}
/* nmosfet.rs:113-166: MN<name> drain gate source bulk model [k=v]*.  Stamps nothing (stampable.rs:842-844). */
static int parse_nmos(const char *line, ko_elem *e) {
    const char *p = line + 2;
    char bulk[KO_NAME];
    if (!(p = ident1(p, e->name))) return 0;
    if (!(p = space1(p)) || !(p = ident1(p, e->plus))) return 0;
    if (!(p = space1(p)) || !(p = ident1(p, e->minus))) return 0;
    if (!(p = space1(p)) || !(p = ident1(p, e->third))) return 0;
    if (!(p = space1(p)) || !(p = ident1(p, bulk))) return 0;
    if (!(p = space1(p)) || !(p = ident1(p, e->model_name))) return 0;
    e->kind = KO_NMOS;
    snprintf(e->ident, sizeof(e->ident), "M%s", e->name);
    return 1;
}
/* models.rs:52-95 + models/diode.rs:26-39 */
static int parse_model_line(const char *line, ko_model *m) {
    const char *p = line + 6;
    if (!(p = space1(p))) return 0;
    if (!(p = ident1(p, m->name))) return 0;
    if (!(p = space1(p))) return 0;
    if (!strncmp(p, "NMOS", 4)) { m->type = 1; p += 4; }
    else if (!strncmp(p, "PMOS", 4)) { m->type = 2; p += 4; }
    else if (*p == 'D') { m->type = 0; p += 1; }
    else return 0;
    m->is = 1e-12; m->n_emis = 1.0;
    m->w = 1e-6; m->l = 1e-6; m->vto = 0.0; m->kp = 2e-5; m->lambda = 0.02;   /* models/nmosfet.rs:24-36 */
    p = space0(p);
    if (*p == '(') p++;
    for (;;) {
        const char *q = space0(p);
        char key[KO_NAME];
        const char *r = ident1(q, key);
        if (!r) break;
        r = space0(r);
        if (*r != '=') break;
        r = space0(r + 1);
        double v;
        if (!(r = value_parser(r, &v))) break;
        for (char *k = key; *k; k++) *k = (char)tolower((unsigned char)*k);
        if (m->type == 0) { if (!strcmp(key, "is")) m->is = v; else if (!strcmp(key, "n")) m->n_emis = v; }
        else if (m->type == 1) {                                                     /* models/nmosfet.rs:48-61 */
            if (!strcmp(key, "w")) m->w = v; else if (!strcmp(key, "l")) m->l = v; else if (!strcmp(key, "vto")) m->vto = v;
            else if (!strcmp(key, "kp")) m->kp = v; else if (!strcmp(key, "lambda")) m->lambda = v;
        }
        p = r;
    }
    return 1;
}

static int starts_with_nc(const char *s, const char *t) { return tag_nc(s, t) > 0; }

static void set_err(int *err, int code, char *msg, int len, const char *fmt, int line, const char *text) {
    if (err) *err = code;
    if (msg && len > 0) snprintf(msg, len, fmt, line, text);
}

/* ---- subcircuits: krets-parser/src/elements/subcircuit.rs ------------------------------------------------------- */
#define KO_MAX_PINS 32
typedef struct { char instance_name[KO_NAME], definition_name[KO_NAME]; char nodes[KO_MAX_PINS][KO_NAME]; int nn; } ko_inst;
typedef struct { int is_instance; ko_elem e; ko_inst x; } ko_defelem;
typedef struct { char name[KO_NAME]; char pins[KO_MAX_PINS][KO_NAME]; int npins; ko_defelem *el; int nel; } ko_subckt;

/* parse_subckt_instance, subcircuit.rs:147-172: x<name> n1 n2 ... <definition> */
static int parse_instance(const char *line, ko_inst *x) {
    const char *p = line + 1;
    char toks[KO_MAX_PINS + 1][KO_NAME];
    int nt = 0;
    if (!(p = ident1(p, x->instance_name))) return 0;
    for (;;) {
        const char *q = space1(p);
        if (!q) break;
        char tok[KO_NAME];
        const char *r = ident1(q, tok);
        if (!r) break;
        if (nt > KO_MAX_PINS) return 0;
        strcpy(toks[nt++], tok);
        p = r;
    }
    if (nt == 0) return 0;
    strcpy(x->definition_name, toks[nt - 1]);
    x->nn = nt - 1;
    for (int i = 0; i < x->nn; i++) strcpy(x->nodes[i], toks[i]);
    return 1;
}

/* one element line (elements.rs:45-55: alt(R, C, L, V, I, D, BJT, NMOS, subckt), each keyed by its leading tag) */
static int parse_element_line(const char *s, ko_defelem *de, int *code) {
    memset(de, 0, sizeof(*de));
    ko_elem *e = &de->e;
    e->ip = e->im = e->ic = -1;
    char c0 = (char)toupper((unsigned char)s[0]);
    *code = KO_ERR_PARSE;
    if (c0 == 'R') return parse_rcli(s, e, KO_R, 'R');
    if (c0 == 'C') return parse_rcli(s, e, KO_C, 'C');
    if (c0 == 'L') return parse_rcli(s, e, KO_L, 'L');
    if (c0 == 'V') return parse_vsource(s, e);
    if (c0 == 'I') return parse_rcli(s, e, KO_I, 'I');
    if (c0 == 'D') return parse_diode(s, e);
    if (c0 == 'M' && toupper((unsigned char)s[1]) == 'N') return parse_nmos(s, e);
    if (c0 == 'X') { de->is_instance = 1; return parse_instance(s, &de->x); }
    if (c0 == 'Q') { *code = KO_ERR_UNSUPPORTED; return 0; }
    return 0;
}

static void push_elem(ko_circuit *c, const ko_elem *e) {
    if (c->ne == c->ecap) { c->ecap = c->ecap ? 2 * c->ecap : 64; c->e = (ko_elem *)realloc(c->e, sizeof(ko_elem) * c->ecap); }
    c->e[c->ne++] = *e;
}

/* map_sub_element, subcircuit.rs:85-116: pins map to the instance's nodes, every other node (a "0" that is not a pin included)
 * and the element name get "<instance>_" in front */
static int map_node(char *node, const ko_subckt *def, const ko_inst *x) {
    for (int k = def->npins - 1; k >= 0; k--)            /* HashMap::collect: a repeated pin keeps its last node */
        if (!strcmp(def->pins[k], node)) { strcpy(node, x->nodes[k]); return 1; }
    char tmp[2 * KO_NAME + 2];
    snprintf(tmp, sizeof(tmp), "%s_%s", x->instance_name, node);
    if (strlen(tmp) >= KO_NAME) return 0;
    strcpy(node, tmp);
    return 1;
}

/* SubcircuitInstance::instantiate, subcircuit.rs:33-82 */
static int instantiate(ko_circuit *c, const ko_inst *x, const ko_subckt *defs, int ndefs, int depth) {
    if (depth > 64) return 0;
    const ko_subckt *def = NULL;
    for (int i = 0; i < ndefs; i++) if (!strcmp(defs[i].name, x->definition_name)) def = &defs[i];   /* later definitions replace earlier ones */
    if (!def || def->npins != x->nn) return 0;
    for (int i = 0; i < def->nel; i++) {
        ko_defelem m = def->el[i];
        char tmp[2 * KO_NAME + 2];
        if (m.is_instance) {
            for (int k = 0; k < m.x.nn; k++) if (!map_node(m.x.nodes[k], def, x)) return 0;
            snprintf(tmp, sizeof(tmp), "%s_%s", x->instance_name, m.x.instance_name);
            if (strlen(tmp) >= KO_NAME) return 0;
            strcpy(m.x.instance_name, tmp);
            if (!instantiate(c, &m.x, defs, ndefs, depth + 1)) return 0;
        } else {
            if (!map_node(m.e.plus, def, x) || !map_node(m.e.minus, def, x)) return 0;
            if (m.e.kind == KO_NMOS && !map_node(m.e.third, def, x)) return 0;
            snprintf(tmp, sizeof(tmp), "%s_%s", x->instance_name, m.e.name);
            if (strlen(tmp) >= KO_NAME) return 0;
            strcpy(m.e.name, tmp);
            snprintf(m.e.ident, sizeof(m.e.ident), "%c%s", m.e.ident[0], m.e.name);                   /* identifier = letter + name */
            push_elem(c, &m.e);
        }
    }
    return 1;
}

/* parse_subcircuits, subcircuit.rs:174-218 (pre-pass over the whole text) */
static int parse_subcircuits(const char *text, ko_subckt **out, int *nout) {
    ko_subckt *defs = NULL;
    int nd = 0, cur = -1;
    const char *p = text;
    while (*p) {
        const char *eol = strchr(p, '\n');
        size_t len = eol ? (size_t)(eol - p) : strlen(p);
        char *buf = (char *)malloc(len + 1);
        memcpy(buf, p, len); buf[len] = 0;
        p = eol ? eol + 1 : p + len;
        char *s = buf;
        while (*s && isspace((unsigned char)*s)) s++;
        char *t = s + strlen(s);
        while (t > s && isspace((unsigned char)t[-1])) *--t = 0;
        int ok = 1;
        if (!*s || *s == '%' || *s == '*') { free(buf); continue; }
        if (starts_with_nc(s, ".subckt")) {
            ko_subckt d; memset(&d, 0, sizeof(d));
            const char *q = space1(s + 7);
            q = q ? ident1(q, d.name) : NULL;
            if (!q) ok = 0;
            while (ok) {
                const char *r = space1(q);
                char pin[KO_NAME];
                if (!r || !(r = ident1(r, pin))) break;
                if (d.npins == KO_MAX_PINS) { ok = 0; break; }
                strcpy(d.pins[d.npins++], pin);
                q = r;
            }
            if (ok) {
                defs = (ko_subckt *)realloc(defs, sizeof(ko_subckt) * (nd + 1));
                defs[nd] = d;
                cur = nd++;
            }
        } else if (starts_with_nc(s, ".ends")) cur = -1;
        else if (cur >= 0) {
            ko_defelem de;
            int code;
            if (!parse_element_line(s, &de, &code)) ok = 0;
            else {
                defs[cur].el = (ko_defelem *)realloc(defs[cur].el, sizeof(ko_defelem) * (defs[cur].nel + 1));
                defs[cur].el[defs[cur].nel++] = de;
            }
        }
        free(buf);
        if (!ok) { for (int i = 0; i < nd; i++) free(defs[i].el); free(defs); return 0; }
    }
    *out = defs; *nout = nd;
    return 1;
}

/* parser.rs:27-154 */
ko_circuit *ko_parse(const char *text, int *err, char *errmsg, int errmsg_len) {
    ko_circuit *c = (ko_circuit *)calloc(1, sizeof(*c));
    if (err) *err = KO_OK;
    int in_control = 0, in_subckt = 0, lineno = 0;
    ko_subckt *defs = NULL;
    int ndefs = 0;
    if (!parse_subcircuits(text, &defs, &ndefs)) {
        set_err(err, KO_ERR_PARSE, errmsg, errmsg_len, "Failed to parse a subcircuit definition%.0d%s", 0, "");
        ko_free(c);
        return NULL;
    }
    const char *p = text;
    while (*p) {
        const char *eol = strchr(p, '\n');
        size_t len = eol ? (size_t)(eol - p) : strlen(p);
        char *buf = (char *)malloc(len + 1);
        memcpy(buf, p, len); buf[len] = 0;
        p = eol ? eol + 1 : p + len;
        lineno++;
        /* str::trim(): unicode whitespace; ASCII subset suffices for netlists */
        char *s = buf;
        while (*s && isspace((unsigned char)*s)) s++;
        char *t = s + strlen(s);
        while (t > s && isspace((unsigned char)t[-1])) *--t = 0;
        int ok = 1, code = KO_ERR_PARSE;
        if (!*s || *s == '%' || *s == '*') goto next;
        if (starts_with_nc(s, ".control")) { in_control = 1; goto next; }
        if (starts_with_nc(s, ".endc")) { in_control = 0; goto next; }
        if (in_control) goto next;
        if (starts_with_nc(s, ".subckt")) { in_subckt = 1; goto next; }
        if (starts_with_nc(s, ".ends")) { in_subckt = 0; goto next; }
        if (in_subckt) goto next;
        if (strlen(s) == 4 && starts_with_nc(s, ".end")) goto next;
        if (starts_with_nc(s, ".model")) {
            char *pc = strchr(s, '%');            /* models.rs:87 */
            if (pc) *pc = 0;
            ko_model m; memset(&m, 0, sizeof(m));
            if (!parse_model_line(s, &m)) { ok = 0; }
            else {
                int found = -1;
                for (int i = 0; i < c->nmodels; i++) if (!strcmp(c->models[i].name, m.name)) found = i;
                if (found < 0) { c->models = (ko_model *)realloc(c->models, sizeof(ko_model) * (c->nmodels + 1)); found = c->nmodels++; }
                c->models[found] = m;
            }
            if (!ok) { set_err(err, code, errmsg, errmsg_len, "Parse error on line %d: %s", lineno, s); free(buf); ko_free(c); return NULL; }
            goto next;
        }
        {
            ko_defelem de;
            ok = parse_element_line(s, &de, &code);
            if (ok && de.is_instance) ok = instantiate(c, &de.x, defs, ndefs, 0);      /* parser.rs:93-98 */
            else if (ok) push_elem(c, &de.e);
            if (!ok) {
                set_err(err, code, errmsg, errmsg_len, "Parse error on line %d: %s", lineno, s);
                free(buf); ko_free(c);
                for (int i = 0; i < ndefs; i++) free(defs[i].el);
                free(defs);
                return NULL;
            }
        }
    next:
        free(buf);
    }
    for (int i = 0; i < ndefs; i++) free(defs[i].el);
    free(defs);
    /* index map, parser.rs:105-125: group-2 current first, then unseen non-ground nodes in nodes() order */
    {
        /* `nodes` HashSet<String>: reuse a second string set via a scratch circuit map */
        ko_circuit seen; memset(&seen, 0, sizeof(seen));
        for (int i = 0; i < c->ne; i++) {
            ko_elem *e = &c->e[i];
            int g2 = (e->kind == KO_V || e->kind == KO_L || e->kind == KO_I) || (e->kind == KO_C && e->g2); /* elements.rs:125-140 */
            char key[KO_NAME + 8];
            if (g2) {
                snprintf(key, sizeof(key), "I(%s)", e->ident);
                if (ko_index_of(c, key) >= 0) {
                    set_err(err, KO_ERR_PARSE, errmsg, errmsg_len, "duplicate element identifier (element %d): %s", i, e->ident);
                    for (int k = 0; k < seen.n; k++) free(seen.uname[k]);
                    free(seen.uname); free(seen.hslot);
                    ko_free(c); return NULL;
                }
                map_insert(c, key);
            }
            const char *nodes[3] = { e->plus, e->minus, e->third };   /* elements.rs:69-81 */
            int nn = (e->kind == KO_NMOS) ? 3 : 2;
            for (int k = 0; k < nn; k++) {
                if (ko_index_of(&seen, nodes[k]) >= 0) continue;
                map_insert(&seen, nodes[k]);
                if (!strcmp(nodes[k], "0")) continue;
                snprintf(key, sizeof(key), "V(%s)", nodes[k]);
                map_insert(c, key);
            }
        }
        for (int k = 0; k < seen.n; k++) free(seen.uname[k]);
        free(seen.uname); free(seen.hslot);
    }
    if (c->ne == 0) { set_err(err, KO_ERR_EMPTY_NETLIST, errmsg, errmsg_len, "The netlist is empty%.0d%s", 0, ""); ko_free(c); return NULL; }
    /* second pass, parser.rs:132-149: apply models */
    for (int i = 0; i < c->ne; i++) {
        ko_elem *e = &c->e[i];
        if (e->kind == KO_D || e->kind == KO_NMOS) {
            int want = (e->kind == KO_D) ? 0 : 1, found = -1;
            for (int k = 0; k < c->nmodels; k++) if (!strcmp(c->models[k].name, e->model_name) && c->models[k].type == want) found = k;
            if (found < 0) { set_err(err, KO_ERR_MISSING_MODEL, errmsg, errmsg_len, "model not defined (element %d): %s", i, e->model_name); ko_free(c); return NULL; }
            e->is = c->models[found].is; e->n_emis = c->models[found].n_emis;
            e->vth = c->models[found].vto; e->lambda = c->models[found].lambda;
            e->beta = c->models[found].kp * (c->models[found].w / c->models[found].l);   /* models/nmosfet.rs:43-45 */
        }
    }
    c->is_current = (int *)malloc(sizeof(int) * (c->n ? c->n : 1));
    for (int i = 0; i < c->n; i++) c->is_current[i] = (c->uname[i][0] == 'I');
    for (int i = 0; i < c->ne; i++) {
        ko_elem *e = &c->e[i];
        char key[KO_NAME + 8];
        snprintf(key, sizeof(key), "V(%s)", e->plus); e->ip = ko_index_of(c, key);
        snprintf(key, sizeof(key), "V(%s)", e->minus); e->im = ko_index_of(c, key);
        snprintf(key, sizeof(key), "I(%s)", e->ident); e->ic = ko_index_of(c, key);
        e->is3 = -1;
        if (e->kind == KO_NMOS) { snprintf(key, sizeof(key), "V(%s)", e->third); e->is3 = ko_index_of(c, key); }
    }
    return c;
}

void ko_free(ko_circuit *c) {
    if (!c) return;
    for (int i = 0; i < c->n; i++) free(c->uname[i]);
    free(c->uname); free(c->e); free(c->is_current); free(c->models); free(c->hslot);
    free(c);
}
static ko_circuit *clone_circuit_shallow(const ko_circuit *c) {   /* private element table, shared names */
    ko_circuit *d = (ko_circuit *)malloc(sizeof(*d));
    *d = *c;
    d->e = (ko_elem *)malloc(sizeof(ko_elem) * c->ne);
    memcpy(d->e, c->e, sizeof(ko_elem) * c->ne);
    return d;
}
static void free_clone(ko_circuit *d) { free(d->e); free(d); }

int ko_num_unknowns(const ko_circuit *c) { return c->n; }
int ko_num_elements(const ko_circuit *c) { return c->ne; }
const char *ko_unknown_name(const ko_circuit *c, int idx) { return (idx >= 0 && idx < c->n) ? c->uname[idx] : NULL; }
int ko_element_index(const ko_circuit *c, const char *ident) {
    for (int i = 0; i < c->ne; i++) if (!strcmp(c->e[i].ident, ident)) return i;   /* dc.rs:21-24: first match */
    return -1;
}
const char *ko_element_ident(const ko_circuit *c, int e) { return c->e[e].ident; }
int ko_element_kind(const ko_circuit *c, int e) { return c->e[e].kind; }
double ko_element_value(const ko_circuit *c, int e) { return c->e[e].value; }
void ko_set_element_value(ko_circuit *c, int e, double v) { c->e[e].value = v; }
void ko_element_indices(const ko_circuit *c, int e, int *ip, int *im, int *ic) { *ip = c->e[e].ip; *im = c->e[e].im; *ic = c->e[e].ic; }
int ko_is_nonlinear(const ko_circuit *c) {   /* elements.rs:143-148 */
    for (int i = 0; i < c->ne; i++) if (c->e[i].kind == KO_D || c->e[i].kind == KO_NMOS) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * device equations
 * ---------------------------------------------------------------------------------------------- */
static const double KO_KB = 1.380649e-23, KO_Q = 1.602176634e-19, KO_TEMP = 300.0;   /* constants.rs:1-3 */
#define KO_VT (KO_KB * KO_TEMP / KO_Q)                                             /* constants.rs:4 */
#define KO_PI 3.14159265358979323846264338327950288                                /* std::f64::consts::PI */

/* a "solution map": NULL x == empty HashMap */
typedef struct { const double *x; int has_extra; double extra; } ko_map;

static double diode_vd(const ko_elem *e, const ko_map *m) {     /* diode.rs:26-40 */
    double vp = (m->x && e->ip >= 0) ? m->x[e->ip] : 0.5;
    double vm = (m->x && e->im >= 0) ? m->x[e->im] : 0.0;
    return vp - vm;
}
static double diode_limit(const ko_elem *e, double vd) {        /* diode.rs:65-70 */
    double n = e->n_emis, is = e->is;
    double vcrit = n * KO_VT * log(1.7976931348623157e308 * n * KO_VT / is);
    if (vd < -vcrit) return -vcrit;
    if (vd > vcrit) return vcrit;
    return vd;
}
static double diode_g(const ko_elem *e, const ko_map *m) {      /* diode.rs:42-48 */
    double v = diode_limit(e, diode_vd(e, m));
    double n = e->n_emis, is = e->is;
    return (is / (n * KO_VT)) * exp(v / (n * KO_VT));
}
static double diode_i(const ko_elem *e, const ko_map *m) {      /* diode.rs:50-56 */
    double v = diode_limit(e, diode_vd(e, m));
    double n = e->n_emis, is = e->is;
    return is * (exp(v / (n * KO_VT)) - 1.0);
}
static double diode_ieq(const ko_elem *e, const ko_map *m) {    /* diode.rs:58-61 */
    double vd = diode_vd(e, m);
    return diode_i(e, m) - diode_g(e, m) * vd;
}
/* nmosfet.rs:51-102: g_m, g_ds, i_d of the level-1 model, region tests and evaluation order as written there.
 * The REFERENCE never reaches these through a stamp (stampable.rs:842-844: bare node names never match the index map); they
 * are used only by the level-1 EXTENSION below (ko_set_device_models), which has no upstream oracle: tests pin it against an
 * independent Newton solve of the KCL equations (tests/test_nmos_level1.py). */
static double nmos_gm(const ko_elem *e, double vgs, double vds) {
    double vth = e->vth, beta = e->beta, lambda = e->lambda;
    if (vgs <= vth) return 0.0;
    else if (vds >= 0.0 && vds <= (vgs - vth)) return beta * vds;
    else if (vds >= (vgs - vth) && vds >= 0.0) return beta * (vgs - vth) * (1.0 + lambda * vds);
    return 0.0;
}
static double nmos_gds(const ko_elem *e, double vgs, double vds) {
    double vth = e->vth, beta = e->beta, lambda = e->lambda;
    if (vgs <= vth) return 0.0;
    else if (vds >= 0.0 && vds <= (vgs - vth)) return beta * (vgs - vth - vds);
    else if (vds >= (vgs - vth) && vds >= 0.0) return (beta / 2.0) * lambda * ((vgs - vth) * (vgs - vth));
    return 0.0;
}
static double nmos_id(const ko_elem *e, double vgs, double vds) {
    double vth = e->vth, beta = e->beta, lambda = e->lambda;
    if (vgs <= vth) return 0.0;
    else if (vds >= 0.0 && vds <= (vgs - vth)) return beta * ((vgs - vth) * vds - ((vds * vds) / 2.0));
    else if (vds >= (vgs - vth) && vds >= 0.0) return (beta / 2.0) * ((vgs - vth) * (vgs - vth)) * (1.0 + lambda * vds);
    return 0.0;
}
/* 0 = reference behaviour (NMOS stamps nothing); 1 = level-1 companion stamps (EXTENSION) */
static int ko_device_models = 0;
void ko_set_device_models(int models) { ko_device_models = models ? 1 : 0; }

static double pulse_value_at(const double *p, double time) {    /* voltage_source.rs:70-93 */
    double v0 = p[0], v1 = p[1], td = p[2], tr = p[3], tf = p[4], pw = p[5], per = p[6];
    if (time < td) return v0;
    double t = fmod(time - td, per);
    if (t < tr) return v0 + (v1 - v0) * (t / tr);
    else if (t < tr + pw) return v1;
    else if (t < tr + pw + tf) return v1 + (v0 - v1) * ((t - tr - pw) / tf);
    else return v0;
}
static double sin_value_at(const double *s, double time) {      /* voltage_source.rs:52-65 */
    double off = s[0], amp = s[1], freq = s[2], delay = s[3], damp = s[4], phase = s[5];
    if (time < delay) return off;
    double w = 2.0 * KO_PI * freq;
    double ph = phase * (KO_PI / 180.0);   /* f64::to_radians */
    double t = time - delay;
    return off + amp * exp(-damp * t) * sin(w * t + ph);
}
static double vsrc_transient_value(const ko_elem *e, double time) {   /* voltage_source.rs:244-252 */
    if (e->has_pulse) return pulse_value_at(e->pulse, time);
    if (e->has_sin) return sin_value_at(e->sinp, time);
    return e->value;
}
double ko_transient_value_at(const ko_circuit *c, int e, double time) { return vsrc_transient_value(&c->e[e], time); }

/* ------------------------------------------------------------------------------------------------
 * stamps (crates/krets-solver/src/stampable.rs).  Triplet lists in element order.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int *r, *c; double *re, *im; int n, cap; int cplx; } ko_trip;
static void trip_init(ko_trip *t, int cplx) { memset(t, 0, sizeof(*t)); t->cplx = cplx; }
static void trip_free(ko_trip *t) { free(t->r); free(t->c); free(t->re); free(t->im); }
static void trip_push(ko_trip *t, int r, int c, double re, double im) {
    if (t->n == t->cap) {
        t->cap = t->cap ? 2 * t->cap : 256;
        t->r = (int *)realloc(t->r, sizeof(int) * t->cap);
        t->c = (int *)realloc(t->c, sizeof(int) * t->cap);
        t->re = (double *)realloc(t->re, sizeof(double) * t->cap);
        if (t->cplx) t->im = (double *)realloc(t->im, sizeof(double) * t->cap);
    }
    t->r[t->n] = r; t->c[t->n] = c; t->re[t->n] = re;
    if (t->cplx) t->im[t->n] = im;
    t->n++;
}
/* the group-1 two-terminal conductance pattern: (p,p)+y (m,m)+y (p,m)-y (m,p)-y, stampable.rs:229-244 etc. */
static void stamp_g1(ko_trip *g, int ip, int im, double re, double imag) {
    if (ip >= 0) trip_push(g, ip, ip, re, imag);
    if (im >= 0) trip_push(g, im, im, re, imag);
    if (ip >= 0 && im >= 0) { trip_push(g, ip, im, -re, -imag); trip_push(g, im, ip, -re, -imag); }
}

enum { MODE_DC = 0, MODE_AC = 1, MODE_TRAN = 2, MODE_TRAP = 3 };   /* MODE_TRAP: trapezoidal companions — NOT in the reference (no oracle upstream) */

/* One element's matrix + RHS contribution.
 *   DC   : guess = NR guess map
 *   AC   : guess = OP solution, w = frequency
 *   TRAN : guess = NR guess, prev = previous time point, w = h */
/* per-capacitor current state of the trapezoidal rule (indexed by element ordinal), set by the TRAP driver */
static __thread const double *ko_trap_iprev = NULL;
static __thread int ko_trap_elem = 0;
static __thread double ko_trap_tnow = 0.0;

static int stamp_element(const ko_elem *e, int mode, const ko_map *guess, const ko_map *prev, double w, ko_trip *g, ko_trip *b) {
    int ip = e->ip, im = e->im, ic = e->ic;
    switch (e->kind) {
    case KO_R: {                                             /* stampable.rs:202-317 (g2 is never set: resistor.rs:47) */
        double gg = 1.0 / e->value;
        stamp_g1(g, ip, im, gg, 0.0);
        break;
    }
    case KO_C:
        if (mode == MODE_DC) break;                          /* :361-368, :423-430 */
        if (mode == MODE_TRAP) {                             /* i_n = (2C/h)(v_n - v_{n-1}) - i_{n-1}  (north-star extension) */
            double gg = e->value / (w / 2.0);
            stamp_g1(g, ip, im, gg, 0.0);
            double vp = (prev->x && ip >= 0) ? prev->x[ip] : 0.0;
            double vm = (prev->x && im >= 0) ? prev->x[im] : 0.0;
            double hist = gg * (vp - vm) + ko_trap_iprev[ko_trap_elem];
            if (ip >= 0) trip_push(b, ip, 0, hist, 0.0);
            if (im >= 0) trip_push(b, im, 0, -hist, 0.0);
            break;
        }
        if (mode == MODE_AC) {                               /* :370-421 */
            double y = 2.0 * KO_PI * w * e->value;
            if (!e->g2) stamp_g1(g, ip, im, 0.0, y);
            else {
                if (ip >= 0 && ic >= 0) trip_push(g, ic, ip, -0.0, -y);
                if (im >= 0 && ic >= 0) trip_push(g, ic, im, 0.0, y);
                if (ic >= 0) trip_push(g, ic, ic, 1.0, 0.0);
            }
        } else {                                             /* :442-505 */
            double gg = e->value / w;
            stamp_g1(g, ip, im, gg, 0.0);
            double vp = (prev->x && ip >= 0) ? prev->x[ip] : 0.0;
            double vm = (prev->x && im >= 0) ? prev->x[im] : 0.0;
            double vprev = vp - vm;
            double ieq = -(e->value / w) * vprev;
            if (ip >= 0) trip_push(b, ip, 0, -ieq, 0.0);
            if (im >= 0) trip_push(b, im, 0, ieq, 0.0);
        }
        break;
    case KO_I:
        if (mode == MODE_AC) {                               /* :532-559 — no (ic,ic) entry */
            if (ip >= 0 && ic >= 0) trip_push(g, ip, ic, 1.0, 0.0);
            if (im >= 0 && ic >= 0) trip_push(g, im, ic, -1.0, 0.0);
        } else {                                             /* :509-530 */
            if (ic >= 0) trip_push(g, ic, ic, 1.0, 0.0);
            if (ip >= 0 && ic >= 0) trip_push(g, ip, ic, 1.0, 0.0);
            if (im >= 0 && ic >= 0) trip_push(g, im, ic, -1.0, 0.0);
        }
        if (ic >= 0) trip_push(b, ic, 0, e->value, 0.0);     /* :561-584 */
        break;
    case KO_D: {
        if (mode == MODE_AC) {                               /* :614-641, no RHS :664-673 */
            stamp_g1(g, ip, im, diode_g(e, guess), 0.0);
        } else {                                             /* :588-612, :643-662 */
            stamp_g1(g, ip, im, diode_g(e, guess), 0.0);
            double ieq = diode_ieq(e, guess);
            if (ip >= 0) trip_push(b, ip, 0, -ieq, 0.0);
            if (im >= 0) trip_push(b, im, 0, ieq, 0.0);
        }
        break;
    }
    case KO_L:
        if (mode == MODE_DC) {                               /* :677-704 */
            if (ip >= 0 && ic >= 0) trip_push(g, ip, ic, 1.0, 0.0);
            if (im >= 0 && ic >= 0) trip_push(g, im, ic, -1.0, 0.0);
            if (ip >= 0 && ic >= 0) trip_push(g, ic, ip, 1.0, 0.0);
            if (im >= 0 && ic >= 0) trip_push(g, ic, im, -1.0, 0.0);
        } else if (mode == MODE_AC) {                        /* :706-746 */
            double z = 2.0 * KO_PI * w * e->value;
            if (ip >= 0 && ic >= 0) trip_push(g, ip, ic, 1.0, 0.0);
            if (im >= 0 && ic >= 0) trip_push(g, im, ic, -1.0, 0.0);
            if (ip >= 0 && ic >= 0) trip_push(g, ic, ip, 1.0, 0.0);
            if (im >= 0 && ic >= 0) trip_push(g, ic, im, -1.0, 0.0);
            if (ic >= 0) trip_push(g, ic, ic, -0.0, -z);
        } else if (mode == MODE_TRAP) {                      /* v_n - (2L/h) i_n = -(2L/h) i_{n-1} - v_{n-1} */
            double zz = e->value / (w / 2.0);
            if (ic >= 0) trip_push(g, ic, ic, -zz, 0.0);
            if (ip >= 0 && ic >= 0) { trip_push(g, ip, ic, 1.0, 0.0); trip_push(g, ic, ip, 1.0, 0.0); }
            if (im >= 0 && ic >= 0) { trip_push(g, im, ic, -1.0, 0.0); trip_push(g, ic, im, -1.0, 0.0); }
            double vp = ip >= 0 ? prev->x[ip] : 0.0, vm = im >= 0 ? prev->x[im] : 0.0;
            if (ic >= 0) { trip_push(b, ic, 0, -(zz * prev->x[ic]), 0.0); trip_push(b, ic, 0, -(vp - vm), 0.0); }
        } else {                                             /* :767-816 */
            if (ic >= 0) trip_push(g, ic, ic, -e->value / w, 0.0);
            if (ip >= 0 && ic >= 0) { trip_push(g, ip, ic, 1.0, 0.0); trip_push(g, ic, ip, 1.0, 0.0); }
            if (im >= 0 && ic >= 0) { trip_push(g, im, ic, -1.0, 0.0); trip_push(g, ic, im, -1.0, 0.0); }
            double iprev = prev->x[ic];                      /* :806-809 unwrap(): always present */
            if (ic >= 0) trip_push(b, ic, 0, -(e->value / w) * iprev, 0.0);
        }
        break;
    case KO_NMOS:
        if (ko_device_models) {
            /* EXTENSION: level-1 companion at the linearisation point (voltages default to 0.0, :825-833); drain d, gate gt, source s.
             * KCL-consistent signs: i_d enters the drain and leaves the source (the unreachable reference stamp :846-868 gives the
             * source row the drain row's signs).  AC: small-signal g_m / g_ds at the operating point, no excitation (:897-912). */
            int d = ip, gt = im, s = e->is3;
            double vg = (guess->x && gt >= 0) ? guess->x[gt] : 0.0, vs = (guess->x && s >= 0) ? guess->x[s] : 0.0;
            double vdn = (guess->x && d >= 0) ? guess->x[d] : 0.0;
            double vgs = vg - vs, vds = vdn - vs;
            double gm = nmos_gm(e, vgs, vds), gds = nmos_gds(e, vgs, vds);
            if (d >= 0) trip_push(g, d, d, gds, 0.0);
            if (s >= 0) trip_push(g, s, s, gds + gm, 0.0);
            if (d >= 0 && s >= 0) { trip_push(g, d, s, -(gds + gm), 0.0); trip_push(g, s, d, -gds, 0.0); }
            if (s >= 0 && gt >= 0) trip_push(g, s, gt, -gm, 0.0);
            if (d >= 0 && gt >= 0) trip_push(g, d, gt, gm, 0.0);
            if (mode != MODE_AC) {
                double ieq = nmos_id(e, vgs, vds) - gds * vds - gm * vgs;   /* :889 */
                if (s >= 0) trip_push(b, s, 0, ieq, 0.0);
                if (d >= 0) trip_push(b, d, 0, -ieq, 0.0);
            }
            break;
        }
        if (mode == MODE_AC) return KO_ERR_UNSUPPORTED;      /* :914-921 todo!() */
        break;                                               /* :842-844,895,899: lookups never match -> no stamps */
    case KO_V:                                               /* :960-1050 */
        if (ip >= 0 && ic >= 0) { trip_push(g, ip, ic, 1.0, 0.0); trip_push(g, ic, ip, 1.0, 0.0); }
        if (im >= 0 && ic >= 0) { trip_push(g, im, ic, -1.0, 0.0); trip_push(g, ic, im, -1.0, 0.0); }
        if (ic >= 0) {
            if (mode == MODE_DC) trip_push(b, ic, 0, e->value, 0.0);
            else if (mode == MODE_AC) trip_push(b, ic, 0, e->ac_amplitude, 0.0);
            else if (mode == MODE_TRAP) trip_push(b, ic, 0, vsrc_transient_value(e, ko_trap_tnow), 0.0);   /* source at t_n */
            else {
                double t = guess->has_extra ? guess->extra : 0.0;   /* :1043 solution_map.get("time") */
                trip_push(b, ic, 0, vsrc_transient_value(e, t), 0.0);
            }
        }
        break;
    default: return KO_ERR_UNSUPPORTED;
    }
    return KO_OK;
}

/* solver.rs:59-70 sum_triplets: HashMap<(row,col),N>, accumulation in input order from N::default() */
typedef struct { int64_t *key; int *idx; int cap; } ko_hash;
static void sum_triplets(const ko_trip *in, int ncols, ko_trip *out) {
    trip_init(out, in->cplx);
    int cap = 16;
    while (cap < 2 * in->n + 2) cap <<= 1;
    int64_t *key = (int64_t *)malloc(sizeof(int64_t) * cap);
    int *idx = (int *)malloc(sizeof(int) * cap);
    for (int i = 0; i < cap; i++) key[i] = -1;
    for (int t = 0; t < in->n; t++) {
        int64_t k = (int64_t)in->r[t] * (int64_t)(ncols > 0 ? ncols : 1) + in->c[t];
        uint64_t h = ((uint64_t)k * 0x9E3779B97F4A7C15ULL) >> 17;
        h &= (uint64_t)(cap - 1);
        while (key[h] >= 0 && key[h] != k) h = (h + 1) & (uint64_t)(cap - 1);
        if (key[h] < 0) {
            key[h] = k;
            idx[h] = out->n;
            trip_push(out, in->r[t], in->c[t], 0.0, 0.0);     /* or_default() */
        }
        out->re[idx[h]] += in->re[t];
        if (in->cplx) out->im[idx[h]] += in->im[t];
    }
    free(key); free(idx);
}

/* SparseColMat::try_new_from_triplets: CSC with sorted rows; explicit zeros kept */
static void build_csc(const ko_trip *t, int n, int *Ap, int *Ai, double *Are, double *Aim) {
    for (int j = 0; j <= n; j++) Ap[j] = 0;
    for (int k = 0; k < t->n; k++) Ap[t->c[k] + 1]++;
    for (int j = 0; j < n; j++) Ap[j + 1] += Ap[j];
    int *w = (int *)malloc(sizeof(int) * (n + 1));
    memcpy(w, Ap, sizeof(int) * (n + 1));
    for (int k = 0; k < t->n; k++) {
        int q = w[t->c[k]]++;
        Ai[q] = t->r[k]; Are[q] = t->re[k];
        if (Aim) Aim[q] = t->im[k];
    }
    free(w);
    for (int j = 0; j < n; j++)                              /* insertion sort rows within each column */
        for (int p = Ap[j] + 1; p < Ap[j + 1]; p++) {
            int r = Ai[p]; double re = Are[p], im = Aim ? Aim[p] : 0.0;
            int q = p - 1;
            while (q >= Ap[j] && Ai[q] > r) { Ai[q + 1] = Ai[q]; Are[q + 1] = Are[q]; if (Aim) Aim[q + 1] = Aim[q]; q--; }
            Ai[q + 1] = r; Are[q + 1] = re; if (Aim) Aim[q + 1] = im;
        }
}

/* assemble every element's stamps for one solve point */
static int assemble(const ko_circuit *c, int mode, const ko_map *guess, const ko_map *prev, double w, ko_trip *g, ko_trip *b) {
    for (int i = 0; i < c->ne; i++) {
        const ko_elem *e = &c->e[i];
        ko_trap_elem = i;
        if (mode == MODE_DC && e->kind == KO_C) continue;    /* op.rs:20-24, dc.rs:60-64 */
        int rc = stamp_element(e, mode, guess, prev, w, g, b);
        if (rc) return rc;
    }
    return KO_OK;
}

/* One "solve point" exactly as op.rs:36-58 does it: stamp, sum, CSC, LU, b, solve (real). */
static int solve_point_real(const ko_circuit *c, int mode, const ko_map *guess, const ko_map *prev, double w, double *x) {
    int n = c->n;
    ko_trip g, b, gs, bs;
    trip_init(&g, 0); trip_init(&b, 0);
    int rc = assemble(c, mode, guess, prev, w, &g, &b);
    if (rc) { trip_free(&g); trip_free(&b); return rc; }
    sum_triplets(&g, n, &gs);
    sum_triplets(&b, 1, &bs);
    int *Ap = (int *)malloc(sizeof(int) * (n + 1));
    int *Ai = (int *)malloc(sizeof(int) * (gs.n + 1));
    double *Ax = (double *)malloc(sizeof(double) * (gs.n + 1));
    build_csc(&gs, n, Ap, Ai, Ax, NULL);
    ko_lu_r f;
    rc = ko_lu_factor_r(n, Ap, Ai, Ax, &f) ? KO_ERR_DECOMPOSITION : KO_OK;
    if (!rc) {
        double *rhs = (double *)calloc(n ? n : 1, sizeof(double));      /* Mat::zeros */
        for (int k = 0; k < bs.n; k++) rhs[bs.r[k]] = bs.re[k];          /* op.rs:54-56 assignment */
        ko_lu_solve_r(&f, rhs, x);
        free(rhs);
        ko_lu_free_r(&f);
    }
    free(Ap); free(Ai); free(Ax);
    trip_free(&g); trip_free(&b); trip_free(&gs); trip_free(&bs);
    return rc;
}
static int solve_point_cplx(const ko_circuit *c, const ko_map *op, double freq, ko_c64 *x) {
    int n = c->n;
    ko_trip g, b, gs, bs;
    trip_init(&g, 1); trip_init(&b, 1);
    int rc = assemble(c, MODE_AC, op, NULL, freq, &g, &b);
    if (rc) { trip_free(&g); trip_free(&b); return rc; }
    sum_triplets(&g, n, &gs);
    sum_triplets(&b, 1, &bs);
    int *Ap = (int *)malloc(sizeof(int) * (n + 1));
    int *Ai = (int *)malloc(sizeof(int) * (gs.n + 1));
    double *Are = (double *)malloc(sizeof(double) * (gs.n + 1));
    double *Aim = (double *)malloc(sizeof(double) * (gs.n + 1));
    build_csc(&gs, n, Ap, Ai, Are, Aim);
    ko_c64 *Ax = (ko_c64 *)malloc(sizeof(ko_c64) * (gs.n + 1));
    for (int k = 0; k < gs.n; k++) { Ax[k].re = Are[k]; Ax[k].im = Aim[k]; }
    ko_lu_c f;
    rc = ko_lu_factor_c(n, Ap, Ai, Ax, &f) ? KO_ERR_DECOMPOSITION : KO_OK;
    if (!rc) {
        ko_c64 *rhs = (ko_c64 *)calloc(n ? n : 1, sizeof(ko_c64));
        for (int k = 0; k < bs.n; k++) { rhs[bs.r[k]].re = bs.re[k]; rhs[bs.r[k]].im = bs.im[k]; }
        ko_lu_solve_c(&f, rhs, x);
        free(rhs);
        ko_lu_free_c(&f);
    }
    free(Ap); free(Ai); free(Are); free(Aim); free(Ax);
    trip_free(&g); trip_free(&b); trip_free(&gs); trip_free(&bs);
    return rc;
}

int ko_stamp_matrix(const ko_circuit *c, int mode, double freq_or_h, const double *guess, int *colptr, int *rowidx, double *vals, int cap) {
    int n = c->n;
    ko_trip g, b, gs;
    int cplx = (mode == MODE_AC);
    trip_init(&g, cplx); trip_init(&b, cplx);
    ko_map gm = { guess, 0, 0.0 };
    double *zeros = (double *)calloc(n ? n : 1, sizeof(double));
    ko_map pm = { zeros, 1, 0.0 };
    int rc = assemble(c, mode, &gm, &pm, freq_or_h, &g, &b);
    if (rc) { trip_free(&g); trip_free(&b); free(zeros); return -1000000 - rc; }
    sum_triplets(&g, n, &gs);
    int nnz = gs.n;
    if (nnz > cap) { trip_free(&g); trip_free(&b); trip_free(&gs); free(zeros); return -nnz; }
    double *re = (double *)malloc(sizeof(double) * (nnz + 1)), *im = cplx ? (double *)malloc(sizeof(double) * (nnz + 1)) : NULL;
    build_csc(&gs, n, colptr, rowidx, re, im);
    for (int k = 0; k < nnz; k++) { if (cplx) { vals[2 * k] = re[k]; vals[2 * k + 1] = im[k]; } else vals[k] = re[k]; }
    free(re); free(im); free(zeros);
    trip_free(&g); trip_free(&b); trip_free(&gs);
    return nnz;
}

/* solver.rs:76-104 convergence_check, including the pseudo-unknown ("step"/"time") compared as a voltage */
static int convergence_check(const ko_circuit *c, const ko_config *cfg, const ko_map *prev, const ko_map *res) {
    if (!prev->x && !prev->has_extra) return 0;              /* previous_result.is_empty() */
    for (int i = 0; i < c->n; i++) {
        double value = res->x[i];
        double pv = prev->x ? prev->x[i] : 0.0;
        double diff = fabs(value - pv);
        double scale = fmax(fabs(value), fabs(pv));
        double atol = c->is_current[i] ? cfg->current_absolute_tolerance : cfg->voltage_absolute_tolerance;
        if (!(diff <= cfg->relative_tolerance * scale + atol)) return 0;
    }
    if (res->has_extra) {
        double value = res->extra;
        double pv = prev->has_extra ? prev->extra : 0.0;
        double diff = fabs(value - pv);
        double scale = fmax(fabs(value), fabs(pv));
        if (!(diff <= cfg->relative_tolerance * scale + cfg->voltage_absolute_tolerance)) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * analyses
 * ---------------------------------------------------------------------------------------------- */
/* op.rs:15-86 */
int ko_solve_op(const ko_circuit *c, const ko_config *cfg, double *x, int *iterations) {
    int n = c->n;
    int nonlinear = ko_is_nonlinear(c);
    double *prevx = (double *)malloc(sizeof(double) * (n ? n : 1));
    ko_map prev = { NULL, 0, 0.0 };
    int rc = KO_OK, it = 0;
    for (int iter = 0; iter < cfg->maximum_iterations; iter++) {
        it = iter + 1;
        rc = solve_point_real(c, MODE_DC, &prev, NULL, 0.0, x);
        if (rc) break;
        if (!nonlinear) break;
        ko_map res = { x, 0, 0.0 };
        if (convergence_check(c, cfg, &prev, &res)) break;
        memcpy(prevx, x, sizeof(double) * n);
        prev.x = prevx;
        if (iter == cfg->maximum_iterations - 1) { rc = KO_ERR_MAX_ITERATIONS; break; }
    }
    if (iterations) *iterations = it;
    free(prevx);
    return rc;
}

int ko_dc_num_steps(double start, double stop, double step) {   /* dc.rs:44-45: `as usize` saturating truncation */
    double v = fabs((stop - start) / step);
    if (!(v == v)) return 1;            /* NaN as usize == 0 */
    if (v > 2147483000.0) return 2147483647;
    return (int)v + 1;
}

/* dc.rs:47-113 body for points [i0, i1) with a private circuit copy */
static int dc_range(ko_circuit *c, const ko_config *cfg, int ei, double start, double step, int i0, int i1, double *out, int warm_start) {
    int n = c->n;
    int nonlinear = ko_is_nonlinear(c);
    double *last = (double *)malloc(sizeof(double) * (n ? n : 1));
    double *prevx = (double *)malloc(sizeof(double) * (n ? n : 1));
    int have_last = 0; double last_step = 0.0;
    int rc = KO_OK;
    for (int i = i0; i < i1 && !rc; i++) {
        c->e[ei].value = start + ((double)i * step);             /* dc.rs:48-55 */
        ko_map prev = { NULL, 0, 0.0 };
        if (warm_start && have_last) { memcpy(prevx, last, sizeof(double) * n); prev.x = prevx; prev.has_extra = 1; prev.extra = last_step; }
        double *x = out + (size_t)i * n;
        for (int iter = 0; iter < cfg->maximum_iterations; iter++) {
            rc = solve_point_real(c, MODE_DC, &prev, NULL, 0.0, x);
            if (rc) break;
            ko_map res = { x, 1, (double)i };                    /* dc.rs:97 "step" */
            if (!nonlinear) break;
            if (convergence_check(c, cfg, &prev, &res)) break;
            memcpy(prevx, x, sizeof(double) * n);
            prev.x = prevx; prev.has_extra = 1; prev.extra = (double)i;
            if (iter == cfg->maximum_iterations - 1) { rc = KO_ERR_MAX_ITERATIONS; break; }
        }
        memcpy(last, x, sizeof(double) * n); have_last = 1; last_step = (double)i;
    }
    free(last); free(prevx);
    return rc;
}

/* cpu_best_effort: what a tuned CPU implementation of the same path does for a LINEAR circuit — the stamped
 * pattern, the triplet -> CSC slot map and all buffers are built once; every point re-stamps the values (same push order, so the
 * per-entry sums are bit-identical to sum_triplets), scatters them into the fixed CSC, runs the numeric LU and solves.  Points are
 * split over `nthreads` OpenMP threads.  A reported baseline, not the reference's work structure (that is ko_solve_dc). */
int ko_solve_dc_best_effort(ko_circuit *c, const char *element, double start, double step, int num, double *out, int nthreads) {
    int ei = ko_element_index(c, element);
    if (ei < 0) return KO_ERR_ELEMENT_NOT_FOUND;
    if (c->e[ei].kind != KO_V && c->e[ei].kind != KO_I) return KO_ERR_INVALID_ELEMENT_FORMAT;
    if (ko_is_nonlinear(c)) return KO_ERR_UNSUPPORTED;
    const int n = c->n;
    ko_map none = { NULL, 0, 0.0 };
    /* one-time analysis */
    ko_trip g0, b0, gs;
    trip_init(&g0, 0); trip_init(&b0, 0);
    int rc = assemble(c, MODE_DC, &none, NULL, 0.0, &g0, &b0);
    if (rc) { trip_free(&g0); trip_free(&b0); return rc; }
    sum_triplets(&g0, n, &gs);
    const int nnz = gs.n;
    int *Ap = (int *)malloc(sizeof(int) * (n + 1)), *Ai = (int *)malloc(sizeof(int) * (nnz + 1));
    double *A0 = (double *)malloc(sizeof(double) * (nnz + 1));
    build_csc(&gs, n, Ap, Ai, A0, NULL);
    int *slot = (int *)malloc(sizeof(int) * (g0.n + 1));
    for (int t = 0; t < g0.n; t++) {
        int lo = Ap[g0.c[t]], hi = Ap[g0.c[t] + 1] - 1;
        while (lo < hi) { int mid = (lo + hi) / 2; if (Ai[mid] < g0.r[t]) lo = mid + 1; else hi = mid; }
        slot[t] = lo;
    }
    const int ntrip = g0.n;
    trip_free(&g0); trip_free(&b0); trip_free(&gs); free(A0);
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
#endif
    {
        int tid = 0, nt = 1;
#ifdef _OPENMP
        tid = omp_get_thread_num(); nt = omp_get_num_threads();
#endif
        int i0 = (int)((int64_t)num * tid / nt), i1 = (int)((int64_t)num * (tid + 1) / nt);
        ko_circuit *d = clone_circuit_shallow(c);
        ko_trip g, b;
        trip_init(&g, 0); trip_init(&b, 0);
        double *Ax = (double *)malloc(sizeof(double) * (nnz + 1)), *rhs = (double *)malloc(sizeof(double) * (n ? n : 1));
        int r = KO_OK;
        for (int i = i0; i < i1 && !r; i++) {
            d->e[ei].value = start + ((double)i * step);
            g.n = 0; b.n = 0;
            r = assemble(d, MODE_DC, &none, NULL, 0.0, &g, &b);
            if (!r && g.n != ntrip) r = KO_ERR_UNSUPPORTED;   /* stamp structure changed between points: not a linear sweep */
            if (r) break;
            memset(Ax, 0, sizeof(double) * nnz);
            for (int t = 0; t < ntrip; t++) Ax[slot[t]] += g.re[t];
            memset(rhs, 0, sizeof(double) * n);
            for (int t = 0; t < b.n; t++) rhs[b.r[t]] += b.re[t];
            ko_lu_r f;
            if (ko_lu_factor_r(n, Ap, Ai, Ax, &f)) { r = KO_ERR_DECOMPOSITION; break; }
            ko_lu_solve_r(&f, rhs, out + (size_t)i * n);
            ko_lu_free_r(&f);
        }
        trip_free(&g); trip_free(&b); free(Ax); free(rhs); free_clone(d);
        if (r) {
#ifdef _OPENMP
#pragma omp critical
#endif
            rc = r;
        }
    }
    free(Ap); free(Ai); free(slot);
    return rc;
}

int ko_solve_dc(ko_circuit *c, const ko_config *cfg, const char *element, double start, double stop, double step, double *out,
                int cap_points, int *npoints, int warm_start, int nthreads) {
    int ei = ko_element_index(c, element);
    if (ei < 0) return KO_ERR_ELEMENT_NOT_FOUND;
    if (c->e[ei].kind != KO_V && c->e[ei].kind != KO_I) return KO_ERR_INVALID_ELEMENT_FORMAT;
    int num = ko_dc_num_steps(start, stop, step);
    if (npoints) *npoints = num;
    if (num > cap_points) return KO_ERR_BUFFER;
    int rc = KO_OK;
    if (nthreads <= 1) {
        ko_circuit *d = clone_circuit_shallow(c);               /* restores the swept value (dc.rs:117-121) by never touching c */
        rc = dc_range(d, cfg, ei, start, step, 0, num, out, warm_start);
        free_clone(d);
    } else {
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
        {
            int tid = 0, nt = 1;
#ifdef _OPENMP
            tid = omp_get_thread_num(); nt = omp_get_num_threads();
#endif
            int i0 = (int)((int64_t)num * tid / nt), i1 = (int)((int64_t)num * (tid + 1) / nt);
            ko_circuit *d = clone_circuit_shallow(c);
            int r = dc_range(d, cfg, ei, start, step, i0, i1, out, warm_start);
            free_clone(d);
            if (r) {
#ifdef _OPENMP
#pragma omp critical
#endif
                rc = r;
            }
        }
    }
    return rc;
}

/* analyses.rs:80-158 */
static int last_lt(const double *f, int n, double bound) { return n == 0 || f[n - 1] < bound; }   /* is_none_or(|&f| f < bound) */
static int cmp_f64(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x < y) ? -1 : (x > y) ? 1 : 0; }
int ko_generate_frequencies(int kind, unsigned points, double fstart, double fstop, double *out, int cap) {
    int n = 0;
#define PUSH(v) do { if (n >= cap) return -1; out[n++] = (v); } while (0)
    if (fstart <= 0.0 || fstop <= 0.0 || fstart > fstop) return 0;
    if (kind == 0) {
        if (points == 1) PUSH(fstart);
        else if (points > 1) {
            double step = (fstop - fstart) / (double)(points - 1);
            for (unsigned i = 0; i < points; i++) PUSH(fstart + (double)i * step);
        }
    } else {
        if (points == 0) return 0;
        double num = (kind == 1) ? log10(fstop / fstart) : log2(fstop / fstart);
        double tp = round(num * (double)points);
        unsigned total = (tp >= 4294967295.0) ? 4294967295u : (tp <= 0.0 ? 0u : (unsigned)tp);   /* `as u32` saturates */
        total = (total == 4294967295u) ? total : total + 1;
        double factor = pow(kind == 1 ? 10.0 : 2.0, 1.0 / (double)points);
        double cur = fstart;
        for (unsigned i = 0; i < total; i++) {
            if (cur > fstop * (1.0 + 1e-9)) break;
            PUSH(cur);
            cur *= factor;
            if (cur > fstop && last_lt(out, n, fstop)) { PUSH(fstop); break; }
        }
        if (last_lt(out, n, fstop * (1.0 - 1e-9))) PUSH(fstop);
    }
#undef PUSH
    qsort(out, n, sizeof(double), cmp_f64);      /* sort_by(partial_cmp) — stable vs unstable is irrelevant for equal f64 */
    int m = 0;
    for (int i = 0; i < n; i++) if (m == 0 || out[i] != out[m - 1]) out[m++] = out[i];   /* dedup */
    return m;
}

/* ac.rs:17-109 */
int ko_solve_ac(const ko_circuit *c, const ko_config *cfg, const double *freqs, int nf, double *out, int *nout, int nthreads) {
    int n = c->n;
    double *op = (double *)malloc(sizeof(double) * (n ? n : 1));
    int rc = ko_solve_op(c, cfg, op, NULL);                  /* ac.rs:25 */
    if (rc) { free(op); return rc; }
    ko_map opm = { op, 0, 0.0 };
    /* rows are emitted only for f > 0 (ac.rs:40-45) */
    int *row = (int *)malloc(sizeof(int) * (nf ? nf : 1));
    int m = 0;
    for (int i = 0; i < nf; i++) row[i] = (freqs[i] <= 0.0) ? -1 : m++;
    if (nout) *nout = m;
    int err = KO_OK;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
#endif
    for (int i = 0; i < nf; i++) {
        if (row[i] < 0) continue;
        int r = solve_point_cplx(c, &opm, freqs[i], (ko_c64 *)(out + (size_t)row[i] * n * 2));
        if (r) {
#ifdef _OPENMP
#pragma omp critical
#endif
            err = r;
        }
    }
    free(op); free(row);
    return err;
}

int ko_tran_num_steps(double h, double stop) {   /* transient.rs:29 */
    double v = round(stop / h);
    if (!(v == v) || v < 0) return 0;
    if (v > 2147483000.0) return 2147483647;
    return (int)v;
}

/* transient.rs:14-114 */
int ko_solve_tran(const ko_circuit *c, const ko_config *cfg, double h, double stop, double *out, double *time, int cap_rows, int *nrows) {
    int n = c->n;
    int nsteps = ko_tran_num_steps(h, stop);
    if (nrows) *nrows = nsteps + 1;
    if (nsteps + 1 > cap_rows) return KO_ERR_BUFFER;
    int rc = ko_solve_op(c, cfg, out, NULL);                  /* transient.rs:21 */
    if (rc) return rc;
    time[0] = 0.0;
    int nonlinear = ko_is_nonlinear(c);
    double *guess = (double *)malloc(sizeof(double) * (n ? n : 1));
    for (int step = 1; step <= nsteps && !rc; step++) {
        double t = (double)step * h;                         /* :44 */
        const double *prevx = out + (size_t)(step - 1) * n;
        ko_map prev = { prevx, 1, time[step - 1] };
        memcpy(guess, prevx, sizeof(double) * n);
        ko_map g = { guess, 1, time[step - 1] };             /* :49 clone of prev row — carries its "time" */
        double *x = out + (size_t)step * n;
        for (int iter = 0; iter < cfg->maximum_iterations; iter++) {
            rc = solve_point_real(c, MODE_TRAN, &g, &prev, h, x);
            if (rc) break;
            ko_map res = { x, 1, t };                        /* :95 */
            if (!nonlinear) break;
            if (convergence_check(c, cfg, &g, &res)) break;
            memcpy(guess, x, sizeof(double) * n);
            g.extra = t;
            if (iter == cfg->maximum_iterations - 1) { rc = KO_ERR_MAX_ITERATIONS; break; }
        }
        time[step] = t;
    }
    free(guess);
    return rc;
}

/* Fixed-step transient with a selectable integrator.  integrator 0 = ko_solve_tran (reference backward Euler with the
 * source lag); 1 = trapezoidal; 2 = backward Euler with the source evaluated at t_n.  1 and 2 are NOT reference
 * behaviour: they restate the north-star extension so that the CUDA path has something to be checked against, and are
 * themselves validated against closed-form RC / RL responses in tests/. */
int ko_solve_tran_ex(const ko_circuit *c, const ko_config *cfg, double h, double stop, int integrator, double *out, double *time,
                     int cap_rows, int *nrows) {
    if (integrator == 0) return ko_solve_tran(c, cfg, h, stop, out, time, cap_rows, nrows);
    int n = c->n;
    int nsteps = ko_tran_num_steps(h, stop);
    if (nrows) *nrows = nsteps + 1;
    if (nsteps + 1 > cap_rows) return KO_ERR_BUFFER;
    int rc = ko_solve_op(c, cfg, out, NULL);
    if (rc) return rc;
    time[0] = 0.0;
    int nonlinear = ko_is_nonlinear(c);
    double *guess = (double *)malloc(sizeof(double) * (n ? n : 1));
    double *iprev = (double *)calloc(c->ne ? c->ne : 1, sizeof(double));   /* capacitor currents, 0 at the DC operating point */
    ko_trap_iprev = iprev;
    for (int step = 1; step <= nsteps && !rc; step++) {
        double t = (double)step * h;
        const double *prevx = out + (size_t)(step - 1) * n;
        ko_map prev = { prevx, 1, time[step - 1] };
        memcpy(guess, prevx, sizeof(double) * n);
        ko_map g = { guess, 1, integrator == 2 ? t : time[step - 1] };
        ko_trap_tnow = t;
        double *x = out + (size_t)step * n;
        for (int iter = 0; iter < cfg->maximum_iterations; iter++) {
            rc = solve_point_real(c, integrator == 1 ? MODE_TRAP : MODE_TRAN, &g, &prev, h, x);
            if (rc) break;
            ko_map res = { x, 1, t };
            if (!nonlinear) break;
            ko_map gchk = { guess, 1, iter == 0 ? time[step - 1] : t };
            if (convergence_check(c, cfg, &gchk, &res)) break;
            memcpy(guess, x, sizeof(double) * n);
            g.extra = t;
            if (iter == cfg->maximum_iterations - 1) { rc = KO_ERR_MAX_ITERATIONS; break; }
        }
        if (integrator == 1)
            for (int i = 0; i < c->ne; i++) {
                const ko_elem *e = &c->e[i];
                if (e->kind != KO_C) continue;
                double gg = e->value / (h / 2.0);
                double vn = (e->ip >= 0 ? x[e->ip] : 0.0) - (e->im >= 0 ? x[e->im] : 0.0);
                double vp = (e->ip >= 0 ? prevx[e->ip] : 0.0) - (e->im >= 0 ? prevx[e->im] : 0.0);
                iprev[i] = gg * (vn - vp) - iprev[i];
            }
        time[step] = t;
    }
    ko_trap_iprev = NULL;
    free(guess); free(iprev);
    return rc;
}

const char *ko_version(void) { return "krets-oracle 1 (restated krets semantics; LU parity unpinned)"; }
