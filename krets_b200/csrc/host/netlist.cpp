#include "netlist.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <unordered_set>

namespace krh {

namespace {

// A tiny cursor over one netlist line, mirroring the nom combinators the reference composes.
struct Cursor {
    const std::string &s;
    size_t i = 0;
    explicit Cursor(const std::string &str) : s(str) {}
    bool eof() const { return i >= s.size(); }
    char peek() const { return i < s.size() ? s[i] : '\0'; }
    static bool blank(char c) { return c == ' ' || c == '\t'; }
    void space0() { while (blank(peek())) ++i; }
    bool space1() { if (!blank(peek())) return false; space0(); return true; }
    bool tag_no_case(const char *t) {
        size_t k = 0;
        while (t[k]) {
            if (i + k >= s.size() || std::tolower((unsigned char)s[i + k]) != std::tolower((unsigned char)t[k])) return false;
            ++k;
        }
        i += k;
        return true;
    }
    // utils.rs:66-69
    bool ident(std::string &out) {
        size_t b = i;
        while (i < s.size() && (((unsigned char)s[i] < 128 && std::isalnum((unsigned char)s[i])) || s[i] == '_')) ++i;
        out = s.substr(b, i - b);
        return i > b;
    }
    // utils.rs:72-78: token = is_not(" \t\r\n()")
    bool value(double &out) {
        size_t b = i;
        while (i < s.size() && !blank(s[i]) && s[i] != '\r' && s[i] != '\n' && s[i] != '(' && s[i] != ')') ++i;
        if (i == b) return false;
        if (parse_value(s.substr(b, i - b), out) != 0) { i = b; return false; }
        return true;
    }
    // nom::number::complete::double (decimal prefix)
    bool nom_double(double &out) {
        size_t b = i, j = i;
        if (j < s.size() && (s[j] == '+' || s[j] == '-')) ++j;
        size_t digits = 0;
        while (j < s.size() && std::isdigit((unsigned char)s[j])) { ++j; ++digits; }
        if (j < s.size() && s[j] == '.') {
            size_t k = j + 1, frac = 0;
            while (k < s.size() && std::isdigit((unsigned char)s[k])) { ++k; ++frac; }
            if (digits + frac > 0) { j = k; digits += frac; }
        }
        if (digits == 0) return false;
        if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
            size_t k = j + 1;
            if (k < s.size() && (s[k] == '+' || s[k] == '-')) ++k;
            if (k >= s.size() || !std::isdigit((unsigned char)s[k])) return false;
            while (k < s.size() && std::isdigit((unsigned char)s[k])) ++k;
            j = k;
        }
        out = std::strtod(s.substr(b, j - b).c_str(), nullptr);
        i = j;
        return true;
    }
};

// <f64 as FromStr> on the upper-cased numeric part
bool rust_f64(const std::string &u, double &out) {
    size_t j = 0;
    if (j < u.size() && (u[j] == '+' || u[j] == '-')) ++j;
    const std::string rest = u.substr(j);
    if (rest == "INF" || rest == "INFINITY") { out = (u[0] == '-') ? -INFINITY : INFINITY; return true; }
    if (rest == "NAN") { out = NAN; return true; }
    size_t digits = 0;
    while (j < u.size() && std::isdigit((unsigned char)u[j])) { ++j; ++digits; }
    if (j < u.size() && u[j] == '.') { ++j; while (j < u.size() && std::isdigit((unsigned char)u[j])) { ++j; ++digits; } }
    if (!digits) return false;
    if (j < u.size() && u[j] == 'E') {
        ++j;
        if (j < u.size() && (u[j] == '+' || u[j] == '-')) ++j;
        if (j >= u.size() || !std::isdigit((unsigned char)u[j])) return false;
        while (j < u.size() && std::isdigit((unsigned char)u[j])) ++j;
    }
    if (j != u.size()) return false;
    out = std::strtod(u.c_str(), nullptr);
    return true;
}

bool two_node_head(Cursor &c, ParsedElement &e) {
    std::string name;
    if (!c.ident(name)) return false;
    e.identifier += name;
    if (!c.space1() || !c.ident(e.nodes[0])) return false;
    if (!c.space1() || !c.ident(e.nodes[1])) return false;
    return true;
}
bool paren_values(Cursor &c, double *v, int count) {
    c.space0();
    if (c.peek() != '(') return false;
    ++c.i;
    c.space0();
    for (int k = 0; k < count; k++) {
        if (k && !c.space1()) return false;
        if (!c.value(v[k])) return false;
    }
    c.space0();
    if (c.peek() != ')') return false;
    ++c.i;
    return true;
}

struct ModelCard {   // type: 0 D, 1 NMOS, 2 PMOS
    int type; double is, n;
    double w = 1e-6, l = 1e-6, vto = 0.0, kp = 2e-5, lambda = 0.02;   // models/nmosfet.rs:24-36 defaults
};

std::string trim(const std::string &s) {
    size_t b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) ++b;
    while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}
bool starts_with_nc(const std::string &s, const char *t) { Cursor c(s); return c.tag_no_case(t); }

}  // namespace

int parse_value(const std::string &token, double &out) {
    std::string u(token);
    for (auto &ch : u) ch = (char)std::toupper((unsigned char)ch);
    if (u.empty()) return 1;
    double mult = 1.0;
    if (u.size() >= 3 && u.compare(u.size() - 3, 3, "MEG") == 0) { u.resize(u.size() - 3); mult = 1e6; }
    else {
        switch (u.back()) {
        case 'F': mult = 1e-15; u.pop_back(); break;
        case 'P': mult = 1e-12; u.pop_back(); break;
        case 'N': mult = 1e-9; u.pop_back(); break;
        case 'U': mult = 1e-6; u.pop_back(); break;
        case 'M': mult = 1e-3; u.pop_back(); break;
        case 'K': mult = 1e3; u.pop_back(); break;
        case 'G': mult = 1e9; u.pop_back(); break;
        case 'T': mult = 1e12; u.pop_back(); break;
        default: break;
        }
    }
    double base;
    if (!rust_f64(u, base)) return 1;
    out = base * mult;
    return 0;
}


namespace {
// x<name> n1 n2 ... <definition>   (subcircuit.rs:147-172)
struct SubcktInstance { std::string instance_name, definition_name; std::vector<std::string> nodes; };
struct DefElement { bool is_instance = false; ParsedElement e; SubcktInstance inst; };
struct SubcktDefinition { std::vector<std::string> pins; std::vector<DefElement> elements; };

// One element line.  Returns PARSE_OK / PARSE_ERROR / PARSE_UNSUPPORTED.
int parse_element_line(const std::string &line, DefElement &de) {
    ParsedElement &e = de.e;
    // element line: the reference tries R, C, L, V, I, D, BJT, NMOS, subckt in that order (elements.rs:45-55);
    // each alternative is keyed by its leading tag, so dispatching on it is equivalent
    e.pod.ip = e.pod.im = e.pod.ic = -1;
    e.pod.n = 1.0; e.pod.is = 1e-12;
    Cursor c(line);
    const char c0 = (char)std::toupper((unsigned char)line[0]);
    bool ok = false;
    auto rcli = [&](int kind, char letter) {
        ++c.i;
        e.identifier = std::string(1, letter);
        e.pod.kind = kind;
        return two_node_head(c, e) && c.space1() && c.value(e.pod.value);
    };
    if (c0 == 'R') ok = rcli(KRB_R, 'R');
    else if (c0 == 'C') {
        ok = rcli(KRB_C, 'C');
        if (ok) { Cursor t = c; if (t.space1() && t.tag_no_case("G2")) e.pod.flags |= KRB_ELEM_FLAG_G2; }   // capacitor.rs:31
    } else if (c0 == 'L') ok = rcli(KRB_L, 'L');
    else if (c0 == 'I') ok = rcli(KRB_I, 'I');
    else if (c0 == 'V') {                                                        // voltage_source.rs:188-235
        ++c.i;
        e.identifier = "V";
        e.pod.kind = KRB_V;
        ok = two_node_head(c, e);
        if (ok) {
            double dc = 0.0, v;
            { Cursor t = c; if (t.space1() && t.nom_double(v)) { dc = v; c.i = t.i; } }
            for (;;) {
                Cursor t = c;
                if (!t.space1()) break;
                Cursor k = t;
                if (k.tag_no_case("dc") && k.space1() && k.nom_double(v)) { dc = v; c.i = k.i; continue; }
                k.i = t.i;
                if (k.tag_no_case("ac") && k.space1() && k.nom_double(v)) { e.pod.ac_amplitude = v; c.i = k.i; continue; }
                k.i = t.i;
                double w[7] = {0, 0, 0, 0, 0, 0, 0};
                if (k.tag_no_case("pulse") && paren_values(k, w, 7)) { e.pod.waveform = KRB_WAVE_PULSE; std::copy(w, w + 7, e.pod.wave); c.i = k.i; continue; }
                k.i = t.i;
                if (k.tag_no_case("sin") && paren_values(k, w, 6)) {
                    // a later SIN does not displace an earlier PULSE: transient_value_at prefers pulse (voltage_source.rs:245-249)
                    if (e.pod.waveform != KRB_WAVE_PULSE) { e.pod.waveform = KRB_WAVE_SIN; std::copy(w, w + 6, e.pod.wave); }
                    c.i = k.i;
                    continue;
                }
                break;
            }
            e.pod.value = dc;
        }
    } else if (c0 == 'D') {                                                      // diode.rs:74-92
        ++c.i;
        e.identifier = "D";
        e.pod.kind = KRB_D;
        ok = two_node_head(c, e);
        e.model_name = "default";
        if (ok) { Cursor t = c; std::string m; if (t.space1() && t.ident(m)) e.model_name = m; }
    } else if (c0 == 'M' && line.size() > 1 && std::toupper((unsigned char)line[1]) == 'N') {   // nmosfet.rs:113-166
        c.i = 2;
        e.identifier = "M";
        e.pod.kind = KRB_NMOS;
        std::string name, bulk;
        ok = c.ident(name) && c.space1() && c.ident(e.nodes[0]) && c.space1() && c.ident(e.nodes[1]) && c.space1() && c.ident(e.nodes[2]) &&
             c.space1() && c.ident(bulk) && c.space1() && c.ident(e.model_name);
        e.identifier += name;
    } else if (c0 == 'X') {                                                      // subcircuit.rs:147-172
        ++c.i;
        SubcktInstance &si = de.inst;
        if (!c.ident(si.instance_name)) return PARSE_ERROR;
        std::vector<std::string> toks;
        for (;;) { Cursor t = c; std::string tok; if (!t.space1() || !t.ident(tok)) break; toks.push_back(tok); c.i = t.i; }
        if (toks.empty()) return PARSE_ERROR;
        si.definition_name = toks.back();
        toks.pop_back();
        si.nodes = toks;
        de.is_instance = true;
        return PARSE_OK;
    } else if (c0 == 'Q') {
        return PARSE_UNSUPPORTED;                                                  // BJT: todo!() in the reference's stamps
    }
    return ok ? PARSE_OK : PARSE_ERROR;
}

// subcircuit.rs:85-116: map the nodes through the pin list (internal nodes get "<instance>_" in front — including a "0"
// that is not a pin, as in the reference) and prefix the element name
void map_sub_element(DefElement &de, const std::unordered_map<std::string, std::string> &port_to_node, const std::string &parent) {
    auto map_node = [&](std::string &node) {
        auto it = port_to_node.find(node);
        node = it != port_to_node.end() ? it->second : parent + "_" + node;
    };
    if (de.is_instance) {
        for (auto &nd : de.inst.nodes) map_node(nd);
        de.inst.instance_name = parent + "_" + de.inst.instance_name;
    } else {
        const int nn = de.e.pod.kind == KRB_NMOS ? 3 : 2;
        for (int k = 0; k < nn; k++) map_node(de.e.nodes[k]);
        de.e.identifier = de.e.identifier.substr(0, 1) + parent + "_" + de.e.identifier.substr(1);   // identifier = letter + name
    }
}

// subcircuit.rs:33-82
int instantiate(const SubcktInstance &inst, const std::unordered_map<std::string, SubcktDefinition> &defs, std::vector<ParsedElement> &out,
                std::string &err, int depth) {
    if (depth > 64) { err = "subcircuit nesting deeper than 64 levels (recursive definition?)"; return PARSE_ERROR; }
    auto it = defs.find(inst.definition_name);
    if (it == defs.end()) { err = "Undefined subcircuit definition: " + inst.definition_name; return PARSE_ERROR; }
    const SubcktDefinition &def = it->second;
    if (inst.nodes.size() != def.pins.size()) {
        err = "Node/port mismatch for instance " + inst.instance_name + ": expected " + std::to_string(def.pins.size()) + ", got " + std::to_string(inst.nodes.size());
        return PARSE_ERROR;
    }
    std::unordered_map<std::string, std::string> port_to_node;
    for (size_t k = 0; k < def.pins.size(); k++) port_to_node[def.pins[k]] = inst.nodes[k];   // later duplicates win, as HashMap::collect
    for (const DefElement &sub : def.elements) {
        DefElement mapped = sub;
        map_sub_element(mapped, port_to_node, inst.instance_name);
        if (mapped.is_instance) {
            const int rc = instantiate(mapped.inst, defs, out, err, depth + 1);
            if (rc) return rc;
        } else out.push_back(mapped.e);
    }
    return PARSE_OK;
}
}  // namespace

int parse_circuit_description(const std::string &text, ParsedCircuit &out, std::string &err) {
    out = ParsedCircuit();
    std::unordered_map<std::string, ModelCard> models;
    bool in_control = false, in_subckt = false;
    size_t pos = 0;
    int lineno = 0;
    // ---- pre-pass: subcircuit definitions (subcircuit.rs:174-218) ----
    std::unordered_map<std::string, SubcktDefinition> subckts;
    {
        std::string current;
        bool inside = false;
        size_t q = 0;
        while (q < text.size()) {
            size_t eol = text.find('\n', q);
            const std::string line = trim(text.substr(q, eol == std::string::npos ? std::string::npos : eol - q));
            q = eol == std::string::npos ? text.size() : eol + 1;
            if (line.empty() || line[0] == '%' || line[0] == '*') continue;
            if (starts_with_nc(line, ".subckt")) {
                Cursor c(line);
                c.i = 7;
                std::string name;
                if (!c.space1() || !c.ident(name)) { err = "Failed to parse subcircuit header: " + line; return PARSE_ERROR; }
                SubcktDefinition def;
                for (;;) { Cursor t = c; std::string pin; if (!t.space1() || !t.ident(pin)) break; def.pins.push_back(pin); c.i = t.i; }
                subckts[name] = def;
                current = name;
                inside = true;
                continue;
            }
            if (starts_with_nc(line, ".ends")) { inside = false; current.clear(); continue; }
            if (inside) {
                DefElement de;
                const int prc = parse_element_line(line, de);
                if (prc) { err = "Failed to parse subcircuit element in '" + current + "': " + line; return prc; }
                subckts[current].elements.push_back(de);
            }
        }
    }
    auto fail = [&](int code, const std::string &line) {
        err = "Parse error on line " + std::to_string(lineno) + ": " + line;
        return code;
    };
    while (pos <= text.size()) {
        size_t eol = text.find('\n', pos);
        std::string raw = text.substr(pos, eol == std::string::npos ? std::string::npos : eol - pos);
        if (eol == std::string::npos) { if (pos >= text.size() && raw.empty()) break; pos = text.size() + 1; } else pos = eol + 1;
        ++lineno;
        const std::string line = trim(raw);
        if (line.empty() || line[0] == '%' || line[0] == '*') continue;             // parser.rs:43-45
        if (starts_with_nc(line, ".control")) { in_control = true; continue; }
        if (starts_with_nc(line, ".endc")) { in_control = false; continue; }
        if (in_control) continue;
        if (starts_with_nc(line, ".subckt")) { in_subckt = true; continue; }
        if (starts_with_nc(line, ".ends")) { in_subckt = false; continue; }
        if (in_subckt) continue;
        if (line.size() == 4 && starts_with_nc(line, ".end")) continue;
        if (starts_with_nc(line, ".model")) {                                       // models.rs:52-95
            std::string body = line.substr(0, line.find('%'));
            Cursor c(body);
            c.i = 6;
            std::string name;
            if (!c.space1() || !c.ident(name) || !c.space1()) return fail(PARSE_ERROR, line);
            ModelCard m;                                                            // models/diode.rs:14-23
            m.type = 0; m.is = 1e-12; m.n = 1.0;
            if (body.compare(c.i, 4, "NMOS") == 0) { m.type = 1; c.i += 4; }
            else if (body.compare(c.i, 4, "PMOS") == 0) { m.type = 2; c.i += 4; }
            else if (c.peek() == 'D') { m.type = 0; c.i += 1; }
            else return fail(PARSE_ERROR, line);
            c.space0();
            if (c.peek() == '(') ++c.i;
            for (;;) {
                Cursor t = c;
                t.space0();
                std::string key;
                double v;
                if (!t.ident(key)) break;
                t.space0();
                if (t.peek() != '=') break;
                ++t.i;
                t.space0();
                if (!t.value(v)) break;
                for (auto &ch : key) ch = (char)std::tolower((unsigned char)ch);
                if (m.type == 0) { if (key == "is") m.is = v; else if (key == "n") m.n = v; }   // models/diode.rs:26-39
                else if (m.type == 1) {                                                         // models/nmosfet.rs:48-61
                    if (key == "w") m.w = v; else if (key == "l") m.l = v; else if (key == "vto") m.vto = v;
                    else if (key == "kp") m.kp = v; else if (key == "lambda") m.lambda = v;
                }
                c.i = t.i;
            }
            models[name] = m;
            continue;
        }
        DefElement de;
        const int prc = parse_element_line(line, de);
        if (prc) return fail(prc, line);
        if (de.is_instance) {                                                          // parser.rs:93-98
            std::string ierr;
            const int irc = instantiate(de.inst, subckts, out.elements, ierr, 0);
            if (irc) { err = ierr; return irc; }
        } else out.elements.push_back(de.e);
    }
    // ---- index map: parser.rs:105-125 ----
    std::unordered_set<std::string> seen_nodes;
    for (auto &e : out.elements) {
        const int kind = e.pod.kind;
        const bool g2 = kind == KRB_V || kind == KRB_L || kind == KRB_I || (kind == KRB_C && (e.pod.flags & KRB_ELEM_FLAG_G2));   // elements.rs:125-140
        if (g2) {
            const std::string key = "I(" + e.identifier + ")";
            if (out.index_map.count(key)) { err = "duplicate element identifier " + e.identifier; return PARSE_ERROR; }
            out.index_map[key] = (int)out.unknown_names.size();
            out.unknown_names.push_back(key);
        }
        const int nn = (kind == KRB_NMOS) ? 3 : 2;
        for (int k = 0; k < nn; k++) {
            if (!seen_nodes.insert(e.nodes[k]).second) continue;
            if (e.nodes[k] == "0") continue;
            const std::string key = "V(" + e.nodes[k] + ")";
            out.index_map[key] = (int)out.unknown_names.size();
            out.unknown_names.push_back(key);
        }
    }
    if (out.elements.empty()) { err = "The netlist is empty"; return PARSE_EMPTY_NETLIST; }   // parser.rs:127-129
    // ---- models: parser.rs:132-149 ----
    for (auto &e : out.elements) {
        if (e.pod.kind == KRB_D || e.pod.kind == KRB_NMOS) {
            auto it = models.find(e.model_name);
            const int want = e.pod.kind == KRB_D ? 0 : 1;
            if (it == models.end() || it->second.type != want) { err = "model '" + e.model_name + "' is not defined"; return PARSE_MISSING_MODEL; }
            e.pod.is = it->second.is;
            e.pod.n = it->second.n;
            if (e.pod.kind == KRB_NMOS) {   // level-1 parameters (nmosfet.rs:39-49): wave[0] = Vth, wave[1] = beta = KP * (W / L), wave[2] = lambda
                e.pod.wave[0] = it->second.vto;
                e.pod.wave[1] = it->second.kp * (it->second.w / it->second.l);   // models/nmosfet.rs:43-45
                e.pod.wave[2] = it->second.lambda;
            }
        }
    }
    auto lookup = [&](const std::string &key) { auto it = out.index_map.find(key); return it == out.index_map.end() ? -1 : it->second; };
    for (auto &e : out.elements) {
        e.pod.ip = lookup("V(" + e.nodes[0] + ")");
        e.pod.im = lookup("V(" + e.nodes[1] + ")");
        e.pod.ic = lookup("I(" + e.identifier + ")");
        if (e.pod.kind == KRB_NMOS) e.pod.ic = lookup("V(" + e.nodes[2] + ")");   // NMOS: the source node (group 1: no branch current)
        out.pod.push_back(e.pod);
    }
    out.is_current.resize(out.unknown_names.size());
    for (size_t i = 0; i < out.unknown_names.size(); i++) out.is_current[i] = out.unknown_names[i][0] == 'I';   // solver.rs:96
    return PARSE_OK;
}

std::vector<double> generate_frequencies(int kind, uint32_t points, double fstart, double fstop) {
    std::vector<double> f;
    if (fstart <= 0.0 || fstop <= 0.0 || fstart > fstop) return f;                 // analyses.rs:85-90
    if (kind == 0) {                                                               // :93-102
        if (points == 1) f.push_back(fstart);
        else if (points > 1) {
            const double step = (fstop - fstart) / (double)(points - 1);
            for (uint32_t i = 0; i < points; i++) f.push_back(fstart + (double)i * step);
        }
    } else {                                                                       // :104-154
        if (points == 0) return f;
        const double span = (kind == 1) ? std::log10(fstop / fstart) : std::log2(fstop / fstart);
        const double r = std::round(span * (double)points);
        uint64_t total = r <= 0.0 ? 0u : (r >= 4294967295.0 ? 4294967295ull : (uint64_t)r);   // `as u32` saturates
        total += 1;
        const double factor = std::pow(kind == 1 ? 10.0 : 2.0, 1.0 / (double)points);
        double cur = fstart;
        for (uint64_t i = 0; i < total; i++) {
            if (cur > fstop * (1.0 + 1e-9)) break;
            f.push_back(cur);
            cur *= factor;
            if (cur > fstop && (f.empty() || f.back() < fstop)) { f.push_back(fstop); break; }
        }
        if (f.empty() || f.back() < fstop * (1.0 - 1e-9)) f.push_back(fstop);
    }
    std::sort(f.begin(), f.end());                                                 // :155-156
    f.erase(std::unique(f.begin(), f.end()), f.end());
    return f;
}

}  // namespace krh
