// Host-side front end (the role krets-parser plays for the solver): netlist text -> element table + MNA index map.
// Only what the solver hot path needs: R/C/L/V/I/D/MN element lines, .model D/NMOS,
// comment / .control skipping, .subckt definitions and instance flattening (subcircuit.rs), the index assignment of parser.rs:105-125, and the AC frequency grids of
// analyses.rs:80-158.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/krets_b200.h"

namespace krh {

enum ParseStatus {
    PARSE_OK = 0,
    PARSE_ERROR = 30,          // krets_parser::Error::ParseError / InvalidFormat
    PARSE_EMPTY_NETLIST = 31,  // krets_parser::Error::EmptyNetlist
    PARSE_UNSUPPORTED = 32,    // BJT: outside the restated subset (the reference's stamps are todo!())
    PARSE_MISSING_MODEL = 33,  // reference: todo!() at parser.rs:138,146
    PARSE_IO = 34
};

struct ParsedElement {
    std::string identifier;    // "R1", "V1", ... (resistor.rs:21-23 and siblings)
    std::string nodes[3];      // plus, minus (NMOS: drain, gate, source)
    std::string model_name;
    krb_element pod{};
};

struct ParsedCircuit {
    std::vector<ParsedElement> elements;
    std::vector<std::string> unknown_names;                 // index -> key ("V(in)", "I(V1)")
    std::unordered_map<std::string, int> index_map;         // circuit.rs:12
    std::vector<krb_element> pod;                           // flattened table handed to krb_circuit_create
    std::vector<uint8_t> is_current;
};

int parse_circuit_description(const std::string &text, ParsedCircuit &out, std::string &err);
int parse_value(const std::string &token, double &out);   // utils.rs:32-64
// analyses.rs:80-158; kind 0 Linear{total_points}, 1 Decade{points_per_decade}, 2 Octave{points_per_octave}
std::vector<double> generate_frequencies(int kind, uint32_t points, double fstart, double fstop);

}  // namespace krh
