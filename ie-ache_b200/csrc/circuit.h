/*
 * circuit.h — the encrypted arithmetic circuits of Cloud/cloud.c as data: an SSA gate DAG,
 * ASAP-levelised, with wire-slot reuse so that every independent gate of one circuit depth,
 * across all bits and all queued expressions, runs in one blind-rotation launch.
 *
 * The reference evaluates these circuits one libtfhe call at a time (Cloud/cloud.c:18-647);
 * bootsNOT / bootsCOPY / bootsCONSTANT are free linear operations there and become wiring here
 * (a negated or constant reference), so only bootsXOR / bootsAND produce gates.  No gate the
 * reference bootstraps is folded away: counts match SURVEY.md App. B.
 */
#ifndef IEACHE_CIRCUIT_H
#define IEACHE_CIRCUIT_H
#include <cstdint>
#include <vector>

#include "kernels.h"

namespace ieache {

/* reference to a wire; wire 0 / 1 are the trivial samples of bootsCONSTANT(0) / (1) */
struct Ref {
    int32_t wire;
    bool neg;
};
constexpr int32_t kWireConst0 = 0, kWireConst1 = 1, kWireFirstInput = 2;

struct CGate {
    uint8_t op;      /* IEACHE_OP_AND / IEACHE_OP_XOR / ... (binary bootstrapped ops only) */
    Ref a, b;
    int32_t level;   /* ASAP depth, 1-based */
};

struct Level {
    std::vector<GateT> tmpl; /* per-instance templates: slot indices + coefficients */
};

struct Circuit {
    int kind = 0, width = 0;
    int n_inputs = 0;              /* input samples per instance (excludes the two constants) */
    std::vector<CGate> gates;      /* wire id of gate i = kWireFirstInput + n_inputs + i */
    std::vector<Ref> outputs;      /* output samples per instance, in answer.data order */

    /* filled by finalize() */
    std::vector<Level> levels;
    std::vector<int32_t> slot_of_wire;
    int n_slots = 0;               /* samples per instance block, including constants and inputs */
    uint64_t n_and = 0, n_xor = 0;
    uint32_t max_width = 0;

    /* --- builder API, mirroring Cloud/cloud.c --- */
    using Word = std::vector<Ref>; /* 32 refs, LSB first */
    Ref input(int idx) const { return Ref{kWireFirstInput + idx, false}; }
    Word input_word(int first) const;
    static Ref const_bit(int v) { return Ref{v ? kWireConst1 : kWireConst0, false}; }
    static Word const_word(uint32_t v);
    static Ref NOT(Ref x) { return Ref{x.wire, !x.neg}; }
    static Word NOTw(const Word &x);
    Ref gate(int op, Ref a, Ref b);
    /* add (cloud.c:18-51): returns sum, sets carry_out */
    Word add(const Word &x, const Word &y, Ref carry_in, Ref *carry_out);
    /* mul32/mul64/mul128 (cloud.c:115-647): sums[0..K], least significant first */
    std::vector<Word> mulK(const std::vector<Word> &x, const Word &m, Ref carry_in);

    void finalize();
};

/* Cloud/cloud.c main() dispatch for one (kind, width); see ieache_circuit_kind */
bool build_circuit(int kind, int width, Circuit &c);

} // namespace ieache
#endif
