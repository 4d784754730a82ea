/*
 * circuit.cpp — builders for the circuits of Cloud/cloud.c and the leveliser.
 *   add    cloud.c:18-51      split  cloud.c:65-113     mulK  cloud.c:115-647
 *   dispatch (ADD / SUB / MUL per width)  cloud.c:870-1190, 1194-2365, 2368-2719
 */
#include "circuit.h"

#include <algorithm>
#include <cstring>

#include "../../include/ieache_b200.h"

namespace ieache {

Circuit::Word Circuit::input_word(int first) const
{
    Word w(32);
    for (int i = 0; i < 32; i++) w[i] = input(first + i);
    return w;
}
Circuit::Word Circuit::const_word(uint32_t v)
{
    Word w(32);
    for (int i = 0; i < 32; i++) w[i] = const_bit((v >> i) & 1);
    return w;
}
Circuit::Word Circuit::NOTw(const Word &x)
{
    Word w(x.size());
    for (size_t i = 0; i < x.size(); i++) w[i] = NOT(x[i]);
    return w;
}
Ref Circuit::gate(int op, Ref a, Ref b)
{
    CGate g;
    g.op = (uint8_t)op; g.a = a; g.b = b; g.level = 0;
    gates.push_back(g);
    return Ref{kWireFirstInput + n_inputs + (int32_t)gates.size() - 1, false};
}

/* cloud.c:25-44: axc = x^c; bxc = y^c; sum = x^bxc; axc = axc&bxc; c = c^axc */
Circuit::Word Circuit::add(const Word &x, const Word &y, Ref carry_in, Ref *carry_out)
{
    Word sum(32);
    Ref carry = carry_in;
    for (int i = 0; i < 32; i++) {
        Ref axc = gate(IEACHE_OP_XOR, x[i], carry);
        Ref bxc = gate(IEACHE_OP_XOR, y[i], carry);
        sum[i] = gate(IEACHE_OP_XOR, x[i], bxc);
        axc = gate(IEACHE_OP_AND, axc, bxc);
        carry = gate(IEACHE_OP_XOR, carry, axc);
    }
    if (carry_out) *carry_out = carry;
    return sum;
}

/* cloud.c:148-198 / 265-359 / 443-612 */
std::vector<Circuit::Word> Circuit::mulK(const std::vector<Word> &x, const Word &m, Ref carry_in)
{
    const int K = (int)x.size();
    std::vector<Word> sums(K + 1, const_word(0));
    std::vector<Word> word(K + 1, const_word(0));
    std::vector<Word> tmp(K, Word(32));
    for (int round = 0; round < 32; round++) {
        for (int kk = 0; kk < 32; kk++)
            for (int q = 0; q < K; q++) tmp[q][kk] = gate(IEACHE_OP_AND, x[q][kk], m[round]);
        const int c1 = 32 - round, c2 = round;
        for (int i = 0; i < round; i++) word[0][i] = const_bit(0);
        for (int i = 0; i < c1; i++) word[0][i + round] = tmp[0][i];
        for (int q = 1; q <= K; q++) {
            for (int i = 0; i < c2; i++) word[q][i] = tmp[q - 1][i + c1];
            if (q < K) for (int i = 0; i < c1; i++) word[q][i + c2] = tmp[q][i];
        }
        Ref cy = carry_in;
        for (int q = 0; q <= K; q++) sums[q] = add(sums[q], word[q], q ? cy : carry_in, &cy);
    }
    return sums;
}

void Circuit::finalize()
{
    const int first_gate = kWireFirstInput + n_inputs;
    const int n_wires = first_gate + (int)gates.size();
    /* ASAP levels */
    std::vector<int32_t> wl(n_wires, 0);
    int32_t depth = 0;
    n_and = n_xor = 0;
    for (size_t i = 0; i < gates.size(); i++) {
        CGate &g = gates[i];
        g.level = 1 + std::max(wl[g.a.wire], wl[g.b.wire]);
        wl[first_gate + i] = g.level;
        depth = std::max(depth, g.level);
        if (g.op == IEACHE_OP_AND) n_and++; else if (g.op == IEACHE_OP_XOR) n_xor++;
    }
    /* last level in which each wire is read; outputs live to the end */
    std::vector<int32_t> last(n_wires, 0);
    for (const CGate &g : gates) {
        last[g.a.wire] = std::max(last[g.a.wire], g.level);
        last[g.b.wire] = std::max(last[g.b.wire], g.level);
    }
    for (const Ref &r : outputs) last[r.wire] = depth + 1;

    std::vector<std::vector<int32_t>> by_level(depth + 1);
    for (size_t i = 0; i < gates.size(); i++) by_level[gates[i].level].push_back((int32_t)i);

    slot_of_wire.assign(n_wires, -1);
    for (int w = 0; w < first_gate; w++) slot_of_wire[w] = w;
    int next_slot = first_gate;
    std::vector<int32_t> free_slots;
    std::vector<std::vector<int32_t>> expire(depth + 2); /* slots to free after level L */
    static const int8_t lin[10][3] = {{+1, -1, -1}, {+1, +1, +1}, {-1, +1, +1}, {+2, +2, +2}, {-2, -2, -2},
                                      {-1, -1, -1}, {-1, -1, +1}, {-1, +1, -1}, {+1, -1, +1}, {+1, +1, -1}};
    levels.assign(depth, Level());
    max_width = 0;
    for (int L = 1; L <= depth; L++) {
        Level &lv = levels[L - 1];
        lv.tmpl.reserve(by_level[L].size());
        for (int32_t gi : by_level[L]) {
            const CGate &g = gates[gi];
            int32_t slot;
            if (!free_slots.empty()) { slot = free_slots.back(); free_slots.pop_back(); }
            else slot = next_slot++;
            const int w = first_gate + gi;
            slot_of_wire[w] = slot;
            /* a wire nobody reads still needs somewhere to be written; free it after its own level */
            const int32_t lu = std::max(last[w], (int32_t)L);
            if (lu <= depth) expire[lu].push_back(slot);
            GateT t;
            t.in0 = slot_of_wire[g.a.wire];
            t.in1 = slot_of_wire[g.b.wire];
            t.out = slot;
            t.c0 = (int8_t)(g.a.neg ? -lin[g.op][1] : lin[g.op][1]);
            t.c1 = (int8_t)(g.b.neg ? -lin[g.op][2] : lin[g.op][2]);
            t.cst_mu = lin[g.op][0];
            lv.tmpl.push_back(t);
        }
        max_width = std::max<uint32_t>(max_width, (uint32_t)lv.tmpl.size());
        for (int32_t s : expire[L]) free_slots.push_back(s);
    }
    n_slots = next_slot;
}

static int add_chain(Circuit &c, std::vector<Circuit::Word> &res, const std::vector<Circuit::Word> &x,
                     const std::vector<Circuit::Word> &y, Ref carry_in)
{
    Ref cy = carry_in;
    for (size_t j = 0; j < x.size(); j++) res.push_back(c.add(x[j], y[j], j ? cy : carry_in, &cy));
    return 0;
}

bool build_circuit(int kind, int width, Circuit &c)
{
    using Word = Circuit::Word;
    c = Circuit();
    c.kind = kind; c.width = width;
    const int nc = width / 32;
    if (width % 32) return false;
    auto words = [&](int first_chunk, int count) {
        std::vector<Word> v;
        for (int j = 0; j < count; j++) v.push_back(c.input_word((first_chunk + j) * 32));
        return v;
    };
    std::vector<Word> res;
    if (kind == IEACHE_CIRC_ADD || kind == IEACHE_CIRC_SUB) {
        if (!(nc == 1 || nc == 2 || nc == 4 || nc == 8)) return false;
        c.n_inputs = (2 * nc + 1) * 32;
        const std::vector<Word> A = words(0, nc), B = words(nc, nc);
        const Ref carry1 = c.input(2 * nc * 32);
        if (kind == IEACHE_CIRC_ADD) {
            add_chain(c, res, A, B, carry1);                         /* cloud.c:891,951-952,1020-1023,1109-1116 */
        } else {
            /* cloud.c:1225-1245 (and wider variants): inverse = NOT(B); twos = inverse + 1; A + twos */
            std::vector<Word> tw;
            Ref cy = Circuit::const_bit(0);
            for (int j = 0; j < nc; j++)
                tw.push_back(c.add(Circuit::NOTw(B[j]), j ? Circuit::const_word(0) : Circuit::const_word(1),
                                   j ? cy : Circuit::const_bit(0), &cy));
            add_chain(c, res, A, tw, carry1);
        }
    } else if (kind == IEACHE_CIRC_MUL) {
        if (!(nc == 1 || nc == 2 || nc == 4)) return false;
        c.n_inputs = (2 * nc + 1) * 32;
        const std::vector<Word> A = words(0, nc), B = words(nc, nc);
        const Word carryw = c.input_word(2 * nc * 32);
        const Ref carry1 = carryw[0];
        if (nc == 1) {                                               /* cloud.c:2668,2683-2686 */
            res = c.mulK({A[0]}, B[0], carry1);                      /* low, high */
        } else if (nc == 2) {                                        /* cloud.c:2589-2616 */
            std::vector<Word> p0 = c.mulK({A[0], A[1]}, B[0], carry1); /* r3,r2,r1 = p0[0],p0[1],p0[2] */
            std::vector<Word> p1 = c.mulK({A[0], A[1]}, B[1], carry1); /* r6,r5,r4 */
            /* split(a=r1,b=r2,c=r4,d=r5,e=r6): sum=e+b, sum2=d+a+co, sum3=c+co2(as number)+carry */
            Ref co, co2;
            Word sum = c.add(p1[0], p0[1], carry1, &co);
            Word sum2 = c.add(p1[1], p0[2], co, &co2);
            Word co2w = Circuit::const_word(0);
            co2w[0] = co2;
            Word sum3 = c.add(p1[2], co2w, carry1, nullptr);
            res = {p0[0], sum, sum2, sum3};
        } else {                                                     /* cloud.c:2434-2491 */
            std::vector<Word> r(21);
            for (int y = 0; y < 4; y++) {
                std::vector<Word> p = c.mulK({A[0], A[1], A[2], A[3]}, B[y], carry1);
                for (int q = 0; q < 5; q++) r[5 * y + 1 + q] = p[4 - q];     /* result1 = top ... result5 = lowest */
            }
            std::vector<Word> s(16);
            std::vector<Ref> co(16);
            s[1] = c.add(r[10], r[4], carry1, &co[1]);
            s[2] = c.add(r[9], r[3], co[1], &co[2]);
            s[3] = c.add(r[8], r[2], co[2], &co[3]);
            s[4] = c.add(r[7], r[1], co[3], &co[4]);
            s[5] = c.add(r[6], carryw, co[4], &co[5]);
            s[6] = c.add(s[2], r[15], co[5], &co[6]);
            s[7] = c.add(s[3], r[14], co[6], &co[7]);
            s[8] = c.add(s[4], r[13], co[7], &co[8]);
            s[9] = c.add(s[5], r[12], co[8], &co[9]);
            s[10] = c.add(r[11], carryw, co[9], &co[10]);
            s[11] = c.add(s[7], r[20], co[10], &co[11]);
            s[12] = c.add(s[8], r[19], co[11], &co[12]);
            s[13] = c.add(s[9], r[18], co[12], &co[13]);
            s[14] = c.add(s[10], r[17], co[13], &co[14]);
            s[15] = c.add(r[16], carryw, co[14], &co[15]);
            res = {r[5], s[1], s[6], s[11], s[12], s[13], s[14], s[15]};
        }
    } else if (kind == IEACHE_CIRC_MULADD) {
        /* postfix AB*C+ on the Cloud node: op 4 at width 32, then op 1 at width 64 on
         * [answer || C] (Cloud/dragonfly_cipher_cloud.py:1306-1315; cloud.c:937-952) */
        if (nc != 1) return false;
        c.n_inputs = 5 * 32; /* A, B, C chunk 1, C chunk 2 (encrypted zeros), carry block of A */
        const Word A = c.input_word(0), B = c.input_word(32), C1 = c.input_word(64), C2 = c.input_word(96);
        const Ref carry1 = c.input(128);
        std::vector<Word> prod = c.mulK({A}, B, carry1);
        add_chain(c, res, {prod[0], prod[1]}, {C1, C2}, carry1);
    } else {
        return false;
    }
    for (const Word &w : res) for (const Ref &r : w) c.outputs.push_back(r);
    c.finalize();
    return true;
}

} // namespace ieache
