// cb_specialize.h -- host-side generator of table-specialised block evaluators.
//
// The lean kernel body interprets a policy block's records (conditions = lists of 16-byte DNF terms, rows = 16-byte
// records) for every request.  Most of that work is decoding data that is fixed once the table is loaded.  When a
// table is loaded the library therefore emits, for every distinct block SHAPE (same rows + same conditions), a
// function that calls the very same force-inlined device helpers of cb_core.h -- term_tri(), term_lit(),
// row_apply() -- with every record as a compile-time constant, and a `SpecBlocks` dispatcher (switch on the block
// id) that replaces cb::GenericBlocks in the kernels of cb_kernels.h.  NVRTC then folds the switches, operand kinds
// and slot indices away: what is left per term is the operand loads and the compare.  Semantics are identical by
// construction (same helpers, same order), which tests/ check through a host build of the generated text.
//
// Host-only, no CUDA dependencies: tests/hostsim uses it too.
#pragma once
#include <stdint.h>

#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "cerbos_b200_format.h"

namespace cbspec {

struct Limits {
    uint32_t max_shapes = 8;        // more distinct shapes than this: keep the generic walker (code size, NVRTC time:
    uint32_t max_items = 96;        // terms + rows over all shapes     ~1 s for 10 items, ~17 s for 160 on one host core)
};

inline std::string hex(uint32_t v) {
    char buf[16];
    snprintf(buf, sizeof buf, "0x%xu", v);
    return buf;
}
inline std::string hex64(uint64_t v) {
    char buf[32];
    snprintf(buf, sizeof buf, "0x%llxull", (unsigned long long)v);
    return buf;
}

// One DNF term as source text: an expression of type int (TRI_T / TRI_F / TRI_E) that may raise `slow`.
// Shapes with constant operands get the constants as immediates (no table load, no list walk); everything else
// goes through term_tri() with the term words as compile-time constants.
inline std::string term_expr(const uint32_t *w, const uint64_t *consts, const uint64_t *theap) {
    const uint32_t op = w[0] & 0xFFu, flags = (w[0] >> 8) & 0xFFu;
    const std::string sx = "cols.slot(" + std::to_string(w[1]) + "u)";
    switch (op) {
    case CB_TERM_EQ_SC: return "eq_tri(" + sx + ", " + hex64(consts[w[2]]) + ", slow)";
    case CB_TERM_ORD_SC: return "ord_tri(" + hex(flags & CB_TERM_CI_MASK) + ", " + sx + ", " + hex64(consts[w[2]]) + ", slow)";
    case CB_TERM_IN_SC: {
        const uint64_t lst = consts[w[2]];
        const uint64_t *p = theap + (lst & 0xFFFFFFFFFFFFull);
        const uint32_t n = (uint32_t)p[0];
        if (n > 16) break;
        std::string e = "in_const_tri(" + sx + ", slow";
        for (uint32_t j = 0; j < n; j++) e += ", " + hex64(p[1 + j]);
        return e + ")";
    }
    default: break;
    }
    return "term_tri(t, b, cols, pid, U4{" + hex(w[0]) + ", " + hex(w[1]) + ", " + hex(w[2]) + ", " + hex(w[3]) + "}, slow)";
}

// image: host copy of the table image (section offsets in `off`, indexed by CB_SEC_*).  Returns the generated
// source ("" = the table does not qualify: a condition without flat form, too many shapes ...).
inline std::string generate(const uint8_t *image, const uint32_t *off, const uint32_t *meta, const Limits lim = Limits()) {
    const uint32_t n_blocks = meta[CB_META_N_BLOCKS];
    if (n_blocks == 0) return "";
    const uint32_t *blocks = reinterpret_cast<const uint32_t *>(image + off[CB_SEC_BLOCKS]);      // {row_start, n_rows, cond_base, n_conds}
    const uint32_t *rows = reinterpret_cast<const uint32_t *>(image + off[CB_SEC_ROWS]);          // 4 words each
    const uint32_t *conds = reinterpret_cast<const uint32_t *>(image + off[CB_SEC_CONDS]);        // {code_off, code_len, flat_off, flat_info}
    const uint32_t *code = reinterpret_cast<const uint32_t *>(image + off[CB_SEC_CODE]);          // 8-byte instruction slots
    const uint32_t *bs_off = reinterpret_cast<const uint32_t *>(image + off[CB_SEC_BLOCK_SLOTS_OFF]);
    const uint32_t *bs = reinterpret_cast<const uint32_t *>(image + off[CB_SEC_BLOCK_SLOTS]);
    const uint64_t *consts = reinterpret_cast<const uint64_t *>(image + off[CB_SEC_CONSTS_V64]);
    const uint64_t *theap = reinterpret_cast<const uint64_t *>(image + off[CB_SEC_THEAP]);

    std::map<std::vector<uint32_t>, uint32_t> shape_ids;
    std::vector<std::vector<uint32_t>> shape_blocks;     // shape -> block ids
    uint32_t items = 0;
    for (uint32_t bid = 0; bid < n_blocks; bid++) {
        const uint32_t *bl = blocks + 4 * bid;
        if (bl[3] > 31) return "";
        std::vector<uint32_t> key;
        key.push_back(bl[1]);
        key.push_back(bl[3]);
        for (uint32_t r = 0; r < bl[1]; r++) {
            const uint32_t *row = rows + 4 * (bl[0] + r);
            key.push_back(row[0]);            // role | cond << 16
            key.push_back(row[1] & 0xFFFFu);  // drcond
            key.push_back(row[2] & 0xFFu);    // effect
        }
        for (uint32_t c = 0; c < bl[3]; c++) {
            const uint32_t *cd = conds + 4 * (bl[2] + c);
            if (cd[3] == 0 || ((cd[3] >> 16) & 0xFF) != CB_FLAT_DNF) return "";   // no flat form: generic interpreter needed
            key.push_back(cd[2]);
            key.push_back(cd[3]);
        }
        for (uint32_t q = bs_off[bid]; q < bs_off[bid + 1]; q++) key.push_back(bs[q]);
        auto it = shape_ids.find(key);
        if (it == shape_ids.end()) {
            it = shape_ids.emplace(key, (uint32_t)shape_blocks.size()).first;
            shape_blocks.emplace_back();
            items += bl[1];
            for (uint32_t c = 0; c < bl[3]; c++) items += conds[4 * (bl[2] + c) + 3] & 0xFFFFu;
        }
        shape_blocks[it->second].push_back(bid);
    }
    if (shape_blocks.size() > lim.max_shapes || items > lim.max_items) return "";

    std::string s;
    s += "// generated by cb_specialize.h from the loaded table: one straight-line evaluator per block shape\n";
    s += "namespace cb {\n";
    const char *args_decl =
        "const TableView t, const BatchView &b, const Cols &cols, uint32_t pid, uint64_t rp, uint32_t RCP, uint32_t role_all, uint32_t alive, "
        "const uint64_t *ram, uint32_t &D, uint32_t &A, bool &defer";
    for (uint32_t sh = 0; sh < shape_blocks.size(); sh++) {
        const uint32_t bid = shape_blocks[sh][0];
        const uint32_t *bl = blocks + 4 * bid;
        s += "template <typename Cols>\nCB_HD void spec_shape_" + std::to_string(sh) + "(" + args_decl + ") {\n";
        if (bs_off[bid + 1] > bs_off[bid]) {
            s += "    if (!cols.staged()) {";
            for (uint32_t q = bs_off[bid]; q < bs_off[bid + 1]; q++) s += " cols.prefetch_slot(" + std::to_string(bs[q]) + "u);";
            s += " }\n";
        }
        s += "    bool slow = false;\n    uint32_t val = 1u;\n";
        for (uint32_t c = 0; c < bl[3]; c++) {
            const uint32_t *cd = conds + 4 * (bl[2] + c);
            const uint32_t nt = cd[3] & 0xFFFFu, negate = (cd[3] >> 24) & 1u;
            s += "    {   // condition " + std::to_string(c + 1) + "\n        bool any = false, group = true;\n";
            for (uint32_t i = 0; i < nt; i++) {
                const uint32_t *w = code + 2 * (cd[2] + 2 * i);   // a term = two 8-byte instruction slots
                const uint32_t flags = (w[0] >> 8) & 0xFFu;
                s += "        group &= term_lit(" + term_expr(w, consts, theap) + ", " + hex(flags) + ");\n";
                if (flags & CB_TERM_GROUP_END) s += "        any |= group; group = true;\n";
            }
            s += std::string("        val |= (uint32_t)(any != ") + (negate ? "true" : "false") + ") << " + std::to_string(c + 1) + ";\n    }\n";
        }
        s += "    defer |= slow;\n";
        for (uint32_t r = 0; r < bl[1]; r++) {
            const uint32_t *row = rows + 4 * (bl[0] + r);
            s += "    row_apply(ldg(reinterpret_cast<const uint32_t *>(ram + " + std::to_string(r) + ")), U4{" + hex(row[0]) + ", " + hex(row[1] & 0xFFFFu) + ", " +
                 hex(row[2] & 0xFFu) + ", 0u}, rp, RCP, role_all, alive, val, D, A);\n";
        }
        s += "}\n";
    }
    s += "struct SpecBlocks {\n    template <typename Cols>\n    CB_HD void operator()(const TableView t, const BatchView &b, const Cols &cols, uint32_t pid, uint32_t bid, "
         "uint64_t rp, uint32_t RCP, uint32_t role_all, uint32_t alive, const uint64_t *row_am, uint32_t &D, uint32_t &A, bool &defer) const {\n";
    s += "        const uint64_t *ram = row_am + ldg(reinterpret_cast<const uint32_t *>(t.blocks() + bid));   // + row_start\n";
    s += "        switch (bid) {\n";
    for (uint32_t sh = 0; sh < shape_blocks.size(); sh++) {
        s += "       ";
        for (uint32_t bid : shape_blocks[sh]) s += " case " + std::to_string(bid) + ":";
        s += "\n            spec_shape_" + std::to_string(sh) + "(t, b, cols, pid, rp, RCP, role_all, alive, ram, D, A, defer);\n            break;\n";
    }
    s += "        default: defer = true; break;\n        }\n    }\n};\n}  // namespace cb\n";
    return s;
}

// ---- conditions without a flat form: their bytecode programs as straight-line code ------------------------------------
// A condition program (table/bytecode.py: compile_cond) is a condition TREE (all / any / none) over CEL leaf expressions:
//   leaf ... TO_COND [JF_KEEP | JT_KEEP] leaf ... TO_COND (AND | OR) ... [COND_NOT] RET
// A leaf only asks "is the value BOOL true" (errors are plain false: ruletable.go:1467-1486 drops CEL errors), it has no
// side effect, so the tree is a boolean formula over its leaves.  translate_program() splits a program into its leaves
// (ATOMS; the same leaf in several conditions is one atom, evaluated once per request) and the formula; atom_source()
// turns a leaf's instructions into C++ that calls the interpreter's own per-instruction helpers (cb_core.h: op_*,
// do_cmp, do_in ...) with the operand stack as named locals and jumps as gotos -- same helpers, same order, so the
// semantics are the interpreter's by construction; NVRTC then folds constant tags and keeps the stack in registers.
// Programs using instructions outside the supported set (values built in the arena by collecting comprehensions,
// list / map literals, runtime.effectiveDerivedRoles) make the table "not qualify", as before.
struct Ins { uint32_t op, ia, ib, ic; };
struct Atoms {
    std::map<std::vector<uint32_t>, uint32_t> ids;    // normalised instruction list -> atom number
    std::vector<std::string> src;                     // one function per atom
    std::vector<bool> slot_used;
    // per atom: the one slot it reads (CB_NONE32: none or several) and whether it reads anything else of the request
    // (P.id) -- an atom over one slot is a function of that slot's value alone (and of the batch's `now`)
    std::vector<uint32_t> only_slot;
    std::vector<bool> reads_pid;
};

inline bool is_jump(uint32_t op) {
    return op == CB_OP_JF_KEEP || op == CB_OP_JT_KEEP || op == CB_OP_JMP || op == CB_OP_TERN || op == CB_OP_LOOP_INIT || op == CB_OP_LOOP_NEXT || op == CB_OP_LOOP_PRED;
}

// One leaf P[lo, hi) (hi = its TO_COND) -> the body of `template <typename Cols> CB_HD bool uc_atom_K(Ctx &c, const Cols &cols)`.
// "" = an unsupported instruction or a malformed program.
inline std::string atom_source(const std::vector<Ins> &P, uint32_t lo, uint32_t hi, const uint32_t *consts /* cb_const words */, uint32_t n_consts, uint32_t n_slots,
                               std::vector<bool> &slot_used, uint32_t *only_slot = nullptr, bool *reads_pid = nullptr) {
    std::vector<uint32_t> my_slots;
    bool my_pid = false;
    const uint32_t n = hi - lo;
    const int kUnset = -1000;
    std::vector<int> depth(n + 1, kUnset), ldep(n + 1, kUnset);
    std::vector<bool> target(n + 1, false);
    bool bad = false;
    auto set = [&](uint32_t k, int d, int l) {
        if (k > n) { bad = true; return; }
        if (depth[k] == kUnset) { depth[k] = d; ldep[k] = l; }
        else if (depth[k] != d || ldep[k] != l) bad = true;
    };
    auto rel = [&](uint32_t abs) -> uint32_t { if (abs < lo || abs > hi) { bad = true; return 0; } return abs - lo; };
    set(0, 0, 0);
    for (uint32_t k = 0; k < n && !bad; k++) {
        if (depth[k] == kUnset) { bad = true; break; }   // unreachable instruction: not something compile_cond emits
        const Ins &I = P[lo + k];
        const int d = depth[k], l = ldep[k];
        int nd = d;
        bool falls = true;
        switch (I.op) {
        case CB_OP_CONST: case CB_OP_SLOT: case CB_OP_HAS_SLOT: case CB_OP_PID: case CB_OP_NOW: case CB_OP_VAR:
        case CB_OP_CMP_SLOT_CONST: case CB_OP_CMP_SLOT_SLOT: case CB_OP_CMP_SLOT_PID: case CB_OP_IN_SLOT_CONST: case CB_OP_IN_CONST_SLOT:
            nd = d + 1; break;
        case CB_OP_SELECT: case CB_OP_HAS: case CB_OP_NEG: case CB_OP_NOT: case CB_OP_SIZE: case CB_OP_NOERR: case CB_OP_INT: case CB_OP_UINT:
        case CB_OP_DOUBLE: case CB_OP_TIMESTAMP: case CB_OP_DURATION: case CB_OP_DYN: case CB_OP_IN_IP_RANGE: case CB_OP_HIER_SIZE: case CB_OP_TS_GET:
        case CB_OP_MATCHES:
            if (d < 1) bad = true;
            break;
        case CB_OP_INDEX: case CB_OP_EQ: case CB_OP_NE: case CB_OP_LT: case CB_OP_LE: case CB_OP_GT: case CB_OP_GE:
        case CB_OP_ADD: case CB_OP_SUB: case CB_OP_MUL: case CB_OP_DIV: case CB_OP_MOD: case CB_OP_IN:
        case CB_OP_STARTS_WITH: case CB_OP_ENDS_WITH: case CB_OP_CONTAINS: case CB_OP_AND: case CB_OP_OR:
        case CB_OP_HAS_INTERSECTION: case CB_OP_IS_SUBSET: case CB_OP_HIER_REL: case CB_OP_IN_SPLIT:
            if (d < 2) bad = true;
            nd = d - 1; break;
        case CB_OP_HIER_CA: nd = d - (I.ia == 0 ? 1 : 2); if (nd < 1) bad = true; break;
        case CB_OP_FN: if (I.ib < 1 || I.ib > 4 || d < (int)I.ib) bad = true; nd = d - ((int)I.ib - 1); break;
        case CB_OP_JF_KEEP: case CB_OP_JT_KEEP: if (d < 1) bad = true; set(rel(I.ic), d, l); target[rel(I.ic)] = true; break;
        case CB_OP_JMP: set(rel(I.ic), d, l); target[rel(I.ic)] = true; falls = false; break;
        case CB_OP_TERN:
            if (d < 1) bad = true;
            nd = d - 1;
            set(rel(I.ic), d - 1, l); target[rel(I.ic)] = true;
            set(rel(I.ib), d, l); target[rel(I.ib)] = true;
            break;
        case CB_OP_LOOP_INIT:
            if (d < 1 || (I.ib & 0xFF) >= CB_LOOP_MAP || l >= CB_MAX_LOOP_DEPTH) bad = true;
            set(rel(I.ic), d, l); target[rel(I.ic)] = true;      // not entered: the comprehension's value is pushed
            set(k + 1, d - 1, l + 1);
            falls = false;
            break;
        case CB_OP_LOOP_NEXT:
            if (d < 1 || (I.ib & 0xFF) >= CB_LOOP_MAP || l < 1) bad = true;
            set(rel(I.ic), d - 1, l); target[rel(I.ic)] = true;  // next element: back to the body
            set(k + 1, d, l - 1);
            falls = false;
            break;
        default: bad = true; break;   // TO_COND / COND_NOT inside a leaf, MKLIST, MKMAP, LOOP_PRED, RUNTIME_EDR, unknown
        }
        if (nd > CB_MAX_STACK) bad = true;
        if (falls && !bad) set(k + 1, nd, l);
    }
    if (bad || depth[n] != 1 || ldep[n] != 0) return "";
    auto S = [](int i) { return "s" + std::to_string(i); };
    auto lab = [](uint32_t k) { return "P" + std::to_string(k); };
    auto cst = [&](uint32_t k) -> std::string {
        if (k >= n_consts) { bad = true; return "mk_err()"; }
        const uint32_t *w = consts + 4 * k;    // {tag, pad, bits lo, bits hi}
        return "mk(" + hex(w[0]) + ", " + hex64((uint64_t)w[2] | (uint64_t)w[3] << 32) + ")";
    };
    auto slot = [&](uint32_t v) -> std::string {
        if (v >= n_slots) { bad = true; return "0ull"; }
        slot_used[v] = true;
        bool seen = false;
        for (uint32_t q : my_slots) seen |= q == v;
        if (!seen) my_slots.push_back(v);
        return "cols.slot(" + std::to_string(v) + "u)";
    };
    int maxd = 1;
    for (uint32_t k = 0; k <= n; k++) if (depth[k] > maxd) maxd = depth[k];
    std::string s = "    Val";
    for (int i = 0; i < maxd; i++) s += std::string(i ? ", " : " ") + S(i);
    s += ";\n    Loop L0, L1; int st_; (void)st_; (void)L0; (void)L1;\n";
    for (uint32_t k = 0; k < n; k++) {
        const Ins &I = P[lo + k];
        const int d = depth[k], l = ldep[k];
        if (target[k]) s += lab(k) + ":;\n";
        const std::string a = S(d - 1), a2 = S(d - 2), top = S(d);   // a: top of stack, a2: below it, top: next free
        s += "    ";
        switch (I.op) {
        case CB_OP_CONST: s += top + " = " + cst(I.ic) + ";"; break;
        case CB_OP_SLOT: s += top + " = decode_v64(" + slot(I.ic) + ", &st_);"; break;
        case CB_OP_HAS_SLOT: s += "decode_v64(" + slot(I.ic) + ", &st_); " + top + " = op_has_slot(st_);"; break;
        case CB_OP_PID: my_pid = true; s += top + " = mk(CB_T_STRING, c.pid);"; break;
        case CB_OP_NOW: s += top + " = mk(CB_T_TS, (uint64_t)c.b->now);"; break;
        case CB_OP_VAR: if (I.ia >= CB_MAX_VARS) bad = true; s += top + " = c.vars[" + std::to_string(I.ia) + "];"; break;
        case CB_OP_SELECT: s += a + " = op_select(c, " + a + ", " + hex(I.ic) + ");"; break;
        case CB_OP_HAS: s += a + " = op_has(c, " + a + ", " + hex(I.ic) + ");"; break;
        case CB_OP_INDEX: s += a2 + " = do_index(c, " + a2 + ", " + a + ");"; break;
        case CB_OP_EQ: case CB_OP_NE: case CB_OP_LT: case CB_OP_LE: case CB_OP_GT: case CB_OP_GE:
            s += a2 + " = do_cmp(c, " + std::to_string(I.op - CB_OP_EQ) + ", " + a2 + ", " + a + ");"; break;
        case CB_OP_ADD: case CB_OP_SUB: case CB_OP_MUL: case CB_OP_DIV: case CB_OP_MOD:
            s += a2 + " = do_arith(c, " + std::to_string(I.op) + ", " + a2 + ", " + a + ");"; break;
        case CB_OP_NEG: s += a + " = op_neg(" + a + ");"; break;
        case CB_OP_NOT: s += a + " = op_not(" + a + ");"; break;
        case CB_OP_IN: s += a2 + " = do_in(c, " + a2 + ", " + a + ");"; break;
        case CB_OP_SIZE: s += a + " = op_size(c, " + a + ");"; break;
        case CB_OP_STARTS_WITH: case CB_OP_ENDS_WITH: case CB_OP_CONTAINS:
            s += a2 + " = do_str2(c, " + std::to_string(I.op) + ", " + a2 + ", " + a + ");"; break;
        case CB_OP_JF_KEEP: s += "if (" + a + ".tag == CB_T_BOOL && " + a + ".u == 0) goto " + lab(I.ic - lo) + ";"; break;
        case CB_OP_JT_KEEP: s += "if (" + a + ".tag == CB_T_BOOL && " + a + ".u == 1) goto " + lab(I.ic - lo) + ";"; break;
        case CB_OP_AND: s += a2 + " = and_or(false, " + a2 + ", " + a + ");"; break;
        case CB_OP_OR: s += a2 + " = and_or(true, " + a2 + ", " + a + ");"; break;
        case CB_OP_JMP: s += "goto " + lab(I.ic - lo) + ";"; break;
        case CB_OP_TERN:
            s += "if (" + a + ".tag == CB_T_BOOL) { if (!" + a + ".u) goto " + lab(I.ic - lo) + "; } else { " + a + " = mk_err(); goto " + lab(I.ib - lo) + "; }";
            break;
        case CB_OP_HAS_INTERSECTION: s += a2 + " = do_set_pred(c, false, " + a2 + ", " + a + ");"; break;
        case CB_OP_IS_SUBSET: s += a2 + " = do_set_pred(c, true, " + a2 + ", " + a + ");"; break;
        case CB_OP_LOOP_INIT:
            s += "{ const Val r_ = " + a + "; if (!qloop_init(c, L" + std::to_string(l) + ", r_, " + std::to_string(I.ib & 0xFF) + ", " + (((I.ib >> 8) & 1) ? "true" : "false") +
                 ", " + std::to_string(I.ia) + ", &" + a + ")) goto " + lab(I.ic - lo) + "; }";
            break;
        case CB_OP_LOOP_NEXT:
            s += "{ const Val r_ = " + a + "; if (!qloop_next(c, L" + std::to_string(l - 1) + ", r_, " + std::to_string(I.ib & 0xFF) + ", " + (((I.ib >> 8) & 1) ? "true" : "false") +
                 ", " + std::to_string(I.ia) + ", &" + a + ")) goto " + lab(I.ic - lo) + "; }";
            break;
        case CB_OP_NOERR: s += a + " = mk_bool(" + a + ".tag != CB_T_ERR);"; break;
        case CB_OP_INT: s += a + " = conv_int(c, " + a + ");"; break;
        case CB_OP_UINT: s += a + " = conv_uint(c, " + a + ");"; break;
        case CB_OP_DOUBLE: s += a + " = op_double(c, " + a + ");"; break;
        case CB_OP_TIMESTAMP: s += a + " = op_timestamp(c, " + a + ");"; break;
        case CB_OP_DURATION: s += a + " = op_duration(c, " + a + ");"; break;
        case CB_OP_DYN: s += ";"; break;
        case CB_OP_CMP_SLOT_CONST: s += top + " = do_cmp(c, " + std::to_string(I.ia) + ", decode_v64(" + slot(I.ib) + ", &st_), " + cst(I.ic) + ");"; break;
        case CB_OP_CMP_SLOT_SLOT: s += "{ const Val x_ = decode_v64(" + slot(I.ib) + ", &st_); " + top + " = do_cmp(c, " + std::to_string(I.ia) + ", x_, decode_v64(" + slot(I.ic) + ", &st_)); }"; break;
        case CB_OP_CMP_SLOT_PID: my_pid = true; s += top + " = do_cmp(c, " + std::to_string(I.ia) + ", decode_v64(" + slot(I.ib) + ", &st_), mk(CB_T_STRING, c.pid));"; break;
        case CB_OP_IN_SLOT_CONST: s += top + " = do_in(c, decode_v64(" + slot(I.ib) + ", &st_), " + cst(I.ic) + ");"; break;
        case CB_OP_IN_CONST_SLOT: s += top + " = do_in(c, " + cst(I.ic) + ", decode_v64(" + slot(I.ib) + ", &st_));"; break;
        case CB_OP_IN_IP_RANGE: s += a + " = " + a + ".tag == CB_T_ERR ? mk_err() : do_in_ip_range(c, " + a + ", c.t->theap() + " + hex(I.ic) + ");"; break;
        case CB_OP_HIER_REL: s += a2 + " = op_hier_rel(c, " + hex(I.ia) + ", " + hex(I.ib) + ", " + hex(I.ic) + ", " + a2 + ", " + a + ");"; break;
        case CB_OP_TS_GET: s += a + " = op_ts_get(c, " + a + ", " + hex(I.ia) + ", " + hex(I.ib) + ", " + hex(I.ic) + ");"; break;
        case CB_OP_IN_SPLIT: s += a2 + " = op_in_split(c, " + a2 + ", " + a + ", " + hex(I.ib) + ");"; break;
        case CB_OP_HIER_SIZE: s += a + " = op_hier_size(c, " + a + ", " + hex(I.ib) + ");"; break;
        case CB_OP_HIER_CA:
            if (I.ia == 0) s += a2 + " = op_hier_ca2(c, " + a2 + ", " + a + ", " + hex(I.ib) + ", " + hex(I.ic) + ");";
            else s += S(d - 3) + " = op_hier_ca3(c, " + S(d - 3) + ", " + a2 + ", " + a + ", " + hex(I.ib) + ", " + hex(I.ic) + ");";
            break;
        case CB_OP_FN: {
            const int base = d - (int)I.ib;
            s += "{ Val a_[" + std::to_string(I.ib) + "] = {";
            for (uint32_t q = 0; q < I.ib; q++) s += std::string(q ? ", " : "") + S(base + (int)q);
            s += "}; " + S(base) + " = op_fn(c, " + hex(I.ia) + ", " + hex(I.ib) + ", a_); }";
            break;
        }
        case CB_OP_MATCHES: s += a + " = op_matches(c, " + a + ", " + hex(I.ic) + ");"; break;
        default: bad = true; break;
        }
        s += "\n";
    }
    if (target[n]) s += lab(n) + ":;\n";
    s += "    return cond_true(s0);\n";
    if (only_slot) *only_slot = my_slots.size() == 1 ? my_slots[0] : CB_NONE32;
    if (reads_pid) *reads_pid = my_pid;
    return bad ? std::string() : s;
}

// A whole condition program -> boolean formula over atoms ("" = does not qualify).  New atoms are appended to `at`.
inline std::string translate_program(const uint32_t *code_words, uint32_t code_off, uint32_t code_len, const uint32_t *consts, uint32_t n_consts, uint32_t n_slots, Atoms &at) {
    if (code_len == 0 || code_len > 4096) return "";
    std::vector<Ins> P(code_len);
    for (uint32_t i = 0; i < code_len; i++) {
        const uint32_t w0 = code_words[2 * (code_off + i)], w1 = code_words[2 * (code_off + i) + 1];
        P[i] = Ins{w0 & 0xFFu, (w0 >> 8) & 0xFFu, w0 >> 16, w1};
    }
    std::vector<std::string> st;    // the condition-level stack, symbolically
    uint32_t i = 0;
    bool ret = false;
    while (i < code_len && !ret) {
        const Ins &I = P[i];
        auto cond_level = [&](uint32_t op) { return op == CB_OP_AND || op == CB_OP_OR || op == CB_OP_COND_NOT || op == CB_OP_RET; };
        switch (I.op) {
        case CB_OP_RET: ret = true; continue;
        case CB_OP_AND: case CB_OP_OR: {
            if (st.size() < 2) return "";
            const std::string b = st.back(); st.pop_back();
            st.back() = "(" + st.back() + (I.op == CB_OP_AND ? " & " : " | ") + b + ")";
            i++;
            continue;
        }
        case CB_OP_COND_NOT: if (st.empty()) return ""; st.back() = "!" + st.back(); i++; continue;
        case CB_OP_JF_KEEP: case CB_OP_JT_KEEP:
            // condition-level short circuit: every leaf is evaluated anyway (no side effects), the formula is what matters
            if (st.empty() || I.ic <= i || I.ic > code_len) return "";
            i++;
            continue;
        default: break;
        }
        // a leaf starts here: it ends at the next TO_COND
        uint32_t e = i;
        while (e < code_len && P[e].op != CB_OP_TO_COND) e++;
        bool literal = false;
        if (I.op == CB_OP_CONST && i + 1 < code_len) {   // all[] / any[] / none[]: a bare constant at condition level
            const uint32_t nx = P[i + 1].op;
            if (cond_level(nx)) literal = true;
            if ((nx == CB_OP_JF_KEEP || nx == CB_OP_JT_KEEP) && (e >= code_len || P[i + 1].ic > e)) literal = true;
        }
        if (literal) {
            if (I.ic >= n_consts || consts[4 * I.ic] != CB_T_BOOL) return "";
            st.push_back(consts[4 * I.ic + 2] ? "true" : "false");
            i++;
            continue;
        }
        if (e >= code_len) return "";
        std::vector<uint32_t> key;
        for (uint32_t k = i; k < e; k++) {
            const Ins &J = P[k];
            const bool j = is_jump(J.op);
            key.push_back(J.op | J.ia << 8 | (J.op == CB_OP_TERN ? (J.ib - i) : J.ib) << 16);
            key.push_back(j ? J.ic - i : J.ic);
        }
        auto it = at.ids.find(key);
        if (it == at.ids.end()) {
            if (at.slot_used.size() < n_slots) at.slot_used.resize(n_slots, false);
            uint32_t one = CB_NONE32;
            bool pid = false;
            const std::string body = atom_source(P, i, e, consts, n_consts, n_slots, at.slot_used, &one, &pid);
            if (body.empty()) return "";
            const uint32_t id = (uint32_t)at.src.size();
            at.only_slot.push_back(one);
            at.reads_pid.push_back(pid);
            at.src.push_back("template <typename Cols>\nCB_HD bool uc_atom_" + std::to_string(id) + "(Ctx &c, const Cols &cols) {\n" + body + "}\n");
            it = at.ids.emplace(key, id).first;
        }
        st.push_back("a" + std::to_string(it->second));
        i = e + 1;
    }
    if (!ret || st.size() != 1) return "";
    return st[0];
}

// ---- unique-condition form (cb_uc.h / cb::eval_request_uc) -----------------------------------------------------------
// For tables whose blocks differ in shape the per-shape inlining above explodes; their DISTINCT conditions are few.
// generate_uc() emits `SpecConds`: load() pulls every attribute slot the conditions read into registers (all loads in
// flight at once, coalesced), operator() evaluates every distinct DNF term once (shared between conditions) and
// combines them into the request's condition word.  Rows stay data (4-byte records walked by the generic loop).
// uc: the compact image (cb_uc.h) and its layout.  "" = does not qualify (a distinct condition without flat form ...).
struct UcLimits {
    uint32_t max_terms = 512;   // distinct terms
};
struct UcSource {
    std::string src;            // "" = does not qualify
    uint32_t n_strpred = 0;     // string predicates served by the per-string pre-pass (BatchView::strpred)
    uint32_t n_atoms = 0;       // leaf programs translated to straight-line code (conditions without a flat form)
};
inline UcSource generate_uc(const uint8_t *uc_image, const uint32_t *off, uint32_t uc_conds_off, uint32_t n_uconds, uint32_t n_slots, uint32_t n_consts = 0,
                            const UcLimits lim = UcLimits()) {
    UcSource out;
    if (n_uconds == 0 || n_uconds > 127) return out;
    const uint32_t *const_words = reinterpret_cast<const uint32_t *>(uc_image + off[CB_SEC_CONSTS]);   // cb_const: {tag, pad, bits}
    Atoms atoms;
    std::vector<std::string> formula(n_uconds + 1);   // per distinct condition without flat form: boolean formula over atoms
    const uint32_t *uconds = reinterpret_cast<const uint32_t *>(uc_image + uc_conds_off);    // [n_uconds + 1] x {code_off, code_len, flat_off, flat_info}
    const uint32_t *code = reinterpret_cast<const uint32_t *>(uc_image + off[CB_SEC_CODE]);
    const uint64_t *consts = reinterpret_cast<const uint64_t *>(uc_image + off[CB_SEC_CONSTS_V64]);
    const uint64_t *theap = reinterpret_cast<const uint64_t *>(uc_image + off[CB_SEC_THEAP]);
    const uint32_t kUseMask = ~((uint32_t)(CB_TERM_LIT_F | CB_TERM_GROUP_END) << 8);
    std::map<std::vector<uint32_t>, uint32_t> term_ids;
    std::vector<std::vector<uint32_t>> terms;
    const uint32_t ns = n_slots ? n_slots : 1;
    std::vector<bool> slot_used(ns, false), slot_list(ns, false);
    auto is_slot_kind = [](uint32_t kind) { return kind == CB_OPK_SLOT || kind == CB_OPK_SLOT_ELEM || kind == CB_OPK_SLOT_SIZE; };
    auto use_slot = [&](uint32_t kind, uint32_t v) { if (is_slot_kind(kind) && v < ns) slot_used[v] = true; };
    auto v64_tag = [](uint64_t b) { const uint32_t top = (uint32_t)(b >> 48); return (top & 0xFFF0u) == 0xFFF0u ? (top & 0xFu) : 0u; };
    for (uint32_t u = 1; u <= n_uconds; u++) {
        const uint32_t *cd = uconds + 4 * u;
        if (cd[3] == 0 || ((cd[3] >> 16) & 0xFF) != CB_FLAT_DNF) {
            formula[u] = translate_program(code, cd[0], cd[1], const_words, n_consts, ns, atoms);
            if (formula[u].empty()) return out;
            continue;
        }
        for (uint32_t i = 0, nt = cd[3] & 0xFFFFu; i < nt; i++) {
            const uint32_t *w = code + 2 * (cd[2] + 2 * i);
            std::vector<uint32_t> key = {w[0] & kUseMask, w[1], w[2], w[3]};
            if (term_ids.emplace(key, (uint32_t)terms.size()).second) terms.push_back(key);
        }
    }
    if (terms.size() > lim.max_terms) return out;
    // per term: which form it takes in the specialised code
    //   'L' list registers (IN with a slot list, set predicates over two slot lists), 'P' string-predicate word, 'G' generic
    std::vector<char> form(terms.size(), 'G');
    std::vector<uint32_t> pred_of(terms.size(), 0);
    std::vector<uint32_t> pred_terms;   // term index of every string predicate
    for (uint32_t q = 0; q < terms.size(); q++) {
        const uint32_t *w = terms[q].data();
        const uint32_t op = w[0] & 0xFF, xk = (w[0] >> 16) & 0xFF, yk = w[0] >> 24;
        switch (op) {   // which operands are slots (bytecode._specialize_term shapes first)
        case CB_TERM_EQ_SS: case CB_TERM_ORD_SS: use_slot(CB_OPK_SLOT, w[1]); use_slot(CB_OPK_SLOT, w[2]); break;
        case CB_TERM_EQ_SC: case CB_TERM_EQ_SP: case CB_TERM_ORD_SC: case CB_TERM_IN_SC: use_slot(CB_OPK_SLOT, w[1]); break;
        case CB_TERM_IN_SS: use_slot(CB_OPK_SLOT, w[1]); use_slot(CB_OPK_SLOT, w[2]); if (w[2] < ns) { slot_list[w[2]] = true; form[q] = 'L'; } break;
        case CB_TERM_IN_CS: use_slot(CB_OPK_SLOT, w[2]); if (w[2] < ns) { slot_list[w[2]] = true; form[q] = 'L'; } break;
        default:
            use_slot(xk, w[1]);
            if (op != CB_TERM_HAS) use_slot(yk, w[2]);
            if (op == CB_TERM_IN && yk == CB_OPK_SLOT && w[2] < ns) { slot_list[w[2]] = true; form[q] = 'L'; }
            if ((op == CB_TERM_INTERSECTS || op == CB_TERM_SUBSET) && xk == CB_OPK_SLOT && yk == CB_OPK_SLOT && w[1] < ns && w[2] < ns) {
                slot_list[w[1]] = slot_list[w[2]] = true;
                form[q] = 'L';
            }
            if ((op == CB_TERM_STARTS || op == CB_TERM_ENDS || op == CB_TERM_CONTAINS) && xk == CB_OPK_SLOT && yk == CB_OPK_CONST &&
                v64_tag(consts[w[2]]) == CB_V64_STRING && pred_terms.size() < 32) {
                form[q] = 'P';
                pred_of[q] = (uint32_t)pred_terms.size();
                pred_terms.push_back(q);
            }
            break;
        }
    }
    // Leaf programs over ONE attribute slot (and nothing else of the request) are functions of that slot's value: for a
    // string value the pre-pass evaluates them once per distinct dictionary string -- timestamp(<claim>) > now(),
    // "x" in <claim>.split(" ") ... -- and the request kernel reads two bits: the value, and "evaluate in place" (the
    // pre-pass met a value the device forms cannot hold: the in-place evaluation then raises it for this request).
    std::vector<int> atom_pred(atoms.src.size(), -1);
    uint32_t n_pred_bits = (uint32_t)pred_terms.size();
    for (uint32_t a = 0; a < atoms.src.size(); a++)
        if (atoms.only_slot[a] != CB_NONE32 && !atoms.reads_pid[a] && n_pred_bits + 2 <= 32) { atom_pred[a] = (int)n_pred_bits; n_pred_bits += 2; }
    auto sl = [](uint32_t v) { return "cols.slot(" + std::to_string(v) + "u)"; };
    auto term_code = [&](uint32_t q) -> std::string {
        const uint32_t *w = terms[q].data();
        const uint32_t op = w[0] & 0xFF, xk = (w[0] >> 16) & 0xFF;
        if (form[q] == 'P') return "strpred_tri(b, " + sl(w[1]) + ", " + std::to_string(pred_of[q]) + "u)";
        if (form[q] == 'L') {
            if (op == CB_TERM_IN_CS) return "list_in_tri(" + hex64(consts[w[1]]) + ", cols.l" + std::to_string(w[2]) + ", slow)";
            if (op == CB_TERM_IN_SS) return "list_in_tri(" + sl(w[1]) + ", cols.l" + std::to_string(w[2]) + ", slow)";
            if (op == CB_TERM_IN)
                return "list_in_tri(term_operand(t, b, cols, pid, " + hex(xk) + ", " + hex(w[1]) + ", " + hex(w[3] & 0xFFFFu) + "), cols.l" + std::to_string(w[2]) + ", slow)";
            return std::string("list_set_tri(") + (op == CB_TERM_SUBSET ? "true" : "false") + ", cols.l" + std::to_string(w[1]) + ", cols.l" + std::to_string(w[2]) + ", slow)";
        }
        return term_expr(w, consts, theap);
    };
    // Slots read by the flat terms live in registers (all loads in flight at once); slots only the leaf programs read --
    // and every slot once the table reads more than kMaxRegSlots of them -- are loaded where they are used (L1-allocating loads).
    const uint32_t kMaxRegSlots = 16;
    const bool have_atoms = !atoms.src.empty();
    atoms.slot_used.resize(ns, false);
    uint32_t n_reg = 0;
    for (uint32_t v = 0; v < ns; v++) n_reg += slot_used[v] || slot_list[v];
    if (n_reg > kMaxRegSlots)
        for (uint32_t v = 0; v < ns; v++) if (!slot_list[v]) slot_used[v] = false;
    std::string s;
    s += "// generated by cb_specialize.h (generate_uc) from the loaded table: every distinct condition, straight-line\n";
    s += "namespace cb {\n";
    for (const std::string &a : atoms.src) s += a;
    s += "struct SpecRegs {\n    CachedCols g;\n";
    for (uint32_t v = 0; v < ns; v++)
        if (slot_used[v] || slot_list[v]) s += "    uint64_t s" + std::to_string(v) + ";\n";
    for (uint32_t v = 0; v < ns; v++)
        if (slot_list[v]) s += "    ListRegs l" + std::to_string(v) + ";\n";
    s += "    CB_HD uint64_t slot(uint32_t v) const {\n        switch (v) {\n";
    for (uint32_t v = 0; v < ns; v++)
        if (slot_used[v] || slot_list[v]) s += "        case " + std::to_string(v) + "u: return s" + std::to_string(v) + ";\n";
    s += "        default: return v < " + std::to_string(ns) + "u ? g.slot(v) : (uint64_t)(CB_V64_BOX_BASE | CB_V64_ERROR) << 48;\n        }\n    }\n};\n";
    s += "struct SpecConds {\n    static constexpr uint32_t n_strpred = " + std::to_string(n_pred_bits) + "u;\n";
    s += std::string("    static constexpr int kForm = ") + (n_uconds <= 31 ? "CB_UC_FORM_MASK32" : n_uconds <= 63 ? "CB_UC_FORM_MASK64" : "CB_UC_FORM_INDEX") + ";   // how the rows name their conditions\n";
    s += std::string("    static constexpr bool kPrograms = ") + (have_atoms ? "true" : "false") + ";   // leaf programs: needs the value helpers of cb_core.h\n";
    s += "    template <typename Cols>\n    CB_HD SpecRegs load(const TableView t, const BatchView &b, const Cols &c) const {\n        SpecRegs r;\n        r.g.b = c.b; r.g.n = c.n;\n";
    for (uint32_t v = 0; v < ns; v++)
        if (slot_used[v] || slot_list[v]) s += "        r.s" + std::to_string(v) + " = c.slot(" + std::to_string(v) + "u);\n";
    for (uint32_t v = 0; v < ns; v++)
        if (slot_list[v]) s += "        r.l" + std::to_string(v) + " = list_load(t, b, r.s" + std::to_string(v) + ");\n";
    s += "        return r;\n    }\n";
    s += "    // the predicate word of one string (pre-pass over the string dictionary; bit p = predicate p holds)\n";
    s += "    CB_HD uint32_t strpred(const TableView t, const BatchView &b, uint32_t id) const {\n";
    s += "        OneCols cols; cols.x = ((uint64_t)(CB_V64_BOX_BASE | CB_V64_STRING) << 48) | id;\n        bool slow = false; uint32_t bits = 0u; const uint32_t pid = 0u;\n";
    for (uint32_t p = 0; p < pred_terms.size(); p++) {
        const uint32_t *w = terms[pred_terms[p]].data();
        s += "        bits |= (uint32_t)(term_tri(t, b, cols, pid, U4{" + hex(w[0]) + ", 0x0u, " + hex(w[2]) + ", " + hex(w[3]) + "}, slow) == TRI_T) << " + std::to_string(p) + ";\n";
    }
    {
        bool any = false;
        for (int p : atom_pred) any |= p >= 0;
        if (any) {
            s += "        Ctx c; c.t = &t; c.b = &b; c.req = 0; c.pid = 0; c.edr = 0;\n";
            for (uint32_t a = 0; a < atoms.src.size(); a++)
                if (atom_pred[a] >= 0)
                    s += "        c.unsupported = 0; c.scr_used = 0; bits |= (uint32_t)uc_atom_" + std::to_string(a) + "(c, cols) << " + std::to_string(atom_pred[a]) +
                         "; bits |= (uint32_t)(c.unsupported != 0) << " + std::to_string(atom_pred[a] + 1) + ";\n";
        }
    }
    s += "        (void)slow; (void)pid;\n        return bits;\n    }\n";
    s += "    CB_HD CondWord operator()(const TableView t, const BatchView &b, const SpecRegs &cols, uint32_t pid, uint64_t n, bool &slow) const {\n";
    for (uint32_t q = 0; q < terms.size(); q++)
        s += "        const int q" + std::to_string(q) + " = " + term_code(q) + ";\n";
    if (have_atoms) {
        s += "        Ctx c; c.t = &t; c.b = &b; c.req = n; c.pid = pid; c.unsupported = 0; c.edr = 0; c.scr_used = 0;\n";
        for (uint32_t a = 0; a < atoms.src.size(); a++) {
            const std::string A = std::to_string(a);
            if (atom_pred[a] >= 0) {
                const std::string P = std::to_string(atom_pred[a]), V = sl(atoms.only_slot[a]);
                s += "        bool a" + A + ";\n        {\n            const uint64_t x = " + V + ";\n            const uint32_t w = v64_tag(x) == CB_V64_STRING && !v64_bad(x) ? ldg(b.strpred + (uint32_t)(x & 0xFFFFFFFFu)) >> " + P + " : 2u;\n";
                s += "            if (w & 2u) { c.scr_used = 0; a" + A + " = uc_atom_" + A + "(c, cols); } else a" + A + " = (w & 1u) != 0;\n        }\n";
            } else {
                s += "        c.scr_used = 0; const bool a" + A + " = uc_atom_" + A + "(c, cols);\n";
            }
        }
        s += "        slow |= c.unsupported != 0;   // a value the device forms cannot hold: the general kernel reports it\n";
    } else s += "        (void)n;\n";
    s += "        CondWord val; val.lo = 1ull; val.hi = 0ull;\n";
    for (uint32_t u = 1; u <= n_uconds; u++) {
        const uint32_t *cd = uconds + 4 * u;
        const std::string word = u < 64 ? "val.lo" : "val.hi", sh = std::to_string(u & 63u);
        if (!formula[u].empty()) {
            s += "        " + word + " |= (uint64_t)(" + formula[u] + ") << " + sh + ";   // distinct condition " + std::to_string(u) + " (program)\n";
            continue;
        }
        const uint32_t nt = cd[3] & 0xFFFFu, negate = (cd[3] >> 24) & 1u;
        s += "        {   // distinct condition " + std::to_string(u) + "\n            bool any = false, group = true;\n";
        for (uint32_t i = 0; i < nt; i++) {
            const uint32_t *w = code + 2 * (cd[2] + 2 * i);
            const uint32_t flags = (w[0] >> 8) & 0xFFu;
            const uint32_t q = term_ids[{w[0] & kUseMask, w[1], w[2], w[3]}];
            s += "            group &= term_lit(q" + std::to_string(q) + ", " + hex(flags) + ");\n";
            if (flags & CB_TERM_GROUP_END) s += "            any |= group; group = true;\n";
        }
        s += "            " + word + std::string(" |= (uint64_t)(any != ") + (negate ? "true" : "false") + ") << " + sh + ";\n        }\n";
    }
    s += "        return val;\n    }\n};\n}  // namespace cb\n";
    out.src = s;
    out.n_strpred = n_pred_bits;
    out.n_atoms = (uint32_t)atoms.src.size();
    return out;
}

}  // namespace cbspec
