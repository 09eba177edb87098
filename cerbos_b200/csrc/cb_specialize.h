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
};
inline UcSource generate_uc(const uint8_t *uc_image, const uint32_t *off, uint32_t uc_conds_off, uint32_t n_uconds, uint32_t n_slots, const UcLimits lim = UcLimits()) {
    UcSource out;
    if (n_uconds == 0 || n_uconds > 63) return out;
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
        if (cd[3] == 0 || ((cd[3] >> 16) & 0xFF) != CB_FLAT_DNF) return out;
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
    std::string s;
    s += "// generated by cb_specialize.h (generate_uc) from the loaded table: every distinct condition, straight-line\n";
    s += "namespace cb {\nstruct SpecRegs {\n";
    for (uint32_t v = 0; v < ns; v++)
        if (slot_used[v]) s += "    uint64_t s" + std::to_string(v) + ";\n";
    for (uint32_t v = 0; v < ns; v++)
        if (slot_list[v]) s += "    ListRegs l" + std::to_string(v) + ";\n";
    s += "    CB_HD uint64_t slot(uint32_t v) const {\n        switch (v) {\n";
    for (uint32_t v = 0; v < ns; v++)
        if (slot_used[v]) s += "        case " + std::to_string(v) + "u: return s" + std::to_string(v) + ";\n";
    s += "        default: return (uint64_t)(CB_V64_BOX_BASE | CB_V64_ERROR) << 48;\n        }\n    }\n};\n";
    s += "struct SpecConds {\n    static constexpr uint32_t n_strpred = " + std::to_string(pred_terms.size()) + "u;\n";
    s += std::string("    static constexpr bool kVal32 = ") + (n_uconds <= 31 ? "true" : "false") + ";   // the condition word fits 32 bits\n";
    s += "    template <typename Cols>\n    CB_HD SpecRegs load(const TableView t, const BatchView &b, const Cols &c) const {\n        SpecRegs r;\n";
    for (uint32_t v = 0; v < ns; v++)
        if (slot_used[v]) s += "        r.s" + std::to_string(v) + " = c.slot(" + std::to_string(v) + "u);\n";
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
    s += "        (void)slow; (void)pid;\n        return bits;\n    }\n";
    s += "    CB_HD uint64_t operator()(const TableView t, const BatchView &b, const SpecRegs &cols, uint32_t pid, uint64_t, bool &slow) const {\n";
    for (uint32_t q = 0; q < terms.size(); q++)
        s += "        const int q" + std::to_string(q) + " = " + term_code(q) + ";\n";
    s += "        uint64_t val = 1ull;\n";
    for (uint32_t u = 1; u <= n_uconds; u++) {
        const uint32_t *cd = uconds + 4 * u;
        const uint32_t nt = cd[3] & 0xFFFFu, negate = (cd[3] >> 24) & 1u;
        s += "        {   // distinct condition " + std::to_string(u) + "\n            bool any = false, group = true;\n";
        for (uint32_t i = 0; i < nt; i++) {
            const uint32_t *w = code + 2 * (cd[2] + 2 * i);
            const uint32_t flags = (w[0] >> 8) & 0xFFu;
            const uint32_t q = term_ids[{w[0] & kUseMask, w[1], w[2], w[3]}];
            s += "            group &= term_lit(q" + std::to_string(q) + ", " + hex(flags) + ");\n";
            if (flags & CB_TERM_GROUP_END) s += "            any |= group; group = true;\n";
        }
        s += std::string("            val |= (uint64_t)(any != ") + (negate ? "true" : "false") + ") << " + std::to_string(u) + ";\n        }\n";
    }
    s += "        return val;\n    }\n};\n}  // namespace cb\n";
    out.src = s;
    out.n_strpred = (uint32_t)pred_terms.size();
    return out;
}

}  // namespace cbspec
