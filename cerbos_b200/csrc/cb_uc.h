// cb_uc.h -- host-side builder of the "unique condition" image of a loaded table.
//
// The flattened table (cerbos_b200/table/flatten.py) keeps one condition list per policy block, the way the reference
// keeps one compiled condition per rule (ruletable.go:105-416).  Across a policy set the same conditions recur: shared
// derived roles, the same ownership / tenancy test on every resource kind.  This builder
//   * numbers the DISTINCT conditions of the table 1..U (same DNF term list, or -- no flat form -- same bytecode program),
//   * rewrites every row to {original index, role, effect, mask of the condition bits it needs} (16 bytes, DENY rows first per block),
//   * copies only the sections the unique-condition kernels read into a compact image (C3: 49 KB blob -> ~10 KB),
// so that a kernel can evaluate every distinct condition of a request once, with all lanes in lock step, and walk the
// rows as mask algebra (cb::eval_request_uc).  Built once per cgpu_table_load; the Python blob format is unchanged.
//
// Host-only, no CUDA dependencies: tests/hostsim uses it too.
#pragma once
#include <stdint.h>

#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "cb_core.h"

namespace cbuc {

constexpr uint32_t kMaxUconds = 127;      // bit 0 of the condition word is "no condition"
constexpr uint32_t kMaxMaskUconds = 63;   // up to here rows carry need MASKS; above, the two condition NUMBERS (cb_core.h: CB_UC_FORM_INDEX)

struct Image {
    bool ok = false;
    std::string why;                    // when !ok
    std::vector<uint8_t> bytes;         // compact image (16-byte aligned sections)
    cb::TableLayout lay{};              // offsets into `bytes` + the dims of the source layout
    uint32_t n_uconds = 0, n_flat = 0;  // distinct conditions; how many of them have a flat (DNF) form
    // programs among the conditions, or rows in index form: only the run-time specialised kernel can evaluate the image
    bool needs_spec() const { return n_flat != n_uconds || n_uconds > kMaxMaskUconds; }
    uint32_t n_gids = 0, n_gids_flat = 0;   // table conditions (per-block lists), and how many of them have a flat form
    std::vector<uint32_t> ucond_of_gid; // table condition id -> distinct condition number (1..U)
};

// image: the table image (blob bytes); off / len: section offset and byte length by section id; lay: its layout
inline Image build(const uint8_t *image, const uint32_t *off, const uint64_t *len, const uint32_t *meta, const cb::TableLayout &lay) {
    Image out;
    const uint32_t n_blocks = meta[CB_META_N_BLOCKS], n_rows = meta[CB_META_N_ROWS], n_conds = meta[CB_META_N_CONDS];
    if (n_blocks == 0) { out.why = "no policy blocks"; return out; }
    if (lay.nR > 64) { out.why = "more than 64 roles"; return out; }
    const uint32_t *blocks = reinterpret_cast<const uint32_t *>(image + off[CB_SEC_BLOCKS]);
    const uint32_t *rows = reinterpret_cast<const uint32_t *>(image + off[CB_SEC_ROWS]);
    const uint32_t *conds = reinterpret_cast<const uint32_t *>(image + off[CB_SEC_CONDS]);
    if (len[CB_SEC_BLOCKS] < (uint64_t)n_blocks * 16 || len[CB_SEC_ROWS] < (uint64_t)n_rows * 16 || len[CB_SEC_CONDS] < (uint64_t)n_conds * 16) {
        out.why = "section sizes do not match META";
        return out;
    }
    // distinct conditions
    std::map<std::tuple<uint32_t, uint32_t, uint32_t>, uint32_t> ids;
    std::vector<uint32_t> ucond_rec;   // 4 words per distinct condition, entry 0 unused
    ucond_rec.assign(4, 0);
    out.ucond_of_gid.assign(n_conds, 0);
    for (uint32_t g = 0; g < n_conds; g++) {
        const uint32_t *cd = conds + 4 * g;   // {code_off, code_len, flat_off, flat_info}
        const bool flat = cd[3] != 0 && ((cd[3] >> 16) & 0xFF) == CB_FLAT_DNF;
        const auto key = flat ? std::make_tuple(1u, cd[2], cd[3]) : std::make_tuple(0u, cd[0], cd[1]);
        auto it = ids.find(key);
        if (it == ids.end()) {
            const uint32_t u = (uint32_t)ids.size() + 1;
            if (u > kMaxUconds) { out.why = "more than 127 distinct conditions"; return out; }
            it = ids.emplace(key, u).first;
            ucond_rec.insert(ucond_rec.end(), {cd[0], cd[1], flat ? cd[2] : 0u, flat ? cd[3] : 0u});
            out.n_flat += flat;
        }
        out.ucond_of_gid[g] = it->second;
        out.n_gids++;
        out.n_gids_flat += flat;
    }
    out.n_uconds = (uint32_t)ids.size();
    // Conditions without a flat form stay in the image as programs: only the run-time specialised kernel evaluates them
    // (cb_specialize.h turns their bytecode into straight-line code); the generic unique-condition kernels defer such
    // requests, so the library uses the image of a table with programs only once its specialised kernel is loaded.
    // rows: DENY rows first inside every block (within a scope every matching row is evaluated and DENY beats ALLOW,
    // ruletable.go:1083-1118, so the order of rows inside a block is free); 16 bytes each, see cb_core.h
    std::vector<uint32_t> urows(4 * (size_t)(n_rows ? n_rows : 1), 0);
    std::vector<uint32_t> ublocks(blocks, blocks + 4 * (size_t)n_blocks);   // {row_start, n_rows, DENY rows, 0}
    const bool index_form = out.n_uconds > kMaxMaskUconds;
    for (uint32_t b = 0; b < n_blocks; b++) {
        const uint32_t *bl = blocks + 4 * b;   // {row_start, n_rows, cond_base, n_conds}
        if ((uint64_t)bl[0] + bl[1] > n_rows || (uint64_t)bl[2] + bl[3] > n_conds) { out.why = "block out of range"; return out; }
        uint32_t at = bl[0], n_deny = 0;
        for (int pass = 0; pass < 2; pass++) {
            for (uint32_t r = 0; r < bl[1]; r++) {
                const uint32_t *row = rows + 4 * (bl[0] + r);
                const uint32_t role = row[0] & 0xFFFFu, c = row[0] >> 16, dc = row[1] & 0xFFFFu, effect = row[2] & 0xFFu;
                if ((effect == CB_EFFECT_DENY) != (pass == 0)) continue;
                if ((c && c > bl[3]) || (dc && dc > bl[3])) { out.why = "row condition out of range"; return out; }
                if (role != CB_ROLE_ANY && role >= 64) { out.why = "role id out of range"; return out; }
                const uint32_t uc = c ? out.ucond_of_gid[bl[2] + c - 1] : 0, udc = dc ? out.ucond_of_gid[bl[2] + dc - 1] : 0;
                const uint64_t need = index_form ? 0ull : (1ull | 1ull << uc | 1ull << udc);
                uint32_t *u = urows.data() + 4 * (size_t)at++;
                u[0] = bl[0] + r;
                u[1] = (role == CB_ROLE_ANY ? 0xFFu : role) | effect << 8;
                u[2] = index_form ? (uc | udc << 8) : (uint32_t)need;
                u[3] = index_form ? 0u : (uint32_t)(need >> 32);
                n_deny += pass == 0;
            }
        }
        ublocks[4 * b + 2] = n_deny;
        ublocks[4 * b + 3] = 0;
    }
    // compact image: the sections the unique-condition kernels (and the interpreter they may call) read
    out.lay = lay;
    for (auto &o : out.lay.off) o = 0;
    auto append = [&](const void *p, uint64_t n) {
        const uint32_t at = (uint32_t)out.bytes.size();
        out.bytes.resize((out.bytes.size() + n + 15) & ~(size_t)15, 0);
        if (n) memcpy(out.bytes.data() + at, p, n);
        return at;
    };
    append("CBUC", 4);   // offset 0 stays unused: a zero offset means "section not present"
    out.lay.off[CB_SEC_BLOCKS] = append(ublocks.data(), ublocks.size() * 4);
    for (int id : {CB_SEC_SCOPE_PARENT, CB_SEC_SCOPE_FLAGS, CB_SEC_RES_BLOCK_MAP, CB_SEC_CODE, CB_SEC_CONSTS, CB_SEC_CONSTS_V64, CB_SEC_THEAP,
                   CB_SEC_STR_OFF, CB_SEC_STR_BYTES})
        out.lay.off[id] = append(image + off[id], len[id]);
    out.lay.uc_conds_off = append(ucond_rec.data(), ucond_rec.size() * 4);
    out.lay.uc_rows_off = append(urows.data(), urows.size() * 4);
    out.lay.n_uconds = out.n_uconds;
    out.lay.theap_words = (uint32_t)(len[CB_SEC_THEAP] / 8);
    out.lay.image_bytes = (uint32_t)out.bytes.size();
    out.ok = true;
    return out;
}

}  // namespace cbuc
