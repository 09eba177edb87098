// tests/hostsim/hostsim.cpp -- DEBUG AID (test infrastructure, never shipped, never timed).
// Compiles the kernels' per-request core (cerbos_b200/csrc/cb_core.h) for the host so its logic can be
// checked against the oracles in this GPU-less container before a gpurun call is spent.  The product
// library (libcerbos_b200.so) contains no such path: without a CUDA device cgpu_init fails.
#include <cstdint>
#include <cstring>
#include <vector>

#include "cb_core.h"
#include "cb_specialize.h"

#ifdef HOSTSIM_SPEC
#include "spec_gen.inc"   // generated for one table by hostsim_generate (tests/test_specialize.py)
typedef cb::SpecBlocks HostBlocks;
#define HOSTSIM_ENTRY hostsim_check_spec
#else
typedef cb::GenericBlocks HostBlocks;
#define HOSTSIM_ENTRY hostsim_check
#endif

extern "C" int HOSTSIM_ENTRY(const void *blob, uint64_t blob_len, uint64_t n, uint32_t max_actions, int64_t now, uint32_t flags,
                             const void *const *cols, const uint64_t *col_bytes, uint8_t *bitmap, int mode) {
    const cb_blob_header *h = static_cast<const cb_blob_header *>(blob);
    if (h->magic != CB_MAGIC || h->version != CB_VERSION) return -1;
    const cb_section_desc *sd = reinterpret_cast<const cb_section_desc *>(static_cast<const char *>(blob) + sizeof(cb_blob_header));
    const uint8_t *base = static_cast<const uint8_t *>(blob);
    uint64_t off[128] = {0};
    for (uint32_t i = 0; i < h->n_sections; i++) if (sd[i].id < 128) off[sd[i].id] = sd[i].offset;
    const uint32_t *meta = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_META]);
    cb::TableLayout lay;
    memset(&lay, 0, sizeof(lay));
    for (int i = 0; i < 28; i++) lay.off[i] = (uint32_t)off[i];
    lay.nV = meta[CB_META_N_VERSIONS]; lay.nRP = meta[CB_META_N_RESPATS]; lay.nS = meta[CB_META_N_SCOPES]; lay.nP = meta[CB_META_N_PRINCIPALS];
    lay.nR = meta[CB_META_N_ROLES]; lay.nAP = meta[CB_META_N_APATS]; lay.nT = meta[CB_META_N_STRINGS]; lay.n_slots = meta[CB_META_N_SLOTS];
    lay.n_rows = meta[CB_META_N_ROWS] ? meta[CB_META_N_ROWS] : 1;
    lay.has_role_policies = meta[CB_META_HAS_ROLE_POLICIES]; lay.has_parent_roles = meta[CB_META_HAS_PARENT_ROLES];
    lay.has_principal_policies = meta[CB_META_HAS_PRINCIPAL_POLICIES];
    cb::TableView t;
    t.base = base; t.L = &lay;

    cb::BatchView b;
    b.hdr0 = static_cast<const cb_hdr0 *>(cols[0]); b.hdr1 = static_cast<const cb_hdr1 *>(cols[1]);
    b.roles = static_cast<const uint32_t *>(cols[2]); b.slots = static_cast<const uint64_t *>(cols[3]);
    b.heap = static_cast<const uint64_t *>(cols[4]); b.bstr_off = static_cast<const uint32_t *>(cols[5]);
    b.bstr_bytes = static_cast<const uint8_t *>(cols[6]); b.class_off = static_cast<const uint32_t *>(cols[7]);
    b.class_pats = static_cast<const uint32_t *>(cols[8]); b.aset_k = static_cast<const uint32_t *>(cols[9]);
    b.aset_spread = static_cast<const uint64_t *>(cols[10]);
    b.row_am = static_cast<const uint64_t *>(cols[11]);
    b.n_rows = meta[CB_META_N_ROWS] ? meta[CB_META_N_ROWS] : 1;
    b.stride = n; b.first = 0; b.count = n;
    b.role_cols = (uint32_t)(col_bytes[2] / (4 * n)); b.n_asets = (uint32_t)(col_bytes[9] / 4);
    uint32_t km = max_actions ? max_actions : 1;
    b.kc = 64 / b.role_cols; if (b.kc > km) b.kc = km;
    b.n_pass = (km + b.kc - 1) / b.kc; b.max_actions = km; b.kbytes = (km + 7) / 8; b.flags = flags; b.now = now;
    cb::finish_batch_view(b);
    uint32_t status = 0;
    // mode 0: what the library would pick; 1: force the general 64-bit body; 2: general 32-bit body;
    // 3: the lean body reading a tile of the columns staged the way the kernel's TMA copies lay it out in shared memory
    const bool narrow = b.n_pass == 1 && (uint64_t)km * b.role_cols <= 32;
    uint32_t rcp = 1; while (rcp < b.role_cols) rcp <<= 1;
    const bool fast = narrow && !lay.has_principal_policies && !lay.has_role_policies && !lay.has_parent_roles &&
                      meta[CB_META_DIRECT_KINDS] && (uint64_t)lay.nR * rcp <= 64 && b.kbytes <= 4;
    if (mode == 3 && fast) {
        std::vector<uint64_t> tile(cb::tile_cols_bytes(b.role_cols, lay.n_slots) / 8 + 2);
        uint8_t *tb = reinterpret_cast<uint8_t *>(tile.data());
        for (uint64_t t0 = 0; t0 < n; t0 += cb::CB_TILE) {
            const uint64_t cnt = n - t0 < cb::CB_TILE ? n - t0 : cb::CB_TILE;
            memcpy(tb, b.hdr0 + t0, cnt * 16);
            memcpy(tb + cb::CB_TILE * 16, b.hdr1 + t0, cnt * 8);
            for (uint32_t i = 0; i < b.role_cols; i++) memcpy(tb + cb::CB_TILE * 24 + i * cb::CB_TILE * 4, b.roles + i * b.stride + t0, cnt * 4);
            const uint32_t so = cb::CB_TILE * (24 + 4 * b.role_cols);
            for (uint32_t v = 0; v < lay.n_slots; v++) memcpy(tb + so + v * cb::CB_TILE * 8, b.slots + v * b.stride + t0, cnt * 8);
            for (uint32_t j = 0; j < cnt; j++) {
                cb::TileCols tc; tc.base = tb; tc.tid = j; tc.slots_off = so; tc.aset_k_s = b.aset_k; tc.row_am_s = b.row_am; tc.res_s = nullptr;
                if (cb::eval_request_fast(t, b, tc, t0 + j, bitmap, nullptr, HostBlocks())) cb::eval_request_general(t.base, t.L, &b, t0 + j, bitmap, nullptr, &status);
            }
        }
        return status ? -2 : 0;
    }
    for (uint64_t i = 0; i < n; i++) {
        if (mode == 0 && fast) {
            cb::GlobalCols gc; gc.b = &b; gc.n = i;
            if (cb::eval_request_fast(t, b, gc, i, bitmap, nullptr, HostBlocks())) cb::eval_request_general(t.base, t.L, &b, i, bitmap, nullptr, &status);
        }
        else if (mode == 2 && narrow) cb::eval_request<uint32_t>(t, b, i, bitmap, nullptr, &status);
        else cb::eval_request<uint64_t>(t, b, i, bitmap, nullptr, &status);
    }
    return status ? -2 : 0;
}

#ifndef HOSTSIM_SPEC
// the table-specialised block evaluators the library would hand to NVRTC (source text; "" if the table does not qualify)
extern "C" int64_t hostsim_generate(const void *blob, uint64_t blob_len, char *out, uint64_t cap) {
    const cb_blob_header *h = static_cast<const cb_blob_header *>(blob);
    if (blob_len < sizeof(*h) || h->magic != CB_MAGIC || h->version != CB_VERSION) return -1;
    const cb_section_desc *sd = reinterpret_cast<const cb_section_desc *>(static_cast<const char *>(blob) + sizeof(cb_blob_header));
    uint32_t off[128] = {0};
    for (uint32_t i = 0; i < h->n_sections; i++) if (sd[i].id < 128) off[sd[i].id] = (uint32_t)sd[i].offset;
    const uint8_t *base = static_cast<const uint8_t *>(blob);
    std::string src = cbspec::generate(base, off, reinterpret_cast<const uint32_t *>(base + off[CB_SEC_META]));
    if (src.size() + 1 > cap) return -(int64_t)src.size() - 2;
    memcpy(out, src.c_str(), src.size() + 1);
    return (int64_t)src.size();
}
#endif
