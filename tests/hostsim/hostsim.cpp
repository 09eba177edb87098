// tests/hostsim/hostsim.cpp -- DEBUG AID (test infrastructure, never shipped, never timed).
// Compiles the kernels' per-request core (cerbos_b200/csrc/cb_core.h) for the host so its logic can be
// checked against the oracles in this GPU-less container before a gpurun call is spent.  The product
// library (libcerbos_b200.so) contains no such path: without a CUDA device cgpu_init fails.
#include <cstdint>
#include <cstring>

#include "cb_core.h"

extern "C" int hostsim_check(const void *blob, uint64_t blob_len, uint64_t n, uint32_t max_actions, int64_t now, uint32_t flags,
                             const void *const *cols, const uint64_t *col_bytes, uint8_t *bitmap) {
    const cb_blob_header *h = static_cast<const cb_blob_header *>(blob);
    if (h->magic != CB_MAGIC || h->version != CB_VERSION) return -1;
    const cb_section_desc *sd = reinterpret_cast<const cb_section_desc *>(static_cast<const char *>(blob) + sizeof(cb_blob_header));
    const uint8_t *base = static_cast<const uint8_t *>(blob);
    uint64_t off[128] = {0};
    for (uint32_t i = 0; i < h->n_sections; i++) if (sd[i].id < 128) off[sd[i].id] = sd[i].offset;
    const uint32_t *meta = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_META]);
    cb::TableView t;
    t.scope_parent = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_SCOPE_PARENT]);
    t.scope_flags = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_SCOPE_FLAGS]);
    t.res_block_map = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_RES_BLOCK_MAP]);
    t.res_exists = base + off[CB_SEC_RES_EXISTS];
    t.prin_block_map = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_PRIN_BLOCK_MAP]);
    t.prin_exists = base + off[CB_SEC_PRIN_EXISTS];
    t.prin_of_string = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_PRIN_OF_STRING]);
    t.blocks = reinterpret_cast<const cb_block *>(base + off[CB_SEC_BLOCKS]);
    t.rows = reinterpret_cast<const cb_row *>(base + off[CB_SEC_ROWS]);
    t.conds = reinterpret_cast<const cb_cond *>(base + off[CB_SEC_CONDS]);
    t.code = reinterpret_cast<const cb_instr *>(base + off[CB_SEC_CODE]);
    t.consts = reinterpret_cast<const cb_const *>(base + off[CB_SEC_CONSTS]);
    t.theap = reinterpret_cast<const uint64_t *>(base + off[CB_SEC_THEAP]);
    t.str_off = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_STR_OFF]);
    t.str_bytes = base + off[CB_SEC_STR_BYTES];
    t.par_off = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_ROLE_PARENTS_OFF]);
    t.par_list = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_ROLE_PARENTS]);
    t.rp_off = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_ROLEPOL_OFF]);
    t.rp_entries = reinterpret_cast<const cb_rolepol_entry *>(base + off[CB_SEC_ROLEPOL_ENTRIES]);
    t.rp_rules = reinterpret_cast<const cb_rolepol_rule *>(base + off[CB_SEC_ROLEPOL_RULES]);
    t.rp_apats = reinterpret_cast<const uint32_t *>(base + off[CB_SEC_ROLEPOL_APATS]);
    t.nV = meta[CB_META_N_VERSIONS]; t.nRP = meta[CB_META_N_RESPATS]; t.nS = meta[CB_META_N_SCOPES]; t.nP = meta[CB_META_N_PRINCIPALS];
    t.nR = meta[CB_META_N_ROLES]; t.nAP = meta[CB_META_N_APATS]; t.nT = meta[CB_META_N_STRINGS]; t.n_slots = meta[CB_META_N_SLOTS];
    t.has_role_policies = meta[CB_META_HAS_ROLE_POLICIES]; t.has_parent_roles = meta[CB_META_HAS_PARENT_ROLES];
    t.has_principal_policies = meta[CB_META_HAS_PRINCIPAL_POLICIES];

    cb::BatchView b;
    b.hdr0 = static_cast<const cb_hdr0 *>(cols[0]); b.hdr1 = static_cast<const cb_hdr1 *>(cols[1]);
    b.roles = static_cast<const uint32_t *>(cols[2]); b.slots = static_cast<const uint64_t *>(cols[3]);
    b.heap = static_cast<const uint64_t *>(cols[4]); b.bstr_off = static_cast<const uint32_t *>(cols[5]);
    b.bstr_bytes = static_cast<const uint8_t *>(cols[6]); b.class_off = static_cast<const uint32_t *>(cols[7]);
    b.class_pats = static_cast<const uint32_t *>(cols[8]); b.aset_k = static_cast<const uint32_t *>(cols[9]);
    b.aset_spread = static_cast<const uint64_t *>(cols[10]);
    b.stride = n; b.first = 0; b.count = n;
    b.role_cols = (uint32_t)(col_bytes[2] / (4 * n)); b.n_asets = (uint32_t)(col_bytes[9] / 4);
    uint32_t km = max_actions ? max_actions : 1;
    b.kc = 64 / b.role_cols; if (b.kc > km) b.kc = km;
    b.n_pass = (km + b.kc - 1) / b.kc; b.max_actions = km; b.kbytes = (km + 7) / 8; b.flags = flags; b.now = now;
    uint32_t status = 0;
    for (uint64_t i = 0; i < n; i++) cb::eval_request(t, b, i, bitmap, &status);
    return status ? -2 : 0;
}
