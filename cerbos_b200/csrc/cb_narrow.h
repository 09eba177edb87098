// cb_narrow.h -- native narrowing of an encoded batch into the wire form cgpu_check_narrow takes (include/cerbos_b200.h).
//
// What cerbos_b200/narrow.py specifies in numpy -- per-request columns re-expressed in their narrowest EXACT class, header
// fields that are constant over the batch elided, 16-bit principal offsets, a 16- or 32-bit heap -- done in one pass per
// column on the host, so that a Go / C host goes protobuf -> cgpu_encode -> cgpu_narrow_build -> cgpu_check_narrow without
// Python.  Decision for decision the Python module (tests/test_native_encoder.py compares every array and parameter byte for
// byte); nothing is approximated: a column that fits no class keeps its 8-byte form, a batch whose ids do not fit the
// 16- / 8-bit header fields is not narrowed at all.  Host-only, no CUDA dependencies.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "cerbos_b200_format.h"

namespace cbnarrow {

enum { SLOT_U64 = 0, SLOT_U32_ID = 1, SLOT_U32_HEAP = 2, SLOT_F32 = 3, SLOT_U8 = 4, SLOT_U16_ID = 5, SLOT_U8_NUM = 6 };

struct Narrowed {
    bool ok = false;                      // false: some id does not fit its narrow field -- use the canonical call
    uint64_t n = 0;
    uint32_t role_cols = 1, n_slots = 0;
    std::vector<uint8_t> pid;             // u32[n], or u16[n] offsets from principal_base
    bool pid16 = false;
    uint32_t principal_base = 0;
    std::vector<uint16_t> hdr16;          // [n][fields that vary]
    uint32_t hdr_const_mask = 0;
    uint16_t hdr_const[4] = {0, 0, 0, 0};
    std::vector<uint8_t> versions;        // [n][2], empty when constant
    bool versions_const = false;
    uint8_t versions_value[2] = {0, 0};
    std::vector<uint8_t> roles;           // [role_cols][n]
    std::vector<uint8_t> slot_class;      // [max(n_slots, 1)]
    std::vector<uint32_t> slot_base, slot_base2;
    std::vector<std::vector<uint8_t>> slot_cols;
    std::vector<uint8_t> heap;            // u16 / u32 / u64 words
    uint32_t heap_bits = 0;               // 16, or 0 with heap_u32 saying 32 / 64
    bool heap_u32 = false;
    uint32_t heap_base = 0, heap_base2 = 0;
};

inline uint32_t v64_tag(uint64_t w) { const uint32_t top = (uint32_t)(w >> 48); return (top & 0xFFF0u) == 0xFFF0u ? (top & 0xFu) : 0u; }
constexpr uint64_t kPay = (1ull << 48) - 1;

// ids (non-empty) lie in [base, base + width) or [base2, base2 + width)?
inline bool two_windows(const uint64_t *ids, size_t n, uint64_t width, uint64_t *base, uint64_t *base2) {
    uint64_t lo = ~0ull;
    for (size_t i = 0; i < n; i++) if (ids[i] < lo) lo = ids[i];
    uint64_t lo2 = ~0ull, hi2 = 0;
    bool rest = false;
    for (size_t i = 0; i < n; i++) if (ids[i] >= lo + width) { rest = true; if (ids[i] < lo2) lo2 = ids[i]; if (ids[i] > hi2) hi2 = ids[i]; }
    *base = lo;
    *base2 = rest ? lo2 : lo;
    return !rest || hi2 - lo2 < width;
}

template <typename T>
inline void put(std::vector<uint8_t> &out, size_t i, T v) { memcpy(out.data() + i * sizeof(T), &v, sizeof(T)); }

// one u64 slot column -> class + narrow column (narrow.py: _narrow_slot, same order of attempts)
inline void narrow_slot(const uint64_t *col, uint64_t n, bool v2, uint8_t *cls, std::vector<uint8_t> *out, uint32_t *base, uint32_t *base2) {
    *base = *base2 = 0;
    bool all_boolspec = true, all_strboolspec = true, all_heapspec = true, all_numspec = true, any_str = false;
    uint64_t max_str = 0, max_off = 0;
    bool any_heap = false;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t w = col[i];
        const uint32_t tag = v64_tag(w);
        const bool special = tag == CB_V64_ABSENT || tag == CB_V64_ERROR || tag == CB_V64_NULL;
        const bool is_bool = tag == CB_V64_BOOL, is_str = tag == CB_V64_STRING;
        const bool is_heap = (tag == CB_V64_LIST || tag == CB_V64_MAP) && (w & CB_V64_HEAP_BATCH_BIT);
        all_boolspec &= is_bool || special;
        all_strboolspec &= is_str || is_bool || special;
        all_heapspec &= is_heap || special;
        all_numspec &= tag == 0 || special;
        if (is_str) { any_str = true; if ((w & kPay) > max_str) max_str = w & kPay; }
        if (is_heap) { any_heap = true; const uint64_t off = (w & kPay) & (CB_V64_HEAP_BATCH_BIT - 1); if (off > max_off) max_off = off; }
    }
    auto special_code = [](uint32_t tag, uint64_t absent, uint64_t error, uint64_t null_) { return tag == CB_V64_ABSENT ? absent : tag == CB_V64_ERROR ? error : null_; };
    if (all_boolspec) {
        *cls = SLOT_U8; out->resize(n);
        for (uint64_t i = 0; i < n; i++) { const uint32_t tag = v64_tag(col[i]); (*out)[i] = tag == CB_V64_BOOL ? (uint8_t)(col[i] & kPay) : (uint8_t)special_code(tag, 3, 4, 2); }
        return;
    }
    if (v2 && all_strboolspec && any_str) {
        std::vector<uint64_t> ids;
        ids.reserve(n);
        for (uint64_t i = 0; i < n; i++) if (v64_tag(col[i]) == CB_V64_STRING) ids.push_back(col[i] & kPay);
        uint64_t lo, lo2;
        if (two_windows(ids.data(), ids.size(), 0x7FF0, &lo, &lo2)) {
            *cls = SLOT_U16_ID; *base = (uint32_t)lo; *base2 = (uint32_t)lo2; out->resize(n * 2);
            for (uint64_t i = 0; i < n; i++) {
                const uint32_t tag = v64_tag(col[i]);
                const uint64_t pay = col[i] & kPay;
                uint16_t v;
                if (tag == CB_V64_STRING) v = pay >= lo + 0x7FF0 ? (uint16_t)(pay - lo2 + 0x8000) : (uint16_t)(pay - lo);
                else if (tag == CB_V64_BOOL) v = pay ? 0xFFFB : 0xFFFC;
                else v = (uint16_t)special_code(tag, 0xFFFF, 0xFFFE, 0xFFFD);
                put<uint16_t>(*out, i, v);
            }
            return;
        }
    }
    if (all_strboolspec && (!any_str || max_str < 0xFFFFFFF0ull)) {
        *cls = SLOT_U32_ID; out->resize(n * 4);
        for (uint64_t i = 0; i < n; i++) {
            const uint32_t tag = v64_tag(col[i]);
            const uint64_t pay = col[i] & kPay;
            uint32_t v;
            if (tag == CB_V64_STRING) v = (uint32_t)pay;
            else if (tag == CB_V64_BOOL) v = pay ? 0xFFFFFFFBu : 0xFFFFFFFCu;
            else v = (uint32_t)special_code(tag, 0xFFFFFFFFu, 0xFFFFFFFEu, 0xFFFFFFFDu);
            put<uint32_t>(*out, i, v);
        }
        return;
    }
    if (all_heapspec && (!any_heap || max_off < 0x7FFFFFF0ull)) {
        *cls = SLOT_U32_HEAP; out->resize(n * 4);
        for (uint64_t i = 0; i < n; i++) {
            const uint32_t tag = v64_tag(col[i]);
            uint32_t v;
            if (tag == CB_V64_LIST || tag == CB_V64_MAP) v = (uint32_t)((col[i] & kPay) & (CB_V64_HEAP_BATCH_BIT - 1)) | (tag == CB_V64_MAP ? 0x80000000u : 0u);
            else v = (uint32_t)special_code(tag, 0xFFFFFFFFu, 0xFFFFFFFEu, 0xFFFFFFFDu);
            put<uint32_t>(*out, i, v);
        }
        return;
    }
    if (v2 && all_numspec) {
        bool fits = true;
        for (uint64_t i = 0; i < n && fits; i++) {
            if (v64_tag(col[i]) != 0) continue;
            double d; memcpy(&d, &col[i], 8);
            const uint8_t small = (d >= 0 && d <= 239) ? (uint8_t)d : 0;
            const double back = (double)small;
            uint64_t bb; memcpy(&bb, &back, 8);
            fits = bb == col[i];                  // bit for bit: no -0.0, no fraction, no NaN, nothing above 239
        }
        if (fits) {
            *cls = SLOT_U8_NUM; out->resize(n);
            for (uint64_t i = 0; i < n; i++) {
                const uint32_t tag = v64_tag(col[i]);
                if (tag == 0) { double d; memcpy(&d, &col[i], 8); (*out)[i] = (uint8_t)d; }
                else (*out)[i] = (uint8_t)special_code(tag, 0xFF, 0xFE, 0xFD);
            }
            return;
        }
    }
    if (all_numspec) {
        bool exact = true;
        for (uint64_t i = 0; i < n && exact; i++) {
            if (v64_tag(col[i]) != 0) continue;
            if (col[i] == CB_V64_CANON_NAN) continue;
            double d; memcpy(&d, &col[i], 8);
            const float f = (float)d;
            const double back = (double)f;
            uint64_t bb; memcpy(&bb, &back, 8);
            exact = bb == col[i];
        }
        if (exact) {
            *cls = SLOT_F32; out->resize(n * 4);
            for (uint64_t i = 0; i < n; i++) {
                const uint32_t tag = v64_tag(col[i]);
                uint32_t v;
                if (tag != 0) v = (uint32_t)special_code(tag, 0x7FC00001u, 0x7FC00002u, 0x7FC00003u);
                else if (col[i] == CB_V64_CANON_NAN) v = 0x7FC00000u;
                else { double d; memcpy(&d, &col[i], 8); const float f = (float)d; memcpy(&v, &f, 4); }
                put<uint32_t>(*out, i, v);
            }
            return;
        }
    }
    *cls = SLOT_U64; out->resize(n * 8);
    memcpy(out->data(), col, n * 8);
}

// hdr0: u32[n][4], hdr1: {u16 rv, u16 pv, u32 aset}[n], roles: u32[role_cols][n], slots: u64[n_slots][n], heap: u64[heap_words]
inline Narrowed build(const uint32_t *hdr0, const uint8_t *hdr1, const uint32_t *roles, const uint64_t *slots, const uint64_t *heap, uint64_t heap_words,
                      uint64_t n, uint32_t role_cols, uint32_t n_slots, bool v2 = true) {
    Narrowed r;
    r.n = n; r.role_cols = role_cols; r.n_slots = n_slots;
    std::vector<uint16_t> h16(n * 4);
    std::vector<uint8_t> ver(n * 2);
    auto scope16 = [](uint32_t s, uint16_t *out) {
        if (s == CB_SCOPE_NONE) { *out = 0xFFFF; return true; }
        const uint32_t sid = s & ~(uint32_t)CB_SCOPE_INEXACT_BIT;
        if (sid >= 0x7FFF) return false;
        *out = (uint16_t)(sid | ((s & CB_SCOPE_INEXACT_BIT) ? 0x8000u : 0u));
        return true;
    };
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t kc = hdr0[i * 4 + 1];
        if (kc == CB_KIND_NONE) h16[i * 4] = 0xFFFF;
        else {
            const uint32_t kid = kc & ~(uint32_t)CB_KIND_CLASS_CSR_BIT;
            if (kid >= 0x7FFF) return r;
            h16[i * 4] = (uint16_t)(kid | ((kc & CB_KIND_CLASS_CSR_BIT) ? 0x8000u : 0u));
        }
        if (!scope16(hdr0[i * 4 + 2], &h16[i * 4 + 1]) || !scope16(hdr0[i * 4 + 3], &h16[i * 4 + 2])) return r;
        uint16_t rv, pv; uint32_t aset;
        memcpy(&rv, hdr1 + i * 8, 2); memcpy(&pv, hdr1 + i * 8 + 2, 2); memcpy(&aset, hdr1 + i * 8 + 4, 4);
        if (aset > 0xFFFF) return r;
        h16[i * 4 + 3] = (uint16_t)aset;
        if ((rv != CB_NONE16 && rv >= 0xFF) || (pv != CB_NONE16 && pv >= 0xFF)) return r;
        ver[i * 2] = rv == CB_NONE16 ? 0xFF : (uint8_t)rv;
        ver[i * 2 + 1] = pv == CB_NONE16 ? 0xFF : (uint8_t)pv;
    }
    r.roles.resize((size_t)role_cols * n);
    for (uint64_t j = 0; j < (uint64_t)role_cols * n; j++) {
        const uint32_t x = roles[j];
        if (x == CB_ROLE_PAD) r.roles[j] = 0xFF;
        else if (x == CB_ROLE_UNKNOWN) r.roles[j] = 0xFE;
        else { if (x >= 0xFE) return r; r.roles[j] = (uint8_t)x; }
    }
    const uint32_t ns = n_slots ? n_slots : 1;
    r.slot_class.assign(ns, 0); r.slot_base.assign(ns, 0); r.slot_base2.assign(ns, 0);
    r.slot_cols.resize(n_slots);
    for (uint32_t v = 0; v < n_slots; v++) narrow_slot(slots + (uint64_t)v * n, n, v2, &r.slot_class[v], &r.slot_cols[v], &r.slot_base[v], &r.slot_base2[v]);
    // heap
    bool u32_ok = true, u16_words_ok = heap_words != 0;
    std::vector<uint64_t> hs;
    for (uint64_t i = 0; i < heap_words; i++) {
        const uint64_t w = heap[i];
        const bool is_str = v64_tag(w) == CB_V64_STRING;
        u32_ok &= w < (1ull << 31) || (is_str && (w & kPay) < (1ull << 31));
        u16_words_ok &= is_str || w < (1ull << 15);
        if (is_str && v2) hs.push_back(w & kPay);
    }
    uint64_t hb = 0, hb2 = 0;
    if (v2 && u16_words_ok && (hs.empty() || two_windows(hs.data(), hs.size(), 1ull << 14, &hb, &hb2))) {
        r.heap_bits = 16; r.heap_base = (uint32_t)hb; r.heap_base2 = (uint32_t)hb2;
        r.heap.resize(heap_words * 2);
        for (uint64_t i = 0; i < heap_words; i++) {
            const uint64_t w = heap[i];
            uint16_t x;
            if (v64_tag(w) == CB_V64_STRING) { const uint64_t pay = w & kPay; x = pay >= hb + (1ull << 14) ? (uint16_t)((pay - hb2) | 0xC000) : (uint16_t)((pay - hb) | 0x8000); }
            else x = (uint16_t)w;
            put<uint16_t>(r.heap, i, x);
        }
    } else if (u32_ok) {
        r.heap_u32 = true;
        r.heap.resize(heap_words * 4);
        for (uint64_t i = 0; i < heap_words; i++) {
            const uint64_t w = heap[i];
            put<uint32_t>(r.heap, i, v64_tag(w) == CB_V64_STRING ? (uint32_t)((w & kPay) | (1ull << 31)) : (uint32_t)w);
        }
    } else {
        r.heap.resize(heap_words * 8);
        if (heap_words) memcpy(r.heap.data(), heap, heap_words * 8);
    }
    // principal ids, header fields, versions
    uint32_t pmin = 0xFFFFFFFFu, pmax = 0;
    for (uint64_t i = 0; i < n; i++) { const uint32_t p = hdr0[i * 4]; if (p < pmin) pmin = p; if (p > pmax) pmax = p; }
    if (v2 && n && pmax - pmin <= 0xFFFF) {
        r.pid16 = true; r.principal_base = pmin; r.pid.resize(n * 2);
        for (uint64_t i = 0; i < n; i++) put<uint16_t>(r.pid, i, (uint16_t)(hdr0[i * 4] - pmin));
    } else {
        r.pid.resize(n * 4);
        for (uint64_t i = 0; i < n; i++) put<uint32_t>(r.pid, i, hdr0[i * 4]);
    }
    if (v2 && n) {
        uint32_t keep[4], nk = 0;
        for (uint32_t f = 0; f < 4; f++) {
            bool same = true;
            for (uint64_t i = 1; i < n && same; i++) same = h16[i * 4 + f] == h16[f];
            if (same) { r.hdr_const_mask |= 1u << f; r.hdr_const[f] = h16[f]; } else keep[nk++] = f;
        }
        r.hdr16.resize(n * nk);
        for (uint64_t i = 0; i < n; i++) for (uint32_t q = 0; q < nk; q++) r.hdr16[i * nk + q] = h16[i * 4 + keep[q]];
        bool vsame = true;
        for (uint64_t i = 1; i < n && vsame; i++) vsame = ver[i * 2] == ver[0] && ver[i * 2 + 1] == ver[1];
        if (vsame) { r.versions_const = true; r.versions_value[0] = ver[0]; r.versions_value[1] = ver[1]; } else r.versions = ver;
    } else {
        r.hdr16 = h16;
        r.versions = ver;
    }
    r.ok = true;
    return r;
}

}  // namespace cbnarrow
