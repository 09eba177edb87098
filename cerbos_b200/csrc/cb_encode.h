// cb_encode.h -- native batch encoder: serialized enginev1.CheckInput messages -> SoA request columns.
//
// Host half of the replacement for the per-input string / map work of RuleTable.check (internal/ruletable/ruletable.go:
// 785-884: default version / scope :789-799 + evaluator.go:99-113, namer.SanitizedResource :851, GetAllScopes :611-645)
// and of the glob lookups over actions / resource kinds (internal/util/globs_common.go, glob_map.go:138-186): every
// distinct string becomes a dictionary id / pattern class once per batch; the device sees integers only.  What the Go
// side hands over is what it already holds when svc.CheckResources assembles its inputs (internal/svc/cerbos_svc.go:
// 249-265): `proto.Marshal` of each enginev1.CheckInput (api/public/cerbos/engine/v1/engine.proto `CheckInput`,
// `Principal`, `Resource`, `AuxData`; attributes are google.protobuf.Value trees).  A hand-written protobuf wire reader:
// no generated code, no protoc.
//
// Byte-for-byte the columns cerbos_b200/encode.py builds from the same inputs (tests/test_native_encoder.py), so the
// Python host and a Go host drive the kernels identically.  Host-only, no CUDA dependencies (tests compile it for the CPU).
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "cerbos_b200_format.h"

namespace cbenc {

// ---------------------------------------------------------------------------------------------- tiny JSON (MANIFEST)
struct Json {
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json *get(const char *key) const {
        for (const auto &kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};
struct JsonParser {
    const char *p, *e;
    bool ok = true;
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    static void utf8(std::string &out, uint32_t cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | cp >> 6); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | cp >> 12); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | cp >> 18); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    uint32_t hex4() {
        uint32_t v = 0;
        for (int i = 0; i < 4 && p < e; i++, p++) {
            const char c = *p;
            v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 0);
        }
        return v;
    }
    std::string string() {
        std::string out;
        if (p >= e || *p != '"') { ok = false; return out; }
        p++;
        while (p < e && *p != '"') {
            if (*p == '\\' && p + 1 < e) {
                p++;
                switch (*p++) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; case 'b': out += '\b'; break;
                case 'f': out += '\f'; break; case '/': out += '/'; break; case '\\': out += '\\'; break; case '"': out += '"'; break;
                case 'u': {
                    uint32_t cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && p + 1 < e && p[0] == '\\' && p[1] == 'u') { p += 2; const uint32_t lo = hex4(); cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); }
                    utf8(out, cp);
                    break;
                }
                default: ok = false;
                }
            } else out += *p++;
        }
        if (p < e) p++; else ok = false;
        return out;
    }
    Json value() {
        Json j;
        ws();
        if (p >= e) { ok = false; return j; }
        if (*p == '"') { j.kind = Json::STR; j.str = string(); }
        else if (*p == '[') {
            j.kind = Json::ARR; p++; ws();
            if (p < e && *p == ']') { p++; return j; }
            while (ok) { j.arr.push_back(value()); ws(); if (p < e && *p == ',') { p++; continue; } if (p < e && *p == ']') { p++; break; } ok = false; }
        } else if (*p == '{') {
            j.kind = Json::OBJ; p++; ws();
            if (p < e && *p == '}') { p++; return j; }
            while (ok) {
                ws();
                std::string k = string();
                ws();
                if (p >= e || *p != ':') { ok = false; break; }
                p++;
                j.obj.emplace_back(std::move(k), value());
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == '}') { p++; break; }
                ok = false;
            }
        } else if (!strncmp(p, "true", 4) && e - p >= 4) { j.kind = Json::BOOL; j.b = true; p += 4; }
        else if (!strncmp(p, "false", 5) && e - p >= 5) { j.kind = Json::BOOL; p += 5; }
        else if (!strncmp(p, "null", 4) && e - p >= 4) { p += 4; }
        else { char *end = nullptr; j.kind = Json::NUM; j.num = strtod(p, &end); if (end == p) ok = false; p = end; }
        return j;
    }
};

// ---------------------------------------------------------------------------------------------- naming rules (namer.go)
inline bool name_char(char c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_' || c == '@' || c == '.' || c == '-' || c == '/'; }
inline bool alpha(char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }
// namer.go:213-218 with the patterns at :17-20: names of the old form have every run of characters outside [0-9A-Za-z_.] replaced by "_"
inline std::string sanitize(const std::string &v) {
    size_t i = 0;
    const size_t n = v.size();
    bool old_form = n > 0;
    while (old_form && i < n) {          // segment (":" segment)*, segment = alpha name_char*
        if (!alpha(v[i])) { old_form = false; break; }
        i++;
        while (i < n && name_char(v[i])) i++;
        if (i < n) { if (v[i] != ':' || i + 1 >= n) { old_form = false; break; } i++; }
    }
    if (!old_form) return v;
    std::string out;
    bool in_run = false;
    for (char c : v) {
        const bool keep = (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_' || c == '.';
        if (keep) { out += c; in_run = false; }
        else if (!in_run) { out += '_'; in_run = true; }
    }
    return out;
}
inline std::string scope_value(const std::string &s) { return !s.empty() && s[0] == '.' ? s.substr(1) : s; }   // namer.go:276-278

// ---------------------------------------------------------------------------------------------- globs (util/globs_common.go)
// gobwas/glob v0.2.3 syntax with ':' as the separator: `*` any run of non-separator characters, `**` any run, `?` one
// non-separator character, [abc] / [a-z] / [!abc] classes, {a,b} alternatives, backslash escape; a lone "*" means "**"
// (globs_common.go:74-81).  Invalid patterns match nothing.
struct Glob {
    static bool match(const std::string &pat, const std::string &val) {
        const std::string p = pat == "*" ? std::string("**") : pat;
        bool bad = false;
        const bool r = rec(p, 0, p.size(), val, 0, bad);
        return r && !bad;
    }
    // matches p[pi, pe) against s[si, end)
    static bool rec(const std::string &p, size_t pi, size_t pe, const std::string &s, size_t si, bool &bad) {
        while (pi < pe) {
            const char c = p[pi];
            if (c == '\\') {
                if (pi + 1 >= pe) { bad = true; return false; }
                if (si >= s.size() || s[si] != p[pi + 1]) return false;
                pi += 2; si++;
            } else if (c == '*') {
                size_t j = pi;
                while (j < pe && p[j] == '*') j++;
                const bool any = j - pi >= 2;
                for (size_t k = si;; k++) {
                    if (rec(p, j, pe, s, k, bad)) return true;
                    if (bad || k >= s.size() || (!any && s[k] == ':')) return false;
                }
            } else if (c == '?') {
                if (si >= s.size() || s[si] == ':') return false;
                pi++; si++;
            } else if (c == '[') {
                size_t j = pi + 1;
                const bool neg = j < pe && p[j] == '!';
                if (neg) j++;
                if (si >= s.size()) { size_t k = j; while (k < pe && p[k] != ']') k += (p[k] == '\\' && k + 1 < pe) ? 2 : 1; if (k >= pe) bad = true; return false; }
                bool hit = false;
                size_t k = j;
                bool have_prev = false;
                char prev = 0;
                while (k < pe && p[k] != ']') {
                    char lo;
                    if (p[k] == '\\' && k + 1 < pe) { lo = p[k + 1]; k += 2; }
                    else if (p[k] == '-' && have_prev && k + 1 < pe && p[k + 1] != ']') {
                        char hi = p[k + 1];
                        size_t adv = 2;
                        if (hi == '\\' && k + 2 < pe) { hi = p[k + 2]; adv = 3; }
                        if ((unsigned char)s[si] >= (unsigned char)prev && (unsigned char)s[si] <= (unsigned char)hi) hit = true;
                        k += adv;
                        have_prev = false;
                        continue;
                    } else { lo = p[k]; k++; }
                    if (s[si] == lo) hit = true;
                    prev = lo; have_prev = true;
                }
                if (k >= pe) { bad = true; return false; }
                if (hit == neg) return false;
                pi = k + 1; si++;
            } else if (c == '{') {
                // alternatives up to the matching '}' (nested braces allowed), each followed by the rest of the pattern
                size_t depth = 1, k = pi + 1, start = pi + 1;
                std::vector<std::pair<size_t, size_t>> alts;
                while (k < pe && depth) {
                    if (p[k] == '\\' && k + 1 < pe) { k += 2; continue; }
                    if (p[k] == '{') depth++;
                    else if (p[k] == '}') { depth--; if (!depth) break; }
                    else if (p[k] == ',' && depth == 1) { alts.emplace_back(start, k); start = k + 1; }
                    k++;
                }
                if (k >= pe) { bad = true; return false; }
                alts.emplace_back(start, k);
                const std::string rest = p.substr(k + 1, pe - (k + 1));
                for (const auto &a : alts) {
                    const std::string sub = p.substr(a.first, a.second - a.first) + rest;
                    if (rec(sub, 0, sub.size(), s, si, bad)) return true;
                    if (bad) return false;
                }
                return false;
            } else {
                if (si >= s.size() || s[si] != c) return false;
                pi++; si++;
            }
        }
        return si == s.size();
    }
};
// GlobMap semantics (internal/ruletable/internal/glob_map.go:60-75): literal equality, or a glob match when the key contains '*'
inline bool key_matches(const std::string &key, const std::string &val) {
    if (key == val) return true;
    return key.find('*') != std::string::npos && Glob::match(key, val);
}

// ---------------------------------------------------------------------------------------------- protobuf wire reader
struct Span { const uint8_t *p = nullptr; size_t n = 0; };
struct WireIt {
    const uint8_t *p, *e;
    bool bad = false;
    uint32_t fno = 0, wt = 0;
    uint64_t u = 0;
    Span s;
    explicit WireIt(Span sp) : p(sp.p), e(sp.p + sp.n) {}
    bool varint(uint64_t *out) {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p >= e) return false;
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7F) << shift;
            if (!(b & 0x80)) { *out = v; return true; }
        }
        return false;
    }
    bool next() {
        if (p >= e) return false;
        uint64_t key;
        if (!varint(&key)) { bad = true; return false; }
        fno = (uint32_t)(key >> 3); wt = (uint32_t)(key & 7);
        if (wt == 0) { if (!varint(&u)) { bad = true; return false; } }
        else if (wt == 1) { if (e - p < 8) { bad = true; return false; } memcpy(&u, p, 8); p += 8; }
        else if (wt == 5) { if (e - p < 4) { bad = true; return false; } uint32_t w; memcpy(&w, p, 4); u = w; p += 4; }
        else if (wt == 2) {
            uint64_t ln;
            if (!varint(&ln) || (uint64_t)(e - p) < ln) { bad = true; return false; }
            s.p = p; s.n = (size_t)ln; p += ln;
        } else { bad = true; return false; }
        return true;
    }
};
inline std::string str_of(Span s) { return std::string(reinterpret_cast<const char *>(s.p), s.n); }
// one entry of a map<string, X> field: key bytes + value bytes
inline bool map_entry(Span ent, Span *key, Span *val) {
    WireIt it(ent);
    *key = Span(); *val = Span();
    while (it.next()) { if (it.fno == 1 && it.wt == 2) *key = it.s; else if (it.fno == 2 && it.wt == 2) *val = it.s; }
    return !it.bad;
}

// ---------------------------------------------------------------------------------------------- byte-keyed dictionary
// string -> u32 keyed by raw bytes (no std::string temporaries on the lookup path): open addressing, keys copied into one arena
inline uint64_t hash_bytes(const uint8_t *p, size_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xD6E8FEB86659FD93ull);
    while (n >= 8) { uint64_t w; memcpy(&w, p, 8); h = (h ^ w) * 0x9FB21C651E98DF25ull; h ^= h >> 29; p += 8; n -= 8; }
    uint64_t w = 0;
    if (n) { memcpy(&w, p, n); h = (h ^ w) * 0x9FB21C651E98DF25ull; }
    h ^= h >> 32;
    return h * 0xD6E8FEB86659FD93ull;
}
struct BytesMap {
    struct Ent { uint64_t h; uint32_t off, len, val; };
    std::vector<uint32_t> tab;    // entry index + 1, 0 = empty
    std::vector<Ent> ents;
    std::vector<uint8_t> arena;
    BytesMap() { tab.assign(64, 0); }
    size_t size() const { return ents.size(); }
    const uint8_t *key(size_t i) const { return arena.data() + ents[i].off; }
    const uint32_t *find(const uint8_t *p, size_t n, uint64_t h) const {
        const size_t mask = tab.size() - 1;
        for (size_t i = (size_t)(h >> 7) & mask;; i = (i + 1) & mask) {
            const uint32_t e = tab[i];
            if (!e) return nullptr;
            const Ent &x = ents[e - 1];
            if (x.h == h && x.len == n && (n == 0 || !memcmp(arena.data() + x.off, p, n))) return &x.val;   // (an empty key may come with a null pointer)
        }
    }
    const uint32_t *find(const uint8_t *p, size_t n) const { return find(p, n, hash_bytes(p, n)); }
    void insert(const uint8_t *p, size_t n, uint64_t h, uint32_t val) {   // the key must not be present
        if ((ents.size() + 1) * 2 > tab.size()) {
            std::vector<uint32_t> nt(tab.size() * 2, 0);
            const size_t mask = nt.size() - 1;
            for (size_t e = 0; e < ents.size(); e++) {
                size_t i = (size_t)(ents[e].h >> 7) & mask;
                while (nt[i]) i = (i + 1) & mask;
                nt[i] = (uint32_t)e + 1;
            }
            tab.swap(nt);
        }
        ents.push_back(Ent{h, (uint32_t)arena.size(), (uint32_t)n, val});
        arena.insert(arena.end(), p, p + n);
        const size_t mask = tab.size() - 1;
        size_t i = (size_t)(h >> 7) & mask;
        while (tab[i]) i = (i + 1) & mask;
        tab[i] = (uint32_t)ents.size();
    }
};
inline bool span_eq(Span a, Span b) { return a.n == b.n && (a.n == 0 || !memcmp(a.p, b.p, a.n)); }

// ---------------------------------------------------------------------------------------------- encoder
constexpr uint64_t box(uint32_t tag, uint64_t payload = 0) { return ((uint64_t)(CB_V64_BOX_BASE | tag) << 48) | (payload & 0xFFFFFFFFFFFFull); }
constexpr uint64_t V_ABSENT = box(CB_V64_ABSENT), V_ERROR = box(CB_V64_ERROR), V_NULL = box(CB_V64_NULL);

struct Conf {
    std::string default_version = "default", default_scope;
    bool lenient = false;
};

struct Columns {   // the twelve cgpu_batch columns, in order; buffers owned here
    uint64_t n = 0;
    uint32_t max_actions = 1, role_cols = 1, kc = 1, n_pass = 1;
    std::vector<uint32_t> hdr0;        // [n][4]
    std::vector<uint8_t> hdr1;         // [n] x {u16 rv, u16 pv, u32 aset}
    std::vector<uint32_t> roles;       // [role_cols][n]
    std::vector<uint64_t> slots;       // [max(n_slots, 1)][n]
    std::vector<uint64_t> heap;
    std::vector<uint32_t> bstr_off;
    std::vector<uint8_t> bstr_bytes;
    std::vector<uint32_t> class_off, class_pats, aset_k;
    std::vector<uint64_t> aset_spread, row_am;
    const void *ptr(int i) const {
        switch (i) {
        case 0: return hdr0.data(); case 1: return hdr1.data(); case 2: return roles.data(); case 3: return slots.data(); case 4: return heap.data();
        case 5: return bstr_off.data(); case 6: return bstr_bytes.data(); case 7: return class_off.data(); case 8: return class_pats.data();
        case 9: return aset_k.data(); case 10: return aset_spread.data(); default: return row_am.data();
        }
    }
    size_t bytes(int i) const {
        switch (i) {
        case 0: return hdr0.size() * 4; case 1: return hdr1.size(); case 2: return roles.size() * 4; case 3: return slots.size() * 8; case 4: return heap.size() * 8;
        case 5: return bstr_off.size() * 4; case 6: return bstr_bytes.size(); case 7: return class_off.size() * 4; case 8: return class_pats.size() * 4;
        case 9: return aset_k.size() * 4; case 10: return aset_spread.size() * 8; default: return row_am.size() * 8;
        }
    }
};

struct PrincipalView { Span id, version, scope; std::vector<Span> roles; std::vector<std::pair<Span, Span>> attr; };
struct ResourceView { Span kind, version, id, scope; std::vector<std::pair<Span, Span>> attr; };
struct InputView {
    PrincipalView p; ResourceView r; std::vector<Span> actions; std::vector<std::pair<Span, Span>> jwt; bool has_aux = false;
    void clear() {   // keeps the vectors' capacity: one view per thread is reused for every message of its shard
        p.id = p.version = p.scope = Span{}; p.roles.clear(); p.attr.clear();
        r.kind = r.version = r.id = r.scope = Span{}; r.attr.clear();
        actions.clear(); jwt.clear(); has_aux = false;
    }
};

class Encoder {
  public:
    Conf conf;
    std::string error;
    std::vector<std::string> versions, scopes, respats, roles, apats, strings;
    std::vector<std::vector<std::string>> slots;
    std::vector<uint32_t> row_pat_start, row_apats;
    std::unordered_map<std::string, uint32_t> version_ids, scope_ids, role_ids;
    BytesMap table_strings, role_map;
    // where a slot's value comes from, decided once: [0] of the path and, for attribute paths, where the walk starts
    enum SlotSrc { SRC_ERROR, SRC_AUX, SRC_P_ATTR, SRC_R_ATTR, SRC_P_ROLES, SRC_R_ROLES, SRC_P_SCOPE, SRC_R_SCOPE, SRC_P_VERSION, SRC_R_VERSION, SRC_P_ID, SRC_R_ID,
                   SRC_P_KIND, SRC_R_KIND, SRC_EMPTY };
    std::vector<SlotSrc> slot_src;

    // blob: the table blob (its MANIFEST section carries the dictionaries)
    bool init(const void *blob, size_t len, const Conf &c) {
        conf = c;
        const cb_blob_header *h = static_cast<const cb_blob_header *>(blob);
        if (len < sizeof(*h) || h->magic != CB_MAGIC || h->version != CB_VERSION) { error = "not a cerbos_b200 table blob (bad magic / version)"; return false; }
        const cb_section_desc *sd = reinterpret_cast<const cb_section_desc *>(static_cast<const char *>(blob) + sizeof(cb_blob_header));
        const char *man = nullptr;
        size_t man_len = 0;
        for (uint32_t i = 0; i < h->n_sections; i++)
            if (sd[i].id == CB_SEC_MANIFEST && sd[i].offset <= len && sd[i].n_bytes <= len - sd[i].offset) { man = static_cast<const char *>(blob) + sd[i].offset; man_len = sd[i].n_bytes; }
        if (!man) { error = "table blob has no MANIFEST section"; return false; }
        JsonParser jp{man, man + man_len};
        const Json root = jp.value();
        if (!jp.ok || root.kind != Json::OBJ) { error = "MANIFEST is not valid JSON"; return false; }
        auto strs = [&](const char *key, std::vector<std::string> *out) {
            const Json *a = root.get(key);
            if (a) for (const Json &x : a->arr) out->push_back(x.str);
        };
        strs("versions", &versions); strs("scopes", &scopes); strs("respats", &respats); strs("roles", &roles); strs("apats", &apats); strs("strings", &strings);
        if (const Json *a = root.get("slots")) for (const Json &path : a->arr) { slots.emplace_back(); for (const Json &seg : path.arr) slots.back().push_back(seg.str); }
        if (const Json *a = root.get("row_pat_start")) for (const Json &x : a->arr) row_pat_start.push_back((uint32_t)x.num);
        if (const Json *a = root.get("row_apats")) for (const Json &x : a->arr) row_apats.push_back((uint32_t)x.num);
        for (uint32_t i = 0; i < versions.size(); i++) version_ids[versions[i]] = i;
        for (uint32_t i = 0; i < scopes.size(); i++) scope_ids[scopes[i]] = i;
        for (uint32_t i = 0; i < roles.size(); i++) role_ids[roles[i]] = i;
        for (uint32_t i = 0; i < strings.size(); i++) {
            const uint8_t *sp = reinterpret_cast<const uint8_t *>(strings[i].data());
            const uint64_t hh = hash_bytes(sp, strings[i].size());
            if (!table_strings.find(sp, strings[i].size(), hh)) table_strings.insert(sp, strings[i].size(), hh, i);
        }
        for (uint32_t i = 0; i < roles.size(); i++) {
            const uint8_t *sp = reinterpret_cast<const uint8_t *>(roles[i].data());
            const uint64_t hh = hash_bytes(sp, roles[i].size());
            if (const uint32_t *old = role_map.find(sp, roles[i].size(), hh)) *const_cast<uint32_t *>(old) = i; else role_map.insert(sp, roles[i].size(), hh, i);
        }
        for (const auto &path : slots) {
            SlotSrc src = SRC_ERROR;
            if (!path.empty()) {
                if (path[0] == "aux_data") src = SRC_AUX;
                else {
                    const bool principal = path[0] == "principal";
                    const std::string &fld = path.size() > 1 ? path[1] : path[0];
                    src = fld == "attr" ? (principal ? SRC_P_ATTR : SRC_R_ATTR) : fld == "roles" ? (principal ? SRC_P_ROLES : SRC_R_ROLES)
                          : fld == "scope" ? (principal ? SRC_P_SCOPE : SRC_R_SCOPE) : fld == "policy_version" ? (principal ? SRC_P_VERSION : SRC_R_VERSION)
                          : fld == "id" ? (principal ? SRC_P_ID : SRC_R_ID) : fld == "kind" ? (principal ? SRC_P_KIND : SRC_R_KIND) : SRC_EMPTY;
                }
            }
            slot_src.push_back(src);
        }
        return true;
    }

    // inputs: n serialized enginev1.CheckInput messages.  n_threads > 1: contiguous shards are encoded concurrently with
    // shard-local dictionaries / heaps, then merged in shard order -- the result is byte for byte what one thread produces
    // (first-appearance order of strings, classes, action sets and heap records is the sequential order).
    // Read-only on the encoder (safe from any number of threads at once); *err receives the reason when it returns false.
    bool encode(const void *const *inputs, const size_t *lens, uint64_t n, Columns *out, unsigned n_threads = 1, std::string *err = nullptr) const {
        std::string err_local;
        std::string *self_error = err ? err : &err_local;
        if (n_threads < 1) n_threads = 1;
        if ((uint64_t)n_threads > (n + 255) / 256) n_threads = (unsigned)((n + 255) / 256);
        const uint64_t per = (n + n_threads - 1) / n_threads;
        auto span_of = [&](unsigned t, uint64_t *lo, uint64_t *hi) { *lo = (uint64_t)t * per; *hi = *lo + per < n ? *lo + per : n; if (*lo > n) *lo = n; };
        std::vector<int> bad(n_threads, 0);
        std::vector<uint32_t> mr(n_threads, 1), ma(n_threads, 1);
        run(n_threads, [&](unsigned t) {
            uint64_t lo, hi;
            span_of(t, &lo, &hi);
            // first pass: only what sizes the columns (roles per principal, actions per input); the messages are parsed for
            // real in the second pass, into one reusable view per thread
            for (uint64_t i = lo; i < hi; i++) {
                uint32_t nr = 0, na = 0;
                if (!scan_counts(Span{static_cast<const uint8_t *>(inputs[i]), lens[i]}, &nr, &na)) { bad[t] = 1; return; }
                if (nr > mr[t]) mr[t] = nr;
                if (na > ma[t]) ma[t] = na;
            }
        });
        uint32_t max_roles = 1, max_actions = 1;
        for (unsigned t = 0; t < n_threads; t++) {
            if (bad[t]) { *self_error = "malformed CheckInput message"; return false; }
            if (mr[t] > max_roles) max_roles = mr[t];
            if (ma[t] > max_actions) max_actions = ma[t];
        }
        if (max_roles > CB_MAX_ROLE_COLS) { *self_error = "more than 16 roles on one principal is not supported"; return false; }
        Columns &c = *out;
        c.n = n; c.role_cols = max_roles; c.max_actions = max_actions;
        c.kc = 64 / max_roles; if (c.kc > max_actions) c.kc = max_actions; if (c.kc < 1) c.kc = 1;
        c.n_pass = (max_actions + c.kc - 1) / c.kc;
        const size_t n_slots = slots.size();
        c.hdr0.assign(n * 4, 0);
        c.hdr1.assign(n * 8, 0);
        c.roles.assign((size_t)max_roles * n, CB_ROLE_PAD);
        c.slots.assign((n_slots ? n_slots : 1) * n, 0);

        std::vector<State> shards;
        std::vector<std::vector<uint64_t>> heaps(n_threads);
        std::vector<std::string> errs(n_threads);
        shards.reserve(n_threads);
        for (unsigned t = 0; t < n_threads; t++) { shards.emplace_back(this, &c); shards.back().heap = n_threads == 1 ? &c.heap : &heaps[t]; }
        run(n_threads, [&](unsigned t) {
            uint64_t lo, hi;
            span_of(t, &lo, &hi);
            State &st = shards[t];
            // Header fields repeat from one request to the next (a CheckResources call shares its principal and actions, a
            // batch its kinds and scopes): each is resolved again only when its bytes differ from the previous request's.
            Span m_kind{}, m_rscope{}, m_pscope{}, m_rver{}, m_pver{};
            bool have = false;
            uint32_t cid = CB_KIND_NONE, rs_id = CB_SCOPE_NONE, ps_id = CB_SCOPE_NONE, aid = 0;
            uint16_t rv = (uint16_t)CB_NONE16, pv = (uint16_t)CB_NONE16;
            std::vector<Span> m_actions;
            bool have_actions = false;
            InputView v;
            for (uint64_t i = lo; i < hi; i++) {
                v.clear();
                if (!parse_input(Span{static_cast<const uint8_t *>(inputs[i]), lens[i]}, &v)) { errs[t] = "malformed CheckInput message"; return; }
                if (!have || !span_eq(v.r.kind, m_kind)) { cid = st.kind_class(str_of(v.r.kind), &errs[t]); m_kind = v.r.kind; if (!errs[t].empty()) return; }
                if (!have || !span_eq(v.r.scope, m_rscope)) { rs_id = resolve_scope(scope_value(v.r.scope.n ? str_of(v.r.scope) : conf.default_scope)); m_rscope = v.r.scope; }
                if (!have || !span_eq(v.p.scope, m_pscope)) { ps_id = resolve_scope(scope_value(v.p.scope.n ? str_of(v.p.scope) : conf.default_scope)); m_pscope = v.p.scope; }
                if (!have || !span_eq(v.r.version, m_rver)) { rv = version_id(v.r.version.n ? str_of(v.r.version) : conf.default_version); m_rver = v.r.version; }
                if (!have || !span_eq(v.p.version, m_pver)) { pv = version_id(v.p.version.n ? str_of(v.p.version) : conf.default_version); m_pver = v.p.version; }
                bool same_actions = have_actions && m_actions.size() == v.actions.size();
                for (size_t k = 0; same_actions && k < v.actions.size(); k++) same_actions = span_eq(v.actions[k], m_actions[k]);
                if (!same_actions) { aid = st.action_set(v.actions); m_actions = v.actions; have_actions = true; }
                have = true;
                c.hdr0[i * 4 + 0] = st.sid(v.p.id);
                c.hdr0[i * 4 + 1] = cid;
                c.hdr0[i * 4 + 2] = rs_id;
                c.hdr0[i * 4 + 3] = ps_id;
                memcpy(&c.hdr1[i * 8], &rv, 2); memcpy(&c.hdr1[i * 8 + 2], &pv, 2); memcpy(&c.hdr1[i * 8 + 4], &aid, 4);
                for (size_t j = 0; j < v.p.roles.size(); j++) {
                    const uint32_t *rid = role_map.find(v.p.roles[j].p, v.p.roles[j].n);
                    c.roles[j * n + i] = rid ? *rid : CB_ROLE_UNKNOWN;
                }
                for (size_t s = 0; s < n_slots; s++) c.slots[s * n + i] = st.slot_value(v, slots[s], slot_src[s]);
            }
        });
        for (unsigned t = 0; t < n_threads; t++) if (!errs[t].empty()) { *self_error = errs[t]; return false; }
        if (n_threads > 1) {
            // merge the shard-local tables into shard 0's, in order; remember how every local id maps
            State &g = shards[0];
            std::vector<std::vector<uint32_t>> str_map(n_threads), class_map(n_threads), aset_map(n_threads);
            std::vector<uint64_t> heap_base(n_threads, 0);
            c.heap = std::move(heaps[0]);
            for (unsigned t = 1; t < n_threads; t++) {
                State &st = shards[t];
                str_map[t].resize(st.bstr.size());
                for (size_t j = 0; j < st.bstr.size(); j++) {
                    const BytesMap::Ent &e = st.bstr.ents[j];       // entry j holds batch string j of the shard (values are assigned in order)
                    const uint8_t *kp = st.bstr.key(j);
                    const uint32_t *hit = g.bstr.find(kp, e.len, e.h);
                    if (hit) str_map[t][j] = *hit;
                    else { str_map[t][j] = (uint32_t)g.bstr.size(); g.bstr.insert(kp, e.len, e.h, (uint32_t)g.bstr.size()); }
                }
                class_map[t].resize(st.class_list.size());
                for (size_t j = 0; j < st.class_list.size(); j++) {
                    auto it = g.classes.find(st.class_list[j]);
                    if (it == g.classes.end()) { it = g.classes.emplace(st.class_list[j], (uint32_t)g.class_list.size()).first; g.class_list.push_back(st.class_list[j]); }
                    class_map[t][j] = it->second;
                }
                aset_map[t].resize(st.aset_list.size());
                for (size_t j = 0; j < st.aset_list.size(); j++) {
                    auto it = g.asets.find(st.aset_list[j]);
                    if (it == g.asets.end()) { it = g.asets.emplace(st.aset_list[j], (uint32_t)g.aset_list.size()).first; g.aset_list.push_back(st.aset_list[j]); }
                    aset_map[t][j] = it->second;
                }
                heap_base[t] = c.heap.size();
                c.heap.resize(c.heap.size() + heaps[t].size());
            }
            const uint64_t nT = strings.size();
            run(n_threads, [&](unsigned t) {
                if (t == 0) return;
                uint64_t lo, hi;
                span_of(t, &lo, &hi);
                const std::vector<uint32_t> &sm = str_map[t];
                const uint64_t hb = heap_base[t];
                auto fix = [&](uint64_t w) -> uint64_t {
                    const uint32_t top = (uint32_t)(w >> 48);
                    if ((top & 0xFFF0u) != 0xFFF0u) return w;
                    const uint32_t tag = top & 0xFu;
                    if (tag == CB_V64_STRING) { const uint64_t id = w & 0xFFFFFFFFFFFFull; return id >= nT ? box(CB_V64_STRING, nT + sm[id - nT]) : w; }
                    if ((tag == CB_V64_LIST || tag == CB_V64_MAP) && (w & CB_V64_HEAP_BATCH_BIT)) return w + hb;
                    return w;
                };
                for (uint64_t i = lo; i < hi; i++) {
                    uint32_t &pid = c.hdr0[i * 4];
                    if (pid >= nT) pid = (uint32_t)(nT + sm[pid - nT]);
                    uint32_t &kc = c.hdr0[i * 4 + 1];
                    if (kc != CB_KIND_NONE && (kc & CB_KIND_CLASS_CSR_BIT)) kc = class_map[t][kc & ~CB_KIND_CLASS_CSR_BIT] | CB_KIND_CLASS_CSR_BIT;
                    uint32_t aid;
                    memcpy(&aid, &c.hdr1[i * 8 + 4], 4);
                    aid = aset_map[t][aid];
                    memcpy(&c.hdr1[i * 8 + 4], &aid, 4);
                    for (size_t s2 = 0; s2 < n_slots; s2++) c.slots[s2 * n + i] = fix(c.slots[s2 * n + i]);
                }
                // heap records: [n, elements...] / [n, keys..., values...] -- the count words are plain integers (never boxed)
                const std::vector<uint64_t> &h = heaps[t];
                for (size_t j = 0; j < h.size(); j++) c.heap[hb + j] = fix(h[j]);
            });
        }
        shards[0].finish();
        return true;
    }

    template <typename F>
    static void run(unsigned n_threads, F f) {
        if (n_threads <= 1) { f(0); return; }
        std::vector<std::thread> th;
        for (unsigned t = 1; t < n_threads; t++) th.emplace_back([&f, t] { f(t); });
        f(0);
        for (auto &x : th) x.join();
    }

  private:
    uint16_t version_id(const std::string &v) const { auto it = version_ids.find(v); return it == version_ids.end() ? (uint16_t)CB_NONE16 : (uint16_t)it->second; }
    uint32_t resolve_scope(const std::string &scope) const {
        auto it = scope_ids.find(scope);
        if (it != scope_ids.end()) return it->second;
        if (conf.lenient)
            for (size_t i = scope.size(); i-- > 0;)      // every dotted prefix from longest to "" (namer.go:77-87)
                if (scope[i] == '.' || i == 0) {
                    auto a = scope_ids.find(scope.substr(0, i));
                    if (a != scope_ids.end()) return a->second | CB_SCOPE_INEXACT_BIT;
                }
        return CB_SCOPE_NONE;
    }
    static bool parse_attr(WireIt &it, std::vector<std::pair<Span, Span>> *attr) {
        Span k, v;
        if (!map_entry(it.s, &k, &v)) return false;
        attr->emplace_back(k, v);
        return true;
    }
    // roles of the principal and actions of one CheckInput message, without building a view
    static bool scan_counts(Span msg, uint32_t *n_roles, uint32_t *n_actions) {
        WireIt it(msg);
        while (it.next()) {
            if (it.wt != 2) continue;
            if (it.fno == 3) {
                WireIt p(it.s);
                while (p.next()) if (p.wt == 2 && p.fno == 3) (*n_roles)++;
                if (p.bad) return false;
            } else if (it.fno == 4) (*n_actions)++;
        }
        return !it.bad;
    }
    static bool parse_input(Span msg, InputView *out) {
        WireIt it(msg);
        while (it.next()) {
            if (it.wt != 2) continue;
            if (it.fno == 2) {
                WireIt r(it.s);
                while (r.next()) {
                    if (r.wt != 2) continue;
                    if (r.fno == 1) out->r.kind = r.s; else if (r.fno == 2) out->r.version = r.s; else if (r.fno == 3) out->r.id = r.s;
                    else if (r.fno == 4) { if (!parse_attr(r, &out->r.attr)) return false; } else if (r.fno == 5) out->r.scope = r.s;
                }
                if (r.bad) return false;
            } else if (it.fno == 3) {
                WireIt p(it.s);
                while (p.next()) {
                    if (p.wt != 2) continue;
                    if (p.fno == 1) out->p.id = p.s; else if (p.fno == 2) out->p.version = p.s; else if (p.fno == 3) out->p.roles.push_back(p.s);
                    else if (p.fno == 4) { if (!parse_attr(p, &out->p.attr)) return false; } else if (p.fno == 5) out->p.scope = p.s;
                }
                if (p.bad) return false;
            } else if (it.fno == 4) out->actions.push_back(it.s);
            else if (it.fno == 5) {
                out->has_aux = true;
                WireIt a(it.s);
                while (a.next()) if (a.fno == 1 && a.wt == 2 && !parse_attr(a, &out->jwt)) return false;
                if (a.bad) return false;
            }
        }
        return !it.bad;
    }

    // per-batch state: string dictionary, heap, kind classes, action sets
    struct State {
        const Encoder *E;
        Columns *c;
        std::vector<uint64_t> *heap;     // where lists / maps go: the batch heap itself, or a shard-local one merged later
        BytesMap bstr;                   // the batch string dictionary: entry j = batch string j (first appearance order)
        std::unordered_map<std::string, std::vector<uint32_t>> class_cache;
        std::map<std::vector<uint32_t>, uint32_t> classes;
        std::vector<std::vector<uint32_t>> class_list;
        std::map<std::vector<std::string>, uint32_t> asets;
        std::vector<std::vector<std::string>> aset_list;
        State(const Encoder *e, Columns *cc) : E(e), c(cc), heap(&cc->heap) {}

        uint32_t sid(const uint8_t *p, size_t n) {
            const uint64_t h = hash_bytes(p, n);
            if (const uint32_t *t = E->table_strings.find(p, n, h)) return *t;
            if (const uint32_t *b = bstr.find(p, n, h)) return (uint32_t)E->strings.size() + *b;
            const uint32_t i = (uint32_t)bstr.size();
            bstr.insert(p, n, h, i);
            return (uint32_t)E->strings.size() + i;
        }
        uint32_t sid(const std::string &s) { return sid(reinterpret_cast<const uint8_t *>(s.data()), s.size()); }
        uint32_t sid(Span s) { return sid(s.p, s.n); }

        uint32_t kind_class(const std::string &kind, std::string *err) {
            auto it = class_cache.find(kind);
            if (it == class_cache.end()) {
                std::vector<uint32_t> pats;
                const std::string sk = sanitize(kind);
                for (uint32_t i = 0; i < E->respats.size(); i++) if (key_matches(E->respats[i], sk)) pats.push_back(i);
                if (pats.size() > CB_MAX_CLASS_PATS) { *err = "resource kind matches more than 8 resource patterns"; return CB_KIND_NONE; }
                it = class_cache.emplace(kind, std::move(pats)).first;
            }
            const std::vector<uint32_t> &kp = it->second;
            if (kp.empty()) return CB_KIND_NONE;
            if (kp.size() == 1) return kp[0];
            auto ci = classes.find(kp);
            if (ci == classes.end()) { ci = classes.emplace(kp, (uint32_t)class_list.size()).first; class_list.push_back(kp); }
            return ci->second | CB_KIND_CLASS_CSR_BIT;
        }
        uint32_t action_set(const std::vector<Span> &actions) {
            std::vector<std::string> key;
            for (Span a : actions) key.push_back(str_of(a));
            auto it = asets.find(key);
            if (it == asets.end()) { it = asets.emplace(key, (uint32_t)aset_list.size()).first; aset_list.push_back(key); }
            return it->second;
        }

        // google.protobuf.Value -> NaN-boxed 8 bytes (lists / maps go to the heap, children before their parent: the layout
        // cerbos_b200/encode.py produces)
        uint64_t v64(Span val) {
            WireIt it(val);
            uint64_t out = V_NULL;
            while (it.next()) {
                switch (it.fno) {
                case 1: out = V_NULL; break;
                case 2: { double d; memcpy(&d, &it.u, 8); out = d != d ? (uint64_t)CB_V64_CANON_NAN : it.u; break; }
                case 3: out = box(CB_V64_STRING, sid(it.s)); break;
                case 4: out = box(CB_V64_BOOL, it.u ? 1 : 0); break;
                case 5: out = v64_struct(it.s); break;
                case 6: {   // (children are written to the heap before their parent: the elements are collected first)
                    uint64_t small[16];
                    std::vector<uint64_t> big;
                    size_t ne = 0;
                    WireIt l(it.s);
                    while (l.next()) if (l.fno == 1 && l.wt == 2) {
                        const uint64_t w = v64(l.s);
                        if (ne < 16) small[ne] = w; else { if (ne == 16) big.assign(small, small + 16); big.push_back(w); }
                        ne++;
                    }
                    const uint64_t off = heap->size();
                    heap->push_back(ne);
                    if (ne <= 16) heap->insert(heap->end(), small, small + ne); else heap->insert(heap->end(), big.begin(), big.end());
                    out = box(CB_V64_LIST, off | CB_V64_HEAP_BATCH_BIT);
                    break;
                }
                default: break;
                }
            }
            return out;
        }
        uint64_t v64_struct(Span st) {     // google.protobuf.Struct { map<string, Value> fields = 1 }
            std::pair<Span, Span> small[12];
            size_t ne = 0;
            WireIt f(st);
            while (f.next()) if (f.fno == 1 && f.wt == 2) {
                Span k, v;
                if (!map_entry(f.s, &k, &v)) continue;
                if (ne == 12) {     // a larger object: the general path
                    std::vector<std::pair<Span, Span>> ents(small, small + 12);
                    ents.emplace_back(k, v);
                    while (f.next()) if (f.fno == 1 && f.wt == 2) { Span k2, v2; if (map_entry(f.s, &k2, &v2)) ents.emplace_back(k2, v2); }
                    return v64_map(ents.data(), ents.size());
                }
                small[ne++] = std::make_pair(k, v);
            }
            return v64_map(small, ne);
        }
        uint64_t v64_map(const std::vector<std::pair<Span, Span>> &ents_in) { return v64_map(ents_in.data(), ents_in.size()); }
        uint64_t v64_map(const std::pair<Span, Span> *ents_in, size_t n_in) {
            // a repeated key keeps its first position and its last value (dict semantics of the JSON path)
            std::pair<Span, Span> small[12];
            std::vector<std::pair<Span, Span>> big;
            if (n_in > 12) big.reserve(n_in);
            std::pair<Span, Span> *ents = n_in > 12 ? nullptr : small;
            size_t ne = 0;
            for (size_t q = 0; q < n_in; q++) {
                const auto &e = ents_in[q];
                std::pair<Span, Span> *cur = n_in > 12 ? big.data() : small;
                bool dup = false;
                for (size_t z = 0; z < ne; z++) if (span_eq(cur[z].first, e.first)) { cur[z].second = e.second; dup = true; break; }
                if (!dup) { if (n_in > 12) big.push_back(e); else small[ne] = e; ne++; }
            }
            ents = n_in > 12 ? big.data() : small;
            // keys are interned in order, then the values are encoded (their heap records precede the map's own)
            uint64_t ksmall[12], vsmall[12];
            std::vector<uint64_t> kbig, vbig;
            uint64_t *keys = ksmall, *vals = vsmall;
            if (ne > 12) { kbig.resize(ne); vbig.resize(ne); keys = kbig.data(); vals = vbig.data(); }
            for (size_t z = 0; z < ne; z++) keys[z] = box(CB_V64_STRING, sid(ents[z].first));
            for (size_t z = 0; z < ne; z++) vals[z] = ents[z].second.p ? v64(ents[z].second) : V_NULL;
            const uint64_t off = heap->size();
            heap->push_back(ne);
            heap->insert(heap->end(), keys, keys + ne);
            heap->insert(heap->end(), vals, vals + ne);
            return box(CB_V64_MAP, off | CB_V64_HEAP_BATCH_BIT);
        }
        uint64_t v64_string(const std::string &s) { return box(CB_V64_STRING, sid(s)); }

        static const Span *find(const std::vector<std::pair<Span, Span>> &ents, const std::string &key) {
            const Span *hit = nullptr;
            for (const auto &e : ents) if (e.first.n == key.size() && (key.empty() || !memcmp(e.first.p, key.data(), key.size()))) hit = &e.second;   // last wins
            return hit;
        }
        // attr[segs[from]] [segs[from + 1]] ...: ABSENT when only the last segment is missing, ERROR when the path leaves the maps
        uint64_t walk(const std::vector<std::pair<Span, Span>> &root, const std::vector<std::string> &segs, size_t from) {
            const Span *cur = find(root, segs[from]);
            if (!cur) return from + 1 == segs.size() ? V_ABSENT : V_ERROR;
            Span val = *cur;
            for (size_t j = from + 1; j < segs.size(); j++) {
                // val must be a struct Value
                Span st{};
                bool is_struct = false;
                WireIt it(val);
                while (it.next()) { if (it.fno == 5 && it.wt == 2) { st = it.s; is_struct = true; } else is_struct = false; }
                if (!is_struct) return V_ERROR;
                Span next{};
                bool found = false;
                WireIt f(st);
                while (f.next()) if (f.fno == 1 && f.wt == 2) { Span k, v; if (map_entry(f.s, &k, &v) && k.n == segs[j].size() && (k.n == 0 || !memcmp(k.p, segs[j].data(), k.n))) { next = v; found = true; } }
                if (!found) return j + 1 == segs.size() ? V_ABSENT : V_ERROR;
                val = next;
            }
            return val.p ? v64(val) : V_NULL;
        }
        uint64_t slot_value(const InputView &v, const std::vector<std::string> &path, SlotSrc src) {
            switch (src) {
            case SRC_ERROR: return V_ERROR;
            case SRC_AUX: return path.size() > 2 ? walk(v.jwt, path, 2) : v64_map(v.jwt);
            case SRC_P_ATTR: return path.size() > 2 ? walk(v.p.attr, path, 2) : v64_map(v.p.attr);
            case SRC_R_ATTR: return path.size() > 2 ? walk(v.r.attr, path, 2) : v64_map(v.r.attr);
            case SRC_P_ROLES: case SRC_R_ROLES: {
                std::vector<uint64_t> elems;
                if (src == SRC_P_ROLES) for (Span r : v.p.roles) elems.push_back(box(CB_V64_STRING, sid(r)));
                const uint64_t off = heap->size();
                heap->push_back(elems.size());
                heap->insert(heap->end(), elems.begin(), elems.end());
                return box(CB_V64_LIST, off | CB_V64_HEAP_BATCH_BIT);
            }
            case SRC_P_SCOPE: case SRC_R_SCOPE: return v64_string(scope_value(str_of(src == SRC_P_SCOPE ? v.p.scope : v.r.scope)));
            case SRC_P_VERSION: case SRC_R_VERSION: return box(CB_V64_STRING, sid(src == SRC_P_VERSION ? v.p.version : v.r.version));
            case SRC_P_ID: case SRC_R_ID: return box(CB_V64_STRING, sid(src == SRC_P_ID ? v.p.id : v.r.id));
            case SRC_R_KIND: return box(CB_V64_STRING, sid(v.r.kind));
            default: return v64_string(std::string());   // SRC_P_KIND, unknown fields: the empty string
            }
        }

        void finish() {
            if (c->heap.empty()) c->heap.push_back(0);
            // kind classes (CSR)
            c->class_off.assign(class_list.size() + 1, 0);
            for (size_t k = 0; k < class_list.size(); k++) { c->class_off[k] = (uint32_t)c->class_pats.size(); c->class_pats.insert(c->class_pats.end(), class_list[k].begin(), class_list[k].end()); }
            c->class_off[class_list.size()] = (uint32_t)c->class_pats.size();
            if (c->class_pats.empty()) c->class_pats.push_back(0);
            // action sets: per (pass, set, action pattern) the (action x role column) bit spread; then OR-ed over the patterns of every table row
            const size_t n_ap = E->apats.empty() ? 1 : E->apats.size(), n_as = aset_list.empty() ? 1 : aset_list.size();
            c->aset_k.assign(n_as, 0);
            c->aset_spread.assign((size_t)c->n_pass * n_as * n_ap, 0);
            std::unordered_map<std::string, std::vector<uint32_t>> apat_cache;
            for (size_t a = 0; a < aset_list.size(); a++) {
                c->aset_k[a] = (uint32_t)aset_list[a].size();
                for (size_t k = 0; k < aset_list[a].size(); k++) {
                    const std::string &act = aset_list[a][k];
                    auto it = apat_cache.find(act);
                    if (it == apat_cache.end()) {
                        std::vector<uint32_t> m;
                        for (uint32_t i = 0; i < E->apats.size(); i++) if (key_matches(E->apats[i], act)) m.push_back(i);
                        it = apat_cache.emplace(act, std::move(m)).first;
                    }
                    const size_t ps = k / c->kc, kk = k % c->kc;
                    for (uint32_t ap : it->second) c->aset_spread[(ps * n_as + a) * n_ap + ap] |= 1ull << (kk * c->role_cols);
                }
            }
            const size_t n_rows = E->row_pat_start.size();
            const size_t nr = n_rows ? n_rows : 1;
            c->row_am.assign((size_t)c->n_pass * n_as * nr, 0);
            for (size_t ps = 0; ps < c->n_pass; ps++)
                for (size_t a = 0; a < n_as; a++)
                    for (size_t r = 0; r < n_rows; r++) {
                        const size_t lo = E->row_pat_start[r], hi = r + 1 < n_rows ? E->row_pat_start[r + 1] : E->row_apats.size();
                        uint64_t m = 0;
                        for (size_t q = lo; q < hi; q++) m |= c->aset_spread[(ps * n_as + a) * n_ap + E->row_apats[q]];
                        c->row_am[(ps * n_as + a) * nr + r] = m;
                    }
            // batch string dictionary
            // (entries are stored in order of first appearance and their keys back to back in the arena: that IS the dictionary)
            c->bstr_off.assign(bstr.size() + 1, 0);
            for (size_t j = 0; j < bstr.size(); j++) c->bstr_off[j] = bstr.ents[j].off;
            c->bstr_off[bstr.size()] = (uint32_t)bstr.arena.size();
            c->bstr_bytes.reserve(bstr.arena.size() + 16);
            c->bstr_bytes.assign(bstr.arena.begin(), bstr.arena.end());
            c->bstr_bytes.insert(c->bstr_bytes.end(), 16, 0);
        }
    };
};

}  // namespace cbenc
