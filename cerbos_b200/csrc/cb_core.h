// cb_core.h -- per-request evaluation core of the B200 CheckResources kernels.
//
// One thread evaluates one request (principal, resource, K actions) against the flattened rule table:
//   * scope chains + existence checks      (reference: ruletable.go:611-645, 804-863; index.go:1089-1172)
//   * ONE pass over the rows of every policy block on the chain; every satisfied row contributes a
//     bit pattern (action x role-column) so the reference's per-action / per-role walk
//     (ruletable.go:885-1152) becomes a handful of 64-bit mask operations ("bit-parallel walk"):
//         within a scope   DENY beats ALLOW                         (:1083-1091)
//         OVERRIDE_PARENT  satisfied ALLOW finishes the (action, role) pair     (:1115-1118)
//         REQUIRE_PARENTAL_CONSENT  ALLOW is dropped, walk continues   (:1113-1114)
//         an action is ALLOWed iff the principal-policy walk allows it, or it is undecided there
//         and some role's resource-policy walk allows it            (:1124-1148)
//   * role-policy DENY synthesis             (index.go:688-776)
//   * CEL conditions by a stack bytecode interpreter with cel-go error semantics
//     (bytecode produced by cerbos_b200/table/bytecode.py; leaf rule ruletable.go:1425-1441)
//
// The file is plain C++ guarded by CB_HD so that tests/hostsim can compile the very same code for the
// host and step through it without a GPU (debug aid only -- the product never runs it on the CPU).
#pragma once
#include <stdint.h>

#include "cerbos_b200_format.h"

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#define CB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CB_HD inline
#define CB_HD_NOINLINE
#endif

namespace cb {

// ---------------------------------------------------------------------------------------------- views
struct TableView {
    const uint32_t *scope_parent, *scope_flags, *res_block_map, *prin_block_map, *prin_of_string;
    const uint8_t *res_exists, *prin_exists;
    const cb_block *blocks;
    const cb_row *rows;
    const cb_cond *conds;
    const cb_instr *code;
    const cb_const *consts;
    const uint64_t *theap;
    const uint32_t *str_off;
    const uint8_t *str_bytes;
    const uint32_t *par_off, *par_list, *rp_off, *rp_apats;
    const cb_rolepol_entry *rp_entries;
    const cb_rolepol_rule *rp_rules;
    uint32_t nV, nRP, nS, nP, nR, nAP, nT, n_slots;
    uint32_t has_role_policies, has_parent_roles, has_principal_policies;
};

struct BatchView {
    const cb_hdr0 *hdr0;
    const cb_hdr1 *hdr1;
    const uint32_t *roles;      // [role_cols][stride]
    const uint64_t *slots;      // [n_slots][stride]
    const uint64_t *heap;
    const uint32_t *bstr_off;
    const uint8_t *bstr_bytes;
    const uint32_t *class_off, *class_pats, *aset_k;
    const uint64_t *aset_spread;  // [n_pass][n_asets][nAP]
    uint64_t stride;            // requests per column (N of the whole batch)
    uint64_t first, count;      // sub-range evaluated by this launch
    uint32_t role_cols, n_asets, kc, n_pass, max_actions, kbytes, flags;
    int64_t now;
};

// Table data may live in shared memory (TMA-staged image) or in global memory, heap references may point
// into either the table or the batch: those loads are plain (generic) loads.  Only the big streaming request
// columns -- read exactly once -- use the read-only, no-L1-allocate path so they do not evict the table.
template <typename T>
CB_HD T ldg(const T *p) { return *p; }

CB_HD uint64_t ldcol64(const uint64_t *p) {
#if defined(__CUDA_ARCH__)
    uint64_t v;
    asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
CB_HD uint32_t ldcol32(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
    uint32_t v;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}

// 128-bit loads of the 16-byte records (their C structs are only 4-byte aligned, the buffers are 16-byte aligned)
struct alignas(16) U4 { uint32_t x, y, z, w; };
CB_HD U4 ld16(const void *p) { return *reinterpret_cast<const U4 *>(p); }
CB_HD U4 ldcol128(const void *p) {
#if defined(__CUDA_ARCH__)
    U4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#else
    return *reinterpret_cast<const U4 *>(p);
#endif
}
CB_HD cb_hdr0 load_hdr0(const cb_hdr0 *p) { U4 v = ldcol128(p); cb_hdr0 h; h.principal_id = v.x; h.kind_class = v.y; h.resource_scope = v.z; h.principal_scope = v.w; return h; }
CB_HD cb_hdr1 load_hdr1(const cb_hdr1 *p) {
    uint64_t v = ldcol64(reinterpret_cast<const uint64_t *>(p));
    cb_hdr1 h; h.resource_version = (uint16_t)(v & 0xFFFF); h.principal_version = (uint16_t)((v >> 16) & 0xFFFF); h.action_set_id = (uint32_t)(v >> 32); return h;
}
CB_HD cb_block load_block(const cb_block *p) { U4 v = ld16(p); cb_block b; b.row_start = v.x; b.n_rows = v.y; b.cond_base = v.z; b.n_conds = v.w; return b; }
CB_HD cb_row load_row(const cb_row *p) {
    U4 v = ld16(p);
    cb_row r; r.apat = (uint16_t)(v.x & 0xFFFF); r.role = (uint16_t)(v.x >> 16); r.cond = (uint16_t)(v.y & 0xFFFF); r.drcond = (uint16_t)(v.y >> 16);
    r.respat = (uint16_t)(v.z & 0xFFFF); r.effect = (uint8_t)((v.z >> 16) & 0xFF); r.flags = (uint8_t)(v.z >> 24); r.pad = v.w; return r;
}
CB_HD cb_rolepol_entry load_rp_entry(const cb_rolepol_entry *p) { U4 v = ld16(p); cb_rolepol_entry e; e.role = v.x; e.rule_start = v.y; e.n_rules = v.z; e.pad = v.w; return e; }
CB_HD cb_rolepol_rule load_rp_rule(const cb_rolepol_rule *p) { U4 v = ld16(p); cb_rolepol_rule e; e.respat = v.x; e.cond = v.y; e.apat_start = v.z; e.n_apats = v.w; return e; }

// ---------------------------------------------------------------------------------------------- values
struct Val {
    uint32_t tag;
    uint64_t u;
};
static constexpr uint64_t kHeapBatch = 1ull << 63;

CB_HD Val mk(uint32_t tag, uint64_t u) { Val v; v.tag = tag; v.u = u; return v; }
CB_HD Val mk_err() { return mk(CB_T_ERR, 0); }
CB_HD Val mk_bool(bool b) { return mk(CB_T_BOOL, b ? 1u : 0u); }
CB_HD Val mk_int(int64_t i) { return mk(CB_T_INT, (uint64_t)i); }
CB_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; __builtin_memcpy(&d, &u, 8); return d;
#endif
}
CB_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; __builtin_memcpy(&u, &d, 8); return u;
#endif
}
CB_HD Val mk_double(double d) { return mk(CB_T_DOUBLE, d != d ? (uint64_t)CB_V64_CANON_NAN : d2u(d)); }

enum { SLOT_VALUE = 0, SLOT_ABSENT = 1, SLOT_ERROR = 2 };

CB_HD Val decode_v64(uint64_t bits, int *state) {
    *state = SLOT_VALUE;
    uint32_t top = (uint32_t)(bits >> 48);
    if ((top & 0xFFF0u) == 0xFFF0u && (top & 0xFu) != 0) {
        uint32_t tag = top & 0xFu;
        uint64_t pay = bits & 0xFFFFFFFFFFFFull;
        switch (tag) {
        case CB_V64_NULL: return mk(CB_T_NULL, 0);
        case CB_V64_BOOL: return mk_bool(pay != 0);
        case CB_V64_STRING: return mk(CB_T_STRING, pay);
        case CB_V64_LIST:
        case CB_V64_MAP: {
            uint64_t off = pay & (CB_V64_HEAP_BATCH_BIT - 1);
            if (pay & CB_V64_HEAP_BATCH_BIT) off |= kHeapBatch;
            return mk(tag == CB_V64_LIST ? CB_T_LIST : CB_T_MAP, off);
        }
        case CB_V64_INT: return mk_int((int64_t)(pay << 16) >> 16);
        case CB_V64_ABSENT: *state = SLOT_ABSENT; return mk_err();
        default: *state = SLOT_ERROR; return mk_err();
        }
    }
    return mk(CB_T_DOUBLE, bits);
}
CB_HD Val decode_elem(uint64_t bits) { int s; return decode_v64(bits, &s); }

struct Ctx {
    const TableView *t;
    const BatchView *b;
    uint64_t req;          // absolute request index (column index)
    uint32_t pid;          // hdr0.principal_id
    uint32_t unsupported;  // sticky
    Val vars[CB_MAX_VARS];
};

CB_HD const uint64_t *heap_ptr(const Ctx &c, uint64_t ref) {
    return (ref & kHeapBatch) ? c.b->heap + (ref & ~kHeapBatch) : c.t->theap + ref;
}
CB_HD void str_get(const Ctx &c, uint64_t id, const uint8_t *&p, uint32_t &len) {
    if (id < c.t->nT) {
        uint32_t o = ldg(c.t->str_off + id);
        p = c.t->str_bytes + o;
        len = ldg(c.t->str_off + id + 1) - o;
    } else {
        uint64_t j = id - c.t->nT;
        uint32_t o = ldg(c.b->bstr_off + j);
        p = c.b->bstr_bytes + o;
        len = ldg(c.b->bstr_off + j + 1) - o;
    }
}

CB_HD bool is_num(const Val &v) { return v.tag == CB_T_INT || v.tag == CB_T_UINT || v.tag == CB_T_DOUBLE; }

// cel-go cross-type numeric comparison (types/compare.go): -1/0/1, 2 = unordered (NaN)
CB_HD int num_cmp(const Val &a, const Val &b) {
    if (a.tag == CB_T_DOUBLE || b.tag == CB_T_DOUBLE) {
        if (a.tag == CB_T_DOUBLE && b.tag == CB_T_DOUBLE) {
            double x = u2d(a.u), y = u2d(b.u);
            if (x != x || y != y) return 2;
            return x < y ? -1 : (x > y ? 1 : 0);
        }
        int sign = 1;
        Val dv = a, iv = b;
        if (a.tag != CB_T_DOUBLE) { dv = b; iv = a; sign = -1; }
        double d = u2d(dv.u);
        if (d != d) return 2;
        int r;
        if (iv.tag == CB_T_UINT) {
            if (d < 0) r = -1;
            else if (d > 18446744073709551615.0) r = 1;
            else { double y = (double)iv.u; r = d < y ? -1 : (d > y ? 1 : 0); }
        } else {
            if (d < -9223372036854775808.0) r = -1;
            else if (d > 9223372036854775807.0) r = 1;
            else { double y = (double)(int64_t)iv.u; r = d < y ? -1 : (d > y ? 1 : 0); }
        }
        return r * sign;
    }
    if (a.tag == b.tag) {
        if (a.tag == CB_T_INT) { int64_t x = (int64_t)a.u, y = (int64_t)b.u; return x < y ? -1 : (x > y ? 1 : 0); }
        return a.u < b.u ? -1 : (a.u > b.u ? 1 : 0);
    }
    if (a.tag == CB_T_INT) {
        int64_t x = (int64_t)a.u;
        if (x < 0) return -1;
        return (uint64_t)x < b.u ? -1 : ((uint64_t)x > b.u ? 1 : 0);
    }
    int64_t y = (int64_t)b.u;
    if (y < 0) return 1;
    return a.u < (uint64_t)y ? -1 : (a.u > (uint64_t)y ? 1 : 0);
}

// scalar (non-container) equality; containers handled by the callers below
CB_HD bool scalar_equal(const Val &a, const Val &b) {
    if (is_num(a) && is_num(b)) return num_cmp(a, b) == 0;
    if (a.tag != b.tag) return false;
    if (a.tag == CB_T_NULL) return true;
    return a.u == b.u;  // BOOL / STRING (interned ids) / TS / DUR
}
CB_HD bool is_container(const Val &v) { return v.tag == CB_T_LIST || v.tag == CB_T_MAP; }

CB_HD bool map_find(const Ctx &c, const Val &m, const Val &key, Val *out) {
    if (key.tag != CB_T_STRING) return false;  // JSON / constant maps have string keys only
    const uint64_t *p = heap_ptr(c, m.u);
    uint64_t n = ldg(p);
    for (uint64_t i = 0; i < n; i++) {
        Val k = decode_elem(ldg(p + 1 + i));
        if (k.tag == CB_T_STRING && k.u == key.u) {
            if (out) *out = decode_elem(ldg(p + 1 + n + i));
            return true;
        }
    }
    return false;
}

// Heterogeneous equality (cel-go types.Equal).  Containers are compared to a nesting depth of 3;
// deeper structures raise the sticky `unsupported` flag (the call then fails loudly).
template <int DEPTH>
struct Eq {
    static CB_HD bool eq(Ctx &c, const Val &a, const Val &b) {
        if (!is_container(a) || !is_container(b)) {
            if (is_container(a) != is_container(b)) return false;
            return scalar_equal(a, b);
        }
        if (a.tag != b.tag) return false;
        const uint64_t *pa = heap_ptr(c, a.u), *pb = heap_ptr(c, b.u);
        uint64_t n = ldg(pa);
        if (n != ldg(pb)) return false;
        if (a.tag == CB_T_LIST) {
            for (uint64_t i = 0; i < n; i++)
                if (!Eq<DEPTH - 1>::eq(c, decode_elem(ldg(pa + 1 + i)), decode_elem(ldg(pb + 1 + i)))) return false;
            return true;
        }
        for (uint64_t i = 0; i < n; i++) {
            Val ov;
            if (!map_find(c, b, decode_elem(ldg(pa + 1 + i)), &ov)) return false;
            if (!Eq<DEPTH - 1>::eq(c, decode_elem(ldg(pa + 1 + n + i)), ov)) return false;
        }
        return true;
    }
};
template <>
struct Eq<0> {
    static CB_HD bool eq(Ctx &c, const Val &a, const Val &b) {
        if (is_container(a) && is_container(b)) { c.unsupported = 1; return false; }
        if (is_container(a) != is_container(b)) return false;
        return scalar_equal(a, b);
    }
};
CB_HD bool val_equal(Ctx &c, const Val &a, const Val &b) { return Eq<3>::eq(c, a, b); }

CB_HD int str_cmp(const Ctx &c, uint64_t ia, uint64_t ib) {
    const uint8_t *pa, *pb;
    uint32_t la, lb;
    str_get(c, ia, pa, la);
    str_get(c, ib, pb, lb);
    uint32_t m = la < lb ? la : lb;
    for (uint32_t i = 0; i < m; i++) {
        uint8_t x = ldg(pa + i), y = ldg(pb + i);
        if (x != y) return x < y ? -1 : 1;
    }
    return la < lb ? -1 : (la > lb ? 1 : 0);
}

// -1/0/1, 3 = error (no such overload / NaN)
CB_HD int val_order(const Ctx &c, const Val &a, const Val &b) {
    if (is_num(a) && is_num(b)) { int r = num_cmp(a, b); return r == 2 ? 3 : r; }
    if (a.tag != b.tag) return 3;
    switch (a.tag) {
    case CB_T_BOOL: return a.u < b.u ? -1 : (a.u > b.u ? 1 : 0);
    case CB_T_STRING: return a.u == b.u ? 0 : str_cmp(c, a.u, b.u);
    case CB_T_TS:
    case CB_T_DUR: { int64_t x = (int64_t)a.u, y = (int64_t)b.u; return x < y ? -1 : (x > y ? 1 : 0); }
    default: return 3;
    }
}

CB_HD Val do_cmp(Ctx &c, int ci, const Val &a, const Val &b) {
    if (a.tag == CB_T_ERR || b.tag == CB_T_ERR) return mk_err();
    if (ci == 0) return mk_bool(val_equal(c, a, b));
    if (ci == 1) return mk_bool(!val_equal(c, a, b));
    int r = val_order(c, a, b);
    if (r == 3) return mk_err();
    switch (ci) {
    case 2: return mk_bool(r < 0);
    case 3: return mk_bool(r <= 0);
    case 4: return mk_bool(r > 0);
    default: return mk_bool(r >= 0);
    }
}

CB_HD Val do_in(Ctx &c, const Val &x, const Val &cont) {
    if (x.tag == CB_T_ERR || cont.tag == CB_T_ERR) return mk_err();
    if (cont.tag == CB_T_LIST) {
        const uint64_t *p = heap_ptr(c, cont.u);
        uint64_t n = ldg(p);
        for (uint64_t i = 0; i < n; i++)
            if (val_equal(c, x, decode_elem(ldg(p + 1 + i)))) return mk_bool(true);
        return mk_bool(false);
    }
    if (cont.tag == CB_T_MAP) return mk_bool(map_find(c, cont, x, nullptr));
    return mk_err();
}

CB_HD Val do_index(Ctx &c, const Val &cont, const Val &key) {
    if (cont.tag == CB_T_ERR || key.tag == CB_T_ERR) return mk_err();
    if (cont.tag == CB_T_LIST) {
        int64_t idx;
        if (key.tag == CB_T_INT) idx = (int64_t)key.u;
        else if (key.tag == CB_T_UINT) { if (key.u > 0x7FFFFFFFFFFFFFFFull) return mk_err(); idx = (int64_t)key.u; }
        else if (key.tag == CB_T_DOUBLE) {
            double d = u2d(key.u);
            if (!(d == (double)(int64_t)d) || !(d > -9.2e18 && d < 9.2e18)) return mk_err();
            idx = (int64_t)d;
        } else return mk_err();
        const uint64_t *p = heap_ptr(c, cont.u);
        if (idx < 0 || (uint64_t)idx >= ldg(p)) return mk_err();
        return decode_elem(ldg(p + 1 + idx));
    }
    if (cont.tag == CB_T_MAP) { Val out; return map_find(c, cont, key, &out) ? out : mk_err(); }
    return mk_err();
}

// ---- Cerbos set functions (cerbos_lib.go:323-431).  When the larger list has > 3 elements that are all
// hashable the reference probes a Go map keyed by ref.Val: identity is (dynamic type, value), i.e. no
// cross-type numeric equality; otherwise it scans with Equal. ----
CB_HD bool hashable(const Val &v) {
    return v.tag == CB_T_STRING || v.tag == CB_T_INT || v.tag == CB_T_UINT || v.tag == CB_T_DOUBLE || v.tag == CB_T_DUR || v.tag == CB_T_TS;
}
CB_HD bool uses_go_map(const Ctx &c, const Val &b) {
    const uint64_t *p = heap_ptr(c, b.u);
    uint64_t n = ldg(p);
    if (n <= 3) return false;
    for (uint64_t i = 0; i < n; i++)
        if (!hashable(decode_elem(ldg(p + 1 + i)))) return false;
    return true;
}
CB_HD bool key_identical(const Val &a, const Val &b) {
    if (a.tag != b.tag) return false;
    if (a.tag == CB_T_DOUBLE) return u2d(a.u) == u2d(b.u);
    return a.u == b.u;
}
CB_HD bool list_member(Ctx &c, bool go_map, const Val &b, const Val &x) {
    const uint64_t *p = heap_ptr(c, b.u);
    uint64_t n = ldg(p);
    for (uint64_t i = 0; i < n; i++) {
        Val e = decode_elem(ldg(p + 1 + i));
        if (go_map ? key_identical(x, e) : val_equal(c, x, e)) return true;
    }
    return false;
}
CB_HD Val do_set_pred(Ctx &c, bool subset, Val a, Val b) {
    if (a.tag != CB_T_LIST || b.tag != CB_T_LIST) return mk_err();
    if (!subset && ldg(heap_ptr(c, a.u)) > ldg(heap_ptr(c, b.u))) { Val t = a; a = b; b = t; }
    bool gm = uses_go_map(c, b);
    const uint64_t *p = heap_ptr(c, a.u);
    uint64_t n = ldg(p);
    for (uint64_t i = 0; i < n; i++) {
        bool m = list_member(c, gm, b, decode_elem(ldg(p + 1 + i)));
        if (subset && !m) return mk_bool(false);
        if (!subset && m) return mk_bool(true);
    }
    return mk_bool(subset);
}

// ---- arithmetic with cel-go overflow rules ----
#if defined(__CUDA_ARCH__)
CB_HD bool add_ovf(int64_t x, int64_t y, int64_t *r) { int64_t s = (int64_t)((uint64_t)x + (uint64_t)y); *r = s; return ((x ^ s) & (y ^ s)) < 0; }
CB_HD bool sub_ovf(int64_t x, int64_t y, int64_t *r) { int64_t s = (int64_t)((uint64_t)x - (uint64_t)y); *r = s; return ((x ^ y) & (x ^ s)) < 0; }
CB_HD bool mul_ovf(int64_t x, int64_t y, int64_t *r) {
    int64_t lo = (int64_t)((uint64_t)x * (uint64_t)y);
    int64_t hi = __mul64hi(x, y);
    *r = lo;
    return hi != (lo >> 63);
}
CB_HD bool umul_ovf(uint64_t x, uint64_t y, uint64_t *r) { *r = x * y; return __umul64hi(x, y) != 0; }
#else
CB_HD bool add_ovf(int64_t x, int64_t y, int64_t *r) { return __builtin_add_overflow(x, y, r); }
CB_HD bool sub_ovf(int64_t x, int64_t y, int64_t *r) { return __builtin_sub_overflow(x, y, r); }
CB_HD bool mul_ovf(int64_t x, int64_t y, int64_t *r) { return __builtin_mul_overflow(x, y, r); }
CB_HD bool umul_ovf(uint64_t x, uint64_t y, uint64_t *r) { return __builtin_mul_overflow(x, y, r); }
#endif

CB_HD Val do_arith(Ctx &c, int op, const Val &a, const Val &b) {
    if (a.tag == CB_T_ERR || b.tag == CB_T_ERR) return mk_err();
    const int64_t kMin = (int64_t)0x8000000000000000ull;
    if (a.tag == CB_T_INT && b.tag == CB_T_INT) {
        int64_t x = (int64_t)a.u, y = (int64_t)b.u, r;
        switch (op) {
        case CB_OP_ADD: return add_ovf(x, y, &r) ? mk_err() : mk_int(r);
        case CB_OP_SUB: return sub_ovf(x, y, &r) ? mk_err() : mk_int(r);
        case CB_OP_MUL: return mul_ovf(x, y, &r) ? mk_err() : mk_int(r);
        case CB_OP_DIV: return (y == 0 || (x == kMin && y == -1)) ? mk_err() : mk_int(x / y);
        default: return (y == 0 || (x == kMin && y == -1)) ? mk_err() : mk_int(x % y);
        }
    }
    if (a.tag == CB_T_UINT && b.tag == CB_T_UINT) {
        uint64_t x = a.u, y = b.u, r;
        switch (op) {
        case CB_OP_ADD: r = x + y; return r < x ? mk_err() : mk(CB_T_UINT, r);
        case CB_OP_SUB: return y > x ? mk_err() : mk(CB_T_UINT, x - y);
        case CB_OP_MUL: return umul_ovf(x, y, &r) ? mk_err() : mk(CB_T_UINT, r);
        case CB_OP_DIV: return y == 0 ? mk_err() : mk(CB_T_UINT, x / y);
        default: return y == 0 ? mk_err() : mk(CB_T_UINT, x % y);
        }
    }
    if (a.tag == CB_T_DOUBLE && b.tag == CB_T_DOUBLE) {
        double x = u2d(a.u), y = u2d(b.u);
        switch (op) {
        case CB_OP_ADD: return mk_double(x + y);
        case CB_OP_SUB: return mk_double(x - y);
        case CB_OP_MUL: return mk_double(x * y);
        case CB_OP_DIV: return mk_double(x / y);
        default: return mk_err();
        }
    }
    int64_t x = (int64_t)a.u, y = (int64_t)b.u, r;
    if (op == CB_OP_ADD) {
        if ((a.tag == CB_T_TS && b.tag == CB_T_DUR) || (a.tag == CB_T_DUR && b.tag == CB_T_TS)) {
            if (add_ovf(x, y, &r)) { c.unsupported = 1; return mk_err(); }
            return mk(CB_T_TS, (uint64_t)r);
        }
        if (a.tag == CB_T_DUR && b.tag == CB_T_DUR) return add_ovf(x, y, &r) ? mk_err() : mk(CB_T_DUR, (uint64_t)r);
        if ((a.tag == CB_T_STRING && b.tag == CB_T_STRING) || (a.tag == CB_T_LIST && b.tag == CB_T_LIST)) {
            c.unsupported = 1;  // concatenation would need device-side allocation
            return mk_err();
        }
    }
    if (op == CB_OP_SUB) {
        if (a.tag == CB_T_TS && b.tag == CB_T_TS) return sub_ovf(x, y, &r) ? mk_err() : mk(CB_T_DUR, (uint64_t)r);
        if (a.tag == CB_T_TS && b.tag == CB_T_DUR) {
            if (sub_ovf(x, y, &r)) { c.unsupported = 1; return mk_err(); }
            return mk(CB_T_TS, (uint64_t)r);
        }
        if (a.tag == CB_T_DUR && b.tag == CB_T_DUR) return sub_ovf(x, y, &r) ? mk_err() : mk(CB_T_DUR, (uint64_t)r);
    }
    return mk_err();
}

// ---- string predicates (byte-wise; UTF-8 makes prefix/suffix/substring tests byte-exact) ----
CB_HD bool bytes_eq(const uint8_t *a, const uint8_t *b, uint32_t n) {
    for (uint32_t i = 0; i < n; i++)
        if (ldg(a + i) != ldg(b + i)) return false;
    return true;
}
CB_HD Val do_str2(const Ctx &c, int op, const Val &s, const Val &t) {
    if (s.tag != CB_T_STRING || t.tag != CB_T_STRING) return mk_err();
    const uint8_t *ps, *pt;
    uint32_t ls, lt;
    str_get(c, s.u, ps, ls);
    str_get(c, t.u, pt, lt);
    if (lt > ls) return mk_bool(false);
    if (op == CB_OP_STARTS_WITH) return mk_bool(bytes_eq(ps, pt, lt));
    if (op == CB_OP_ENDS_WITH) return mk_bool(bytes_eq(ps + (ls - lt), pt, lt));
    for (uint32_t i = 0; i + lt <= ls; i++)
        if (bytes_eq(ps + i, pt, lt)) return mk_bool(true);
    return mk_bool(false);
}
CB_HD uint32_t utf8_len(const uint8_t *p, uint32_t n) {
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; i++) k += (ldg(p + i) & 0xC0) != 0x80;
    return k;
}

// ---- RFC 3339 text -> int64 nanoseconds ----
CB_HD int64_t days_from_civil(int64_t y, int m, int d) {
    y -= m <= 2;
    int64_t era = (y >= 0 ? y : y - 399) / 400;
    int64_t yoe = y - era * 400;
    int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}
CB_HD bool digits(const uint8_t *p, int n, int *out) {
    int v = 0;
    for (int i = 0; i < n; i++) {
        uint8_t ch = ldg(p + i);
        if (ch < '0' || ch > '9') return false;
        v = v * 10 + (ch - '0');
    }
    *out = v;
    return true;
}
CB_HD_NOINLINE Val parse_ts(Ctx &c, const Val &s) {
    const uint8_t *p;
    uint32_t n;
    str_get(c, s.u, p, n);
    int y, mo, d, h, mi, se;
    if (n < 20) return mk_err();
    uint8_t tch = ldg(p + 10);
    if (!digits(p, 4, &y) || ldg(p + 4) != '-' || !digits(p + 5, 2, &mo) || ldg(p + 7) != '-' || !digits(p + 8, 2, &d) ||
        (tch != 'T' && tch != 't') || !digits(p + 11, 2, &h) || ldg(p + 13) != ':' || !digits(p + 14, 2, &mi) ||
        ldg(p + 16) != ':' || !digits(p + 17, 2, &se))
        return mk_err();
    uint32_t i = 19;
    int64_t ns = 0;
    uint8_t ch = ldg(p + i);
    if (ch == '.' || ch == ',') {
        i++;
        int k = 0;
        uint32_t st = i;
        while (i < n) {
            uint8_t dch = ldg(p + i);
            if (dch < '0' || dch > '9') break;
            if (k < 9) { ns = ns * 10 + (dch - '0'); k++; }
            i++;
        }
        if (i == st) return mk_err();
        while (k < 9) { ns *= 10; k++; }
    }
    if (i >= n) return mk_err();
    int64_t off = 0;
    ch = ldg(p + i);
    if (ch == 'Z' || ch == 'z') {
        if (i + 1 != n) return mk_err();
    } else if (ch == '+' || ch == '-') {
        int oh, om;
        if (i + 6 != n || !digits(p + i + 1, 2, &oh) || ldg(p + i + 3) != ':' || !digits(p + i + 4, 2, &om) || oh > 23 || om > 59)
            return mk_err();
        off = (int64_t)(oh * 3600 + om * 60) * (ch == '+' ? 1 : -1);
    } else return mk_err();
    bool leap = (y % 4 == 0 && (y % 100 != 0 || y % 400 == 0));
    int dim = (mo == 2) ? (leap ? 29 : 28) : ((mo == 4 || mo == 6 || mo == 9 || mo == 11) ? 30 : 31);
    if (y < 1 || mo < 1 || mo > 12 || d < 1 || d > dim || h > 23 || mi > 59 || se > 59) return mk_err();
    int64_t secs = days_from_civil(y, mo, d) * 86400 + h * 3600 + mi * 60 + se - off;
    int64_t total;
    if (mul_ovf(secs, 1000000000ll, &total) || add_ovf(total, ns, &total)) {
        c.unsupported = 1;  // valid CEL timestamp outside the int64-nanosecond device range
        return mk_err();
    }
    return mk(CB_T_TS, (uint64_t)total);
}

// ---- IP addresses (Go net.ParseIP / IPNet.Contains) ----
CB_HD bool parse_ipv4(const uint8_t *p, uint32_t n, uint32_t *out) {
    uint32_t v = 0, i = 0;
    for (int part = 0; part < 4; part++) {
        uint32_t st = i;
        int x = 0;
        while (i < n) {
            uint8_t ch = ldg(p + i);
            if (ch < '0' || ch > '9') break;
            x = x * 10 + (ch - '0');
            i++;
            if (i - st > 3) return false;
        }
        if (i == st || x > 255 || (i - st > 1 && ldg(p + st) == '0')) return false;
        v = (v << 8) | (uint32_t)x;
        if (part < 3) {
            if (i >= n || ldg(p + i) != '.') return false;
            i++;
        }
    }
    if (i != n) return false;
    *out = v;
    return true;
}
CB_HD int hexv(uint8_t ch) {
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
    if (ch >= 'A' && ch <= 'F') return ch - 'A' + 10;
    return -1;
}
// groups are accumulated into two 64-bit halves to avoid a dynamically indexed local array
CB_HD void ip6_set(uint64_t &hi, uint64_t &lo, int idx, uint32_t v) {
    if (idx < 4) hi |= (uint64_t)v << (48 - 16 * idx);
    else lo |= (uint64_t)v << (48 - 16 * (idx - 4));
}
CB_HD_NOINLINE bool parse_ipv6(const uint8_t *p, uint32_t n, uint64_t *ohi, uint64_t *olo) {
    // pass 1: count groups before/after "::" ; pass 2: place them
    uint64_t hi = 0, lo = 0;
    int ng = 0, ell = -1;
    uint32_t i = 0;
    uint32_t gv[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 8; q++) gv[q] = 0;
    if (n >= 2 && ldg(p) == ':' && ldg(p + 1) == ':') {
        ell = 0;
        i = 2;
    } else if (n >= 1 && ldg(p) == ':') return false;
    while (i < n) {
        uint32_t j = i;
        bool isv4 = false;
        while (j < n && ldg(p + j) != ':') { if (ldg(p + j) == '.') isv4 = true; j++; }
        if (isv4) {
            uint32_t v4;
            if (j != n || ng > 6 || !parse_ipv4(p + i, n - i, &v4)) return false;
            gv[ng++] = v4 >> 16;
            gv[ng++] = v4 & 0xFFFF;
            i = n;
            break;
        }
        if (j == i || j - i > 4 || ng >= 8) return false;
        uint32_t v = 0;
        for (uint32_t k = i; k < j; k++) {
            int h = hexv(ldg(p + k));
            if (h < 0) return false;
            v = v * 16 + (uint32_t)h;
        }
        gv[ng++] = v;
        i = j;
        if (i < n) {
            i++;
            if (i < n && ldg(p + i) == ':') {
                if (ell >= 0) return false;
                ell = ng;
                i++;
            } else if (i == n) return false;
        }
    }
    if (ell >= 0) {
        if (ng >= 8) return false;
        int tail = ng - ell;
        for (int q = 0; q < ell; q++) ip6_set(hi, lo, q, gv[q]);
        for (int q = 0; q < tail; q++) ip6_set(hi, lo, 8 - tail + q, gv[ell + q]);
    } else {
        if (ng != 8) return false;
        for (int q = 0; q < 8; q++) ip6_set(hi, lo, q, gv[q]);
    }
    *ohi = hi;
    *olo = lo;
    return true;
}
CB_HD_NOINLINE Val do_in_ip_range(Ctx &c, const Val &ip, const uint64_t *cidr) {
    if (ip.tag != CB_T_STRING) return mk_err();
    const uint8_t *p;
    uint32_t n;
    str_get(c, ip.u, p, n);
    bool has_colon = false, has_dot = false;
    for (uint32_t i = 0; i < n; i++) {
        uint8_t ch = ldg(p + i);
        if (ch == ':') has_colon = true;
        if (ch == '.') has_dot = true;
        if (ch == '%') return mk_err();
    }
    uint64_t fam = ldg(cidr), bits = ldg(cidr + 1), hi = ldg(cidr + 2), lo = ldg(cidr + 3);
    bool is4 = false;
    uint32_t v4 = 0;
    uint64_t ihi = 0, ilo = 0;
    if (has_dot && !has_colon) {
        if (!parse_ipv4(p, n, &v4)) return mk_err();
        is4 = true;
    } else if (has_colon) {
        if (!parse_ipv6(p, n, &ihi, &ilo)) return mk_err();
        if (ihi == 0 && (ilo >> 32) == 0xFFFF) { is4 = true; v4 = (uint32_t)ilo; }
    } else return mk_err();
    uint64_t nfam = fam, nbits = bits, nlo = lo;
    if (fam == 6 && hi == 0 && (lo >> 32) == 0xFFFF && bits >= 96) { nfam = 4; nbits = bits - 96; nlo = lo & 0xFFFFFFFFull; }
    if (is4) {
        if (nfam != 4) return mk_bool(false);
        uint32_t mask = nbits == 0 ? 0u : (uint32_t)(0xFFFFFFFFull << (32 - nbits));
        return mk_bool((v4 & mask) == ((uint32_t)nlo & mask));
    }
    if (nfam != 6) return mk_bool(false);
    uint64_t mhi = bits >= 64 ? ~0ull : (bits == 0 ? 0ull : (~0ull << (64 - bits)));
    uint64_t mlo = bits <= 64 ? 0ull : (bits == 128 ? ~0ull : (~0ull << (128 - bits)));
    return mk_bool((ihi & mhi) == (hi & mhi) && (ilo & mlo) == (lo & mlo));
}

// ---- conversions ----
CB_HD_NOINLINE Val conv_int(Ctx &c, const Val &v) {
    switch (v.tag) {
    case CB_T_INT: return v;
    case CB_T_UINT: return v.u > 0x7FFFFFFFFFFFFFFFull ? mk_err() : mk_int((int64_t)v.u);
    case CB_T_DOUBLE: {
        double d = u2d(v.u);
        if (d != d || d <= -9223372036854775808.0 || d >= 9223372036854775808.0) return mk_err();
        return mk_int((int64_t)d);
    }
    case CB_T_STRING: {
        const uint8_t *p;
        uint32_t n;
        str_get(c, v.u, p, n);
        uint32_t i = 0;
        bool neg = false;
        if (n) { uint8_t ch = ldg(p); if (ch == '+' || ch == '-') { neg = ch == '-'; i = 1; } }
        if (i == n) return mk_err();
        uint64_t acc = 0;
        for (; i < n; i++) {
            uint8_t ch = ldg(p + i);
            if (ch < '0' || ch > '9') return mk_err();
            if (acc > (0xFFFFFFFFFFFFFFFFull - 9) / 10) return mk_err();
            acc = acc * 10 + (uint64_t)(ch - '0');
        }
        if (neg) { if (acc > 0x8000000000000000ull) return mk_err(); return mk_int((int64_t)(0 - acc)); }
        if (acc > 0x7FFFFFFFFFFFFFFFull) return mk_err();
        return mk_int((int64_t)acc);
    }
    case CB_T_TS: { int64_t ns = (int64_t)v.u; int64_t s = ns / 1000000000; if (ns % 1000000000 < 0) s--; return mk_int(s); }
    case CB_T_DUR: return mk_int((int64_t)v.u);
    default: return mk_err();
    }
}
CB_HD_NOINLINE Val conv_uint(Ctx &c, const Val &v) {
    switch (v.tag) {
    case CB_T_UINT: return v;
    case CB_T_INT: return (int64_t)v.u < 0 ? mk_err() : mk(CB_T_UINT, v.u);
    case CB_T_DOUBLE: {
        double d = u2d(v.u);
        if (d != d || d < 0 || d >= 18446744073709551616.0) return mk_err();
        return mk(CB_T_UINT, (uint64_t)d);
    }
    case CB_T_STRING: {
        const uint8_t *p;
        uint32_t n;
        str_get(c, v.u, p, n);
        uint32_t i = 0;
        if (n && ldg(p) == '+') i = 1;
        if (i == n) return mk_err();
        uint64_t acc = 0;
        for (; i < n; i++) {
            uint8_t ch = ldg(p + i);
            if (ch < '0' || ch > '9') return mk_err();
            uint64_t dg = (uint64_t)(ch - '0');
            if (acc > (0xFFFFFFFFFFFFFFFFull - dg) / 10) return mk_err();
            acc = acc * 10 + dg;
        }
        return mk(CB_T_UINT, acc);
    }
    default: return mk_err();
    }
}

// ---- 3-valued && / || with cel-go error absorption ----
CB_HD Val and_or(bool is_or, const Val &a, const Val &b) {
    bool ab = a.tag == CB_T_BOOL, bb = b.tag == CB_T_BOOL;
    uint64_t dom = is_or ? 1 : 0;
    if (ab && a.u == dom) return a;
    if (bb && b.u == dom) return b;
    if (ab && bb) return mk_bool(!is_or);
    return mk_err();
}

CB_HD Val load_slot(const Ctx &c, uint32_t s, int *state) {
    return decode_v64(ldcol64(c.b->slots + (uint64_t)s * c.b->stride + c.req), state);
}
CB_HD Val load_const(const Ctx &c, uint32_t k) {
    const cb_const *p = c.t->consts + k;
    return mk(ldg(&p->tag), ldg(&p->bits));
}

struct Loop {
    Val range;
    uint64_t i, n;
    uint32_t any_err;
    int64_t count;
};

CB_HD void loop_bind(Ctx &c, const Loop &L, int var, bool two) {
    const uint64_t *p = heap_ptr(c, L.range.u);
    if (L.range.tag == CB_T_LIST) {
        Val e = decode_elem(ldg(p + 1 + L.i));
        if (two) { c.vars[var] = mk_int((int64_t)L.i); c.vars[var + 1] = e; } else c.vars[var] = e;
    } else {
        Val k = decode_elem(ldg(p + 1 + L.i));
        if (two) { c.vars[var] = k; c.vars[var + 1] = decode_elem(ldg(p + 1 + L.n + L.i)); } else c.vars[var] = k;
    }
}

// Runs one condition program; returns true iff it yields BOOL true (ruletable.go:1425-1441).
CB_HD_NOINLINE bool run_program(Ctx &c, const cb_instr *code) {
    Val st[CB_MAX_STACK + 1];
    Loop loops[CB_MAX_LOOP_DEPTH];
    int sp = 0, ld = 0;
    uint32_t pc = 0;
    for (;;) {
        // one 8-byte instruction fetch
        uint64_t raw = ldg(reinterpret_cast<const uint64_t *>(code + pc));
        pc++;
        uint32_t op = (uint32_t)(raw & 0xFF), ia = (uint32_t)((raw >> 8) & 0xFF), ib = (uint32_t)((raw >> 16) & 0xFFFF);
        uint32_t ic = (uint32_t)(raw >> 32);
        switch (op) {
        case CB_OP_RET: return st[sp - 1].tag == CB_T_BOOL && st[sp - 1].u == 1;
        case CB_OP_CONST: st[sp++] = load_const(c, ic); break;
        case CB_OP_SLOT: { int s; st[sp++] = load_slot(c, ic, &s); break; }
        case CB_OP_HAS_SLOT: { int s; load_slot(c, ic, &s); st[sp++] = s == SLOT_ERROR ? mk_err() : mk_bool(s == SLOT_VALUE); break; }
        case CB_OP_PID: st[sp++] = mk(CB_T_STRING, c.pid); break;
        case CB_OP_NOW: st[sp++] = mk(CB_T_TS, (uint64_t)c.b->now); break;
        case CB_OP_VAR: st[sp++] = c.vars[ia]; break;
        case CB_OP_SELECT: { Val m = st[sp - 1]; Val o; st[sp - 1] = (m.tag == CB_T_MAP && map_find(c, m, mk(CB_T_STRING, ic), &o)) ? o : mk_err(); break; }
        case CB_OP_HAS: { Val m = st[sp - 1]; st[sp - 1] = m.tag == CB_T_MAP ? mk_bool(map_find(c, m, mk(CB_T_STRING, ic), nullptr)) : mk_err(); break; }
        case CB_OP_INDEX: sp--; st[sp - 1] = do_index(c, st[sp - 1], st[sp]); break;
        case CB_OP_EQ: case CB_OP_NE: case CB_OP_LT: case CB_OP_LE: case CB_OP_GT: case CB_OP_GE:
            sp--; st[sp - 1] = do_cmp(c, (int)op - CB_OP_EQ, st[sp - 1], st[sp]); break;
        case CB_OP_ADD: case CB_OP_SUB: case CB_OP_MUL: case CB_OP_DIV: case CB_OP_MOD:
            sp--; st[sp - 1] = do_arith(c, (int)op, st[sp - 1], st[sp]); break;
        case CB_OP_NEG: {
            Val v = st[sp - 1];
            const int64_t kMin = (int64_t)0x8000000000000000ull;
            if (v.tag == CB_T_INT) st[sp - 1] = (int64_t)v.u == kMin ? mk_err() : mk_int(-(int64_t)v.u);
            else if (v.tag == CB_T_DOUBLE) st[sp - 1] = mk_double(-u2d(v.u));
            else if (v.tag == CB_T_DUR) st[sp - 1] = (int64_t)v.u == kMin ? mk_err() : mk(CB_T_DUR, (uint64_t)(-(int64_t)v.u));
            else st[sp - 1] = mk_err();
            break;
        }
        case CB_OP_NOT: { Val v = st[sp - 1]; st[sp - 1] = v.tag == CB_T_BOOL ? mk_bool(!v.u) : mk_err(); break; }
        case CB_OP_IN: sp--; st[sp - 1] = do_in(c, st[sp - 1], st[sp]); break;
        case CB_OP_SIZE: {
            Val v = st[sp - 1];
            if (v.tag == CB_T_STRING) { const uint8_t *p; uint32_t n; str_get(c, v.u, p, n); st[sp - 1] = mk_int(utf8_len(p, n)); }
            else if (is_container(v)) st[sp - 1] = mk_int((int64_t)ldg(heap_ptr(c, v.u)));
            else st[sp - 1] = mk_err();
            break;
        }
        case CB_OP_STARTS_WITH: case CB_OP_ENDS_WITH: case CB_OP_CONTAINS:
            sp--; st[sp - 1] = do_str2(c, (int)op, st[sp - 1], st[sp]); break;
        case CB_OP_JF_KEEP: if (st[sp - 1].tag == CB_T_BOOL && st[sp - 1].u == 0) pc = ic; break;
        case CB_OP_JT_KEEP: if (st[sp - 1].tag == CB_T_BOOL && st[sp - 1].u == 1) pc = ic; break;
        case CB_OP_AND: sp--; st[sp - 1] = and_or(false, st[sp - 1], st[sp]); break;
        case CB_OP_OR: sp--; st[sp - 1] = and_or(true, st[sp - 1], st[sp]); break;
        case CB_OP_JMP: pc = ic; break;
        case CB_OP_TERN: {
            Val v = st[--sp];
            if (v.tag == CB_T_BOOL) { if (!v.u) pc = ic; }
            else { st[sp++] = mk_err(); pc = ib; }
            break;
        }
        case CB_OP_HAS_INTERSECTION: sp--; st[sp - 1] = do_set_pred(c, false, st[sp - 1], st[sp]); break;
        case CB_OP_IS_SUBSET: sp--; st[sp - 1] = do_set_pred(c, true, st[sp - 1], st[sp]); break;
        case CB_OP_LOOP_INIT: {
            Val r = st[--sp];
            int kind = (int)(ib & 0xFF);
            bool two = (ib >> 8) & 1;
            if (!is_container(r)) { st[sp++] = mk_err(); pc = ic; break; }
            Loop &L = loops[ld];
            L.range = r; L.i = 0; L.n = ldg(heap_ptr(c, r.u)); L.any_err = 0; L.count = 0;
            if (L.n == 0) { st[sp++] = mk_bool(kind == CB_LOOP_ALL); pc = ic; break; }
            ld++;
            loop_bind(c, L, (int)ia, two);
            break;
        }
        case CB_OP_LOOP_NEXT: {
            Val r = st[--sp];
            int kind = (int)(ib & 0xFF);
            bool two = (ib >> 8) & 1;
            Loop &L = loops[ld - 1];
            bool done = false;
            Val res = mk_err();
            if (kind == CB_LOOP_EXISTS_ONE) {
                if (r.tag != CB_T_BOOL) L.any_err = 1; else if (r.u) L.count++;
            } else {
                uint64_t dom = kind == CB_LOOP_EXISTS ? 1 : 0;
                if (r.tag == CB_T_BOOL) { if (r.u == dom) { done = true; res = mk_bool(dom != 0); } }
                else L.any_err = 1;
            }
            L.i++;
            if (!done && L.i >= L.n) {
                done = true;
                if (L.any_err) res = mk_err();
                else if (kind == CB_LOOP_EXISTS_ONE) res = mk_bool(L.count == 1);
                else res = mk_bool(kind == CB_LOOP_ALL);
            }
            if (done) { ld--; st[sp++] = res; }
            else { loop_bind(c, L, (int)ia, two); pc = ic; }
            break;
        }
        case CB_OP_TO_COND: { Val v = st[sp - 1]; st[sp - 1] = mk_bool(v.tag == CB_T_BOOL && v.u == 1); break; }
        case CB_OP_COND_NOT: st[sp - 1] = mk_bool(!st[sp - 1].u); break;
        case CB_OP_NOERR: st[sp - 1] = mk_bool(st[sp - 1].tag != CB_T_ERR); break;
        case CB_OP_INT: st[sp - 1] = conv_int(c, st[sp - 1]); break;
        case CB_OP_UINT: st[sp - 1] = conv_uint(c, st[sp - 1]); break;
        case CB_OP_DOUBLE: {
            Val v = st[sp - 1];
            if (v.tag == CB_T_INT) st[sp - 1] = mk_double((double)(int64_t)v.u);
            else if (v.tag == CB_T_UINT) st[sp - 1] = mk_double((double)v.u);
            else if (v.tag == CB_T_STRING) { c.unsupported = 1; st[sp - 1] = mk_err(); }  // strconv.ParseFloat at run time
            else if (v.tag != CB_T_DOUBLE) st[sp - 1] = mk_err();
            break;
        }
        case CB_OP_TIMESTAMP: {
            Val v = st[sp - 1];
            if (v.tag == CB_T_TS) break;
            if (v.tag == CB_T_STRING) st[sp - 1] = parse_ts(c, v);
            else if (v.tag == CB_T_INT) {
                int64_t s = (int64_t)v.u, ns;
                if (s < -62135596800ll || s > 253402300799ll) st[sp - 1] = mk_err();
                else if (mul_ovf(s, 1000000000ll, &ns)) { c.unsupported = 1; st[sp - 1] = mk_err(); }
                else st[sp - 1] = mk(CB_T_TS, (uint64_t)ns);
            } else st[sp - 1] = mk_err();
            break;
        }
        case CB_OP_DURATION: {
            Val v = st[sp - 1];
            if (v.tag == CB_T_DUR) break;
            if (v.tag == CB_T_INT) st[sp - 1] = mk(CB_T_DUR, v.u);
            else if (v.tag == CB_T_STRING) { c.unsupported = 1; st[sp - 1] = mk_err(); }
            else st[sp - 1] = mk_err();
            break;
        }
        case CB_OP_DYN: break;
        case CB_OP_CMP_SLOT_CONST: { int s; Val a = load_slot(c, ib, &s); st[sp++] = do_cmp(c, (int)ia, a, load_const(c, ic)); break; }
        case CB_OP_CMP_SLOT_SLOT: { int s; Val a = load_slot(c, ib, &s); Val b = load_slot(c, ic, &s); st[sp++] = do_cmp(c, (int)ia, a, b); break; }
        case CB_OP_CMP_SLOT_PID: { int s; Val a = load_slot(c, ib, &s); st[sp++] = do_cmp(c, (int)ia, a, mk(CB_T_STRING, c.pid)); break; }
        case CB_OP_IN_SLOT_CONST: { int s; Val a = load_slot(c, ib, &s); st[sp++] = do_in(c, a, load_const(c, ic)); break; }
        case CB_OP_IN_CONST_SLOT: { int s; Val a = load_slot(c, ib, &s); st[sp++] = do_in(c, load_const(c, ic), a); break; }
        case CB_OP_IN_IP_RANGE: st[sp - 1] = st[sp - 1].tag == CB_T_ERR ? mk_err() : do_in_ip_range(c, st[sp - 1], c.t->theap + ic); break;
        default: c.unsupported = 1; return false;
        }
    }
}

// ---------------------------------------------------------------------------------------------- decision walk
CB_HD bool in_class(const BatchView &b, uint32_t c0, uint32_t c1, uint32_t pat) {
    for (uint32_t j = c0; j < c1; j++)
        if (ldg(b.class_pats + j) == pat) return true;
    return false;
}

// is table role `role` in {req_role} U parents(exact resource scope, req_role)   (index.go:805-836)
CB_HD bool role_in_pr(const TableView &t, uint32_t role, uint32_t req_role, uint32_t rscope) {
    if (req_role == role) return true;
    if (!t.has_parent_roles || req_role >= t.nR) return false;
    if (rscope == CB_SCOPE_NONE || (rscope & CB_SCOPE_INEXACT_BIT) || rscope >= t.nS) return false;
    uint64_t idx = (uint64_t)rscope * t.nR + req_role;
    for (uint32_t j = ldg(t.par_off + idx), e = ldg(t.par_off + idx + 1); j < e; j++)
        if (ldg(t.par_list + j) == role) return true;
    return false;
}

struct Memo {
    uint64_t done, val;
};

CB_HD bool cond_sat(Ctx &c, Memo &m, uint32_t cond_base, uint32_t local /*1-based*/) {
    uint32_t li = local - 1;
    if (li < 64 && ((m.done >> li) & 1)) return (m.val >> li) & 1;
    const cb_cond *cd = c.t->conds + (cond_base + li);
    bool s = run_program(c, c.t->code + ldg(&cd->code_off));
    if (li < 64) { m.done |= 1ull << li; m.val |= (uint64_t)s << li; }
    return s;
}

// first scope of the chain for `kind_flag`, honouring strict / lenient search (ruletable.go:626-632)
CB_HD uint32_t chain_start(const TableView &t, uint32_t scope, uint32_t kind_flag, bool lenient) {
    if (scope == CB_SCOPE_NONE) return CB_NONE32;
    uint32_t s = scope & ~CB_SCOPE_INEXACT_BIT;
    if (s >= t.nS) return CB_NONE32;
    if (!(ldg(t.scope_flags + s) & kind_flag)) {
        if (!lenient) return CB_NONE32;
        do { s = ldg(t.scope_parent + s); } while (s != CB_NONE32 && !(ldg(t.scope_flags + s) & kind_flag));
    }
    return s;
}
CB_HD uint32_t chain_next(const TableView &t, uint32_t s, uint32_t kind_flag) {
    do { s = ldg(t.scope_parent + s); } while (s != CB_NONE32 && !(ldg(t.scope_flags + s) & kind_flag));
    return s;
}

// Evaluates request `n` (absolute column index). Writes kbytes bytes of the packed ALLOW bitmap.
CB_HD void eval_request(const TableView &t, const BatchView &b, uint64_t n, uint8_t *bitmap, uint32_t *status) {
    Ctx c;
    c.t = &t; c.b = &b; c.req = n; c.unsupported = 0;
    uint8_t *out = bitmap + n * b.kbytes;
    cb_hdr0 h0 = load_hdr0(b.hdr0 + n);
    cb_hdr1 h1 = load_hdr1(b.hdr1 + n);
    c.pid = h0.principal_id;
    // result bits: actions 0..63 accumulate in a register and are stored once; wider action lists
    // (K > 64, rare) fall back to read-modify-write on the thread's own output bytes
    uint64_t acc = 0;
    const bool wide = b.kbytes > 8;
    if (wide) for (uint32_t q = 0; q < b.kbytes; q++) out[q] = 0;
    struct Store {
        uint8_t *out; uint32_t kbytes; bool wide; const uint64_t *acc;
        CB_HD ~Store() { if (!wide) for (uint32_t q = 0; q < kbytes; q++) out[q] = (uint8_t)(*acc >> (8 * q)); }
    } store_on_exit{out, b.kbytes, wide, &acc};

    uint32_t roles[CB_MAX_ROLE_COLS];
    uint32_t n_roles = 0;
    for (uint32_t i = 0; i < b.role_cols; i++) {
        uint32_t rr = ldcol32(b.roles + (uint64_t)i * b.stride + n);
        roles[i] = rr;
        if (rr != CB_ROLE_PAD) n_roles = i + 1;   // encoder packs roles to the front
    }
    uint32_t aset = h1.action_set_id;
    uint32_t K = aset < b.n_asets ? ldg(b.aset_k + aset) : 0;
    if (n_roles == 0 || K == 0) return;

    bool lenient = (b.flags & CB_BATCH_FLAG_LENIENT) != 0;
    uint32_t p0 = chain_start(t, h0.principal_scope, CB_SCOPE_FLAG_PRINCIPAL, lenient);
    uint32_t r0 = chain_start(t, h0.resource_scope, CB_SCOPE_FLAG_RESOURCE, lenient);
    if (p0 == CB_NONE32 && r0 == CB_NONE32) return;
    uint32_t rv = h1.resource_version, pv = h1.principal_version;
    uint32_t cls0 = ldg(b.class_off + h0.kind_class), cls1 = ldg(b.class_off + h0.kind_class + 1);

    // existence checks (ruletable.go:852-863)
    bool p_exists = false, r_exists = false;
    if (pv != CB_NONE16)
        for (uint32_t s = p0; s != CB_NONE32; s = chain_next(t, s, CB_SCOPE_FLAG_PRINCIPAL))
            p_exists |= ldg(t.prin_exists + (uint64_t)pv * t.nS + s) != 0;
    if (rv != CB_NONE16)
        for (uint32_t s = r0; s != CB_NONE32; s = chain_next(t, s, CB_SCOPE_FLAG_RESOURCE))
            for (uint32_t j = cls0; j < cls1; j++)
                r_exists |= (ldg(t.res_exists + ((uint64_t)rv * t.nRP + ldg(b.class_pats + j)) * t.nS + s) & CB_EXISTS_RESOURCE_KIND) != 0;
    if ((!p_exists && !r_exists) || rv == CB_NONE16) return;

    uint32_t pidx = (t.has_principal_policies && h0.principal_id < t.nT) ? ldg(t.prin_of_string + h0.principal_id) : CB_NONE32;
    const uint32_t RC = b.role_cols;
    const uint64_t role_all = (n_roles >= 64) ? ~0ull : ((1ull << n_roles) - 1);
    const uint32_t nAP = t.nAP ? t.nAP : 1;

    for (uint32_t ps = 0; ps < b.n_pass; ps++) {
        uint32_t kbase = ps * b.kc;
        if (kbase >= K) break;
        uint32_t kn = K - kbase < b.kc ? K - kbase : b.kc;            // actions in this pass
        const uint64_t *spread = b.aset_spread + ((uint64_t)ps * b.n_asets + aset) * nAP;
        // action-only mask: bit kk*RC for every action of this pass
        uint64_t amask = 0;
        for (uint32_t kk = 0; kk < kn; kk++) amask |= 1ull << (kk * RC);

        // ---- principal policies: role agnostic, decided per action (state lives on role column 0) ----
        uint64_t p_allow = 0, p_deny = 0;
        if (pidx != CB_NONE32) {
            uint64_t alive = amask;
            for (uint32_t s = p0; s != CB_NONE32 && alive; s = chain_next(t, s, CB_SCOPE_FLAG_PRINCIPAL)) {
                uint32_t bid = ldg(t.prin_block_map + ((uint64_t)rv * t.nP + pidx) * t.nS + s);
                if (bid == CB_NONE32) continue;
                cb_block bl = load_block(t.blocks + bid);
                Memo memo; memo.done = 0; memo.val = 0;
                uint64_t D = 0, A = 0;
                for (uint32_t ri = 0; ri < bl.n_rows; ri++) {
                    cb_row row = load_row(t.rows + bl.row_start + ri);
                    uint64_t m = ldg(spread + row.apat) & alive;
                    if (!m) continue;
                    if (!in_class(b, cls0, cls1, row.respat)) continue;
                    if (row.effect == CB_EFFECT_DENY ? (m & ~D) == 0 : (m & ~A) == 0) continue;   // nothing new to learn
                    if (row.drcond && !cond_sat(c, memo, bl.cond_base, row.drcond)) continue;
                    if (row.cond && !cond_sat(c, memo, bl.cond_base, row.cond)) continue;
                    if (row.effect == CB_EFFECT_DENY) D |= m; else A |= m;
                }
                p_deny |= D;
                alive &= ~D;
                uint32_t perm = (ldg(t.scope_flags + s) >> CB_SCOPE_PERM_SHIFT) & 3;
                if (perm == 1) { uint64_t a = A & alive; p_allow |= a; alive &= ~a; }
            }
        }

        // ---- resource policies: (action x role) pairs walk the chain together ----
        uint64_t undecided = amask & ~(p_allow | p_deny);
        uint64_t r_allow_pairs = 0;
        if (undecided && r0 != CB_NONE32) {
            uint64_t alive = undecided * role_all;       // every role column of every undecided action
            for (uint32_t s = r0; s != CB_NONE32 && alive; s = chain_next(t, s, CB_SCOPE_FLAG_RESOURCE)) {
                uint64_t D = 0, A = 0;
                bool any_row = false;
                for (uint32_t j = cls0; j < cls1; j++) {
                    uint64_t mi = ((uint64_t)rv * t.nRP + ldg(b.class_pats + j)) * t.nS + s;
                    any_row |= (ldg(t.res_exists + mi) & CB_EXISTS_ANY_ROW) != 0;
                    uint32_t bid = ldg(t.res_block_map + mi);
                    if (bid == CB_NONE32) continue;
                    cb_block bl = load_block(t.blocks + bid);
                    Memo memo; memo.done = 0; memo.val = 0;
                    for (uint32_t ri = 0; ri < bl.n_rows; ri++) {
                        cb_row row = load_row(t.rows + bl.row_start + ri);
                        uint64_t am = ldg(spread + row.apat);
                        if (!am) continue;
                        uint64_t rmask;
                        if (row.role == CB_ROLE_ANY) rmask = role_all;
                        else {
                            rmask = 0;
                            for (uint32_t i = 0; i < n_roles; i++)
                                rmask |= (uint64_t)role_in_pr(t, row.role, roles[i], h0.resource_scope) << i;
                        }
                        uint64_t m = (am * rmask) & alive;
                        if (!m) continue;
                        if (row.effect == CB_EFFECT_DENY ? (m & ~D) == 0 : (m & ~A) == 0) continue;
                        if (row.drcond && !cond_sat(c, memo, bl.cond_base, row.drcond)) continue;
                        if (row.cond && !cond_sat(c, memo, bl.cond_base, row.cond)) continue;
                        if (row.effect == CB_EFFECT_DENY) D |= m; else A |= m;
                    }
                }
                // synthesized role-policy DENY rows (index.go:688-776)
                if (t.has_role_policies && any_row) {
                    uint64_t ro = (uint64_t)rv * t.nS + s;
                    for (uint32_t e = ldg(t.rp_off + ro), ee = ldg(t.rp_off + ro + 1); e < ee; e++) {
                        cb_rolepol_entry en = load_rp_entry(t.rp_entries + e);
                        uint64_t rmask = 0;
                        for (uint32_t i = 0; i < n_roles; i++)
                            rmask |= (uint64_t)role_in_pr(t, en.role, roles[i], h0.resource_scope) << i;
                        if (!rmask) continue;
                        uint64_t matched = 0;   // action bits (role column 0) with at least one matching allow rule
                        for (uint32_t q = 0; q < en.n_rules; q++) {
                            cb_rolepol_rule ru = load_rp_rule(t.rp_rules + en.rule_start + q);
                            if (!in_class(b, cls0, cls1, ru.respat)) continue;
                            uint64_t am = 0;
                            for (uint32_t a = 0; a < ru.n_apats; a++) am |= ldg(spread + ldg(t.rp_apats + ru.apat_start + a));
                            if (!am) continue;
                            matched |= am;
                            if (ru.cond && ((am * rmask) & alive & ~D)) {
                                Memo none; none.done = 0; none.val = 0;
                                if (!cond_sat(c, none, ru.cond - 1, 1)) D |= (am * rmask) & alive;   // DENY none(cond)
                            }
                        }
                        D |= ((amask & ~matched) * rmask) & alive;   // blanket DENY for actions no allow rule matches
                    }
                }
                alive &= ~D;
                uint32_t perm = (ldg(t.scope_flags + s) >> CB_SCOPE_PERM_SHIFT) & 3;
                if (perm == 1) { uint64_t a = A & alive; r_allow_pairs |= a; alive &= ~a; }
            }
        }

        // ---- fold: ALLOW iff principal walk allowed, or undecided there and any role column allowed ----
        for (uint32_t kk = 0; kk < kn; kk++) {
            uint64_t abit = 1ull << (kk * RC);
            bool allow = (p_allow & abit) != 0;
            if (!allow && (undecided & abit)) allow = ((r_allow_pairs >> (kk * RC)) & role_all) != 0;
            uint32_t k = kbase + kk;
            if (allow) { if (wide) out[k >> 3] |= (uint8_t)(1u << (k & 7)); else acc |= 1ull << k; }
        }
    }
    if (c.unsupported && status) {
#if defined(__CUDA_ARCH__)
        atomicOr(status, 1u);
#else
        *status |= 1u;
#endif
    }
}

}  // namespace cb
