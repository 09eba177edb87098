#!/usr/bin/env python3
"""Extracts the reference's golden TEST VECTORS for the CheckResources hot path
into compact JSON fixtures under tests/golden/.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_golden.py

Sources (all test data, no reference source code is copied):
  internal/test/testdata/cel_eval/*.yaml               -> cel_eval.json
      runner: internal/engine/evaluator_test.go:24-50 (now = 2021-04-22T10:05:20.021-05:00)
  internal/conditions/cerbos_lib_test.go:26-134 (expression table) -> cerbos_lib_test.json
  internal/test/testdata/engine/*.yaml,
  .../engine_strict_scope_search/*.yaml,
  .../engine_lenient_scope_search/*.yaml               -> engine_cases.json
      runner: internal/engine/engine_test.go:50-234 (globals {"environment":"test"},
      default version "default", default scope "")
  internal/test/testdata/store/**  (policies the engine goldens run against) -> store_policies.json
      loader skip rules: internal/util/filesystem.go:21-66,139-166,
      internal/storage/index/builder.go:105-144
  internal/test/testdata/server/checks/check_resources/*.yaml -> check_resources_cases.json
      (API-level goldens incl. JWT claims; JWTs are decoded here without verification)
"""
from __future__ import annotations

import base64
import json
import os
import re
import sys

import yaml

REF = "/root/reference/internal"
TD = f"{REF}/test/testdata"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_yaml_docs(path):
    with open(path, encoding="utf-8") as f:
        text = f.read()
    docs = [d for d in yaml.safe_load_all(text) if d is not None]
    return docs


def dump(name, obj):
    p = os.path.join(OUT, name)
    with open(p, "w", encoding="utf-8") as f:
        json.dump(obj, f, indent=1, sort_keys=True, ensure_ascii=False)
        f.write("\n")
    print(f"wrote {p}: {len(obj) if hasattr(obj, '__len__') else ''}")


def cel_eval():
    out = []
    d = f"{TD}/cel_eval"
    for fn in sorted(os.listdir(d)):
        if not fn.endswith(".yaml"):
            continue
        doc = load_yaml_docs(f"{d}/{fn}")[0]
        out.append({"file": fn, "condition": doc["condition"], "request": doc["request"],
                    "want": bool(doc.get("want", False)), "wantError": bool(doc.get("wantError", False))})
    dump("cel_eval.json", out)


def cerbos_lib_table():
    src = open(f"{REF}/conditions/cerbos_lib_test.go", encoding="utf-8").read()
    start = src.index("func TestCerbosLib")
    end = src.index("env, err := cel.NewEnv", start)
    body = src[start:end]
    out = []
    for m in re.finditer(r"\{expr: `([^`]*)`(, wantErr: true)?\}", body):
        out.append({"expr": m.group(1), "wantErr": bool(m.group(2))})
    dump("cerbos_lib_test.json", out)


def engine_cases():
    out = []
    for sub, lenient in (("engine", False), ("engine_strict_scope_search", False),
                         ("engine_lenient_scope_search", True)):
        d = f"{TD}/{sub}"
        for fn in sorted(os.listdir(d)):
            if not fn.endswith((".yaml", ".yml", ".json")):
                continue
            doc = load_yaml_docs(f"{d}/{fn}")[0]
            out.append({"suite": sub, "file": fn, "lenient": lenient,
                        "description": doc.get("description", ""),
                        "inputs": doc.get("inputs", []), "wantOutputs": doc.get("wantOutputs", []),
                        "wantError": bool(doc.get("wantError", False))})
    dump("engine_cases.json", out)


_SUPPORTED_EXT = (".yaml", ".yml", ".json")


def store_policies():
    root = f"{TD}/store"
    out = []
    for dirpath, dirnames, filenames in os.walk(root):
        # skip rules: hidden dirs, _schemas, testdata
        dirnames[:] = sorted(x for x in dirnames if not x.startswith(".") and x not in ("_schemas", "testdata"))
        for fn in sorted(filenames):
            if fn.startswith(".") or not fn.endswith(_SUPPORTED_EXT):
                continue
            stem = fn.rsplit(".", 1)[0]
            if stem.endswith("_test"):
                continue
            rel = os.path.relpath(os.path.join(dirpath, fn), root)
            for doc in load_yaml_docs(os.path.join(dirpath, fn)):
                if not isinstance(doc, dict) or "apiVersion" not in doc:
                    continue
                out.append({"path": rel, "policy": doc})
    dump("store_policies.json", out)


def _jwt_claims(token: str):
    parts = token.split(".")
    if len(parts) != 3:
        return None
    pad = "=" * (-len(parts[1]) % 4)
    try:
        return json.loads(base64.urlsafe_b64decode(parts[1] + pad))
    except Exception:
        return None


def check_resources_cases():
    d = f"{TD}/server/checks/check_resources"
    out = []
    for fn in sorted(os.listdir(d)):
        if not fn.endswith(".yaml"):
            continue
        doc = load_yaml_docs(f"{d}/{fn}")[0]
        cr = doc.get("checkResources")
        if not cr:
            continue
        inp = cr.get("input", {})
        claims = None
        tok = ((inp.get("auxData") or {}).get("jwt") or {}).get("token")
        if tok:
            claims = _jwt_claims(tok)
        out.append({"file": fn, "description": doc.get("description", ""), "wantStatus": doc.get("wantStatus"),
                    "wantError": bool(doc.get("wantError", False)),
                    "input": inp, "jwtClaims": claims, "wantResponse": cr.get("wantResponse")})
    dump("check_resources_cases.json", out)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (authoring container only)")
    cel_eval()
    cerbos_lib_table()
    engine_cases()
    store_policies()
    check_resources_cases()
