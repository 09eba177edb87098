#!/usr/bin/env python
"""Exports a synthetic workload for the Go CPU-baseline harness (baseline/go/engine_gpubaseline_test.go):

    python tools/export_workload.py --workload C3 --out /tmp/c3 [--requests 1048576] [--want]

writes OUT/policies/*.yaml (one policy document per file, disk-store layout), OUT/inputs.jsonl (protojson
enginev1.CheckInput per line: the first --requests requests of the workload's stream) and, with --want, OUT/want.jsonl
(per input {action: "EFFECT_ALLOW" | "EFFECT_DENY"} as the C port of the algorithm answers them -- the same answers the GPU
path is verified against -- for the harness's parity spot-check)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--out", required=True)
    ap.add_argument("--requests", type=int, default=1 << 20)
    ap.add_argument("--want", action="store_true")
    args = ap.parse_args()
    import yaml
    import workloads as W
    w = W.WORKLOADS[args.workload]()
    os.makedirs(os.path.join(args.out, "policies"), exist_ok=True)
    for i, doc in enumerate(w.policies()):
        kind = next(k for k in ("resourcePolicy", "principalPolicy", "rolePolicy", "derivedRoles", "exportVariables", "exportConstants") if k in doc)
        with open(os.path.join(args.out, "policies", f"{i:05d}_{kind}.yaml"), "w") as f:
            yaml.safe_dump(doc, f, sort_keys=False)
    n = min(args.requests, w.default_n)
    fields = w.fields(n)
    with open(os.path.join(args.out, "inputs.jsonl"), "w") as f:
        for s0 in range(0, n, 65536):
            for inp in w.inputs(fields, range(s0, min(n, s0 + 65536))):
                f.write(json.dumps(inp, separators=(",", ":")) + "\n")
    if args.want:
        from oracle import cref
        _, ft, enc = W.build(w)
        m = min(n, 4096)
        inputs = w.inputs(fields, range(m))
        b = enc.encode(inputs)
        eff = cref.check(ft.blob, b.columns, b.n, b.max_actions, n_threads=os.cpu_count() or 1)
        names = {1: "EFFECT_ALLOW", 2: "EFFECT_DENY"}
        with open(os.path.join(args.out, "want.jsonl"), "w") as f:
            for i, inp in enumerate(inputs):
                f.write(json.dumps({a: names[int(eff[i, k])] for k, a in enumerate(inp["actions"])}) + "\n")
    print(f"wrote {len(w.policies())} policies and {n} inputs to {args.out}")


if __name__ == "__main__":
    main()
