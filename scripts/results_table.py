#!/usr/bin/env python
"""Markdown summary of bench.py JSON lines (profiles/r2_bench_*.json) for DESIGN.md / BASELINE.md."""
import json
import sys


def load(path):
    lines = [l for l in open(path) if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main(paths):
    rows = []
    for p in paths:
        d = load(p)
        if not d:
            continue
        r = d.get("roofline") or {}
        acc = d.get("accuracy") or {}
        rows.append((d["n_gpus"], d["config"]["workload"].split(":")[0], d["config"]["n_nodes"], d["value"] / 1e9,
                     d["ms_per_matvec"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_matvec"], r.get("pairs_per_s"),
                     r.get("frac"), acc.get("max_rel_err_vs_oracle_all_ranks", acc.get("max_rel_err_vs_oracle")),
                     d.get("launches_per_matvec_per_rank"), p, d))
    print("| N | workload | nodes | Gpairs/s (device) | ms / matvec | Gpairs/s (e2e) | ms / matvec e2e | sym kernel Gpairs/s | "
          "frac of nominal FP64 | max rel err | launches / matvec / rank | file |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in sorted(rows, key=lambda x: (x[1], x[0])):
        sym = f"{r[7] / 1e9:.0f}" if r[7] else "-"
        fr = f"{r[8]:.3f}" if r[8] else "-"
        err = f"{r[9]:.1e}" if r[9] is not None else "-"
        print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.0f} | {r[4]:.3f} | {r[5]:.0f} | {r[6]:.3f} | {sym} | {fr} | {err} | "
              f"{r[10]:.0f} | `{r[11]}` |")
    for r in sorted(rows, key=lambda x: (x[1], x[0])):
        d = r[12]
        for k in ("matvec_strong_c3", "multi_device_inprocess", "solve", "cpu_baseline", "stokeslet_call",
                  "ref_gpu_baseline", "fiber_operators", "periphery_dense", "error"):
            if d.get(k):
                v = d[k]
                if isinstance(v, dict):
                    v = {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()
                         if not isinstance(b, (dict, str)) or a in ("error", "unavailable")}
                print(f"- N={r[0]} {r[1]} `{k}`: {json.dumps(v)}")


if __name__ == "__main__":
    main(sys.argv[1:])
