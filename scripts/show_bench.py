#!/usr/bin/env python
"""Print the key numbers of bench.py JSON lines (files given on the command line)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        mv = d.get("matvec") or {}
        dn = (d.get("periphery_dense") or {}).get("roofline") or {}
        print(f"{f}: N={d['n_gpus']} value {d['value']:.4g} {d['unit']}  e2e {d['e2e']['value']:.4g}  "
              f"{d['ms_per_step']:.3f} ms/step  acc {d['accuracy']['max_rel_err_vs_oracle']:.2g}  "
              f"roofline frac {d['roofline']['frac']:.3f}  matvec {mv.get('ms')} ms  dense frac {dn.get('frac')}  "
              f"launches {d['gpu_launches']}  clocks {d['clocks']}")
        if "cpu_baseline" in d:
            print("   cpu_baseline:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "cores")
        if "error" in d:
            print("   ERROR:", d["error"])
    except Exception as e:  # noqa: BLE001
        print(f"{f}: cannot parse ({e}); tail: {open(f).read()[-1200:]}")
