#!/usr/bin/env python
"""Markdown rows for BASELINE.md section 4 from bench.py JSON lines.  Usage: tools/baseline_rows.py file.json ..."""
import json
import os
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as ex:  # noqa: BLE001
        print(f"| {os.path.basename(path)} | unreadable: {ex} |")
        continue
    name = os.path.basename(path)[:-5]
    if d.get("impl") == "reference":
        cb = d.get("cpu_baseline", {})
        print(f"| {name} | reference arm (CPU oracle, {cb.get('cores')} threads) | — | {d['value']:.3e} | | | | {cb.get('sample', '')[:60]} |")
        continue
    r = d["roofline"]
    cb = d.get("cpu_baseline") or {}
    cpu = f"{cb.get('single_thread_value', float('nan')):.3e} / {cb.get('value', float('nan')):.3e} ({cb.get('cores')} cores)" if cb else ""
    print(f"| {name} | {d['config']['robots_per_gpu']} | {d['n_gpus']} | {d['value']:.3e} | {r['algorithmic_bytes_per_step']} B | "
          f"{r['frac']:.3f} ({d['ms_per_step'] * 1e3:.1f} us/step, {d['dtype']}) | e2e {d['e2e']['value']:.3e} | {cpu} |")
