#!/usr/bin/env python
"""Tuning experiments on the GPU box: run bench.py once per (library variant, pipeline) and print one compact line
each (ms/step, e2e, per-phase ms).  Full JSON lines go to gpurun_out/ab.jsonl.
usage: python tools/ab_bench.py [--steps K] name[:pipeline] ...   (name 'cur' = the in-tree library)"""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = sys.argv[1:]
    steps = "10"
    if args and args[0] == "--steps":
        steps = args[1]
        args = args[2:]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "ab.jsonl"), "a")
    for spec in args:
        name, _, pipe = spec.partition(":")
        env = dict(os.environ)
        if name != "cur":
            env["CLSN_LIB"] = os.path.join(ROOT, "collision_b200", "variants", f"libclsn_{name}.so")
        if pipe:
            env["CLSN_PIPELINE"] = pipe
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "3", "--no-cpu", "--no-api-default"],
                               env=env, capture_output=True, text=True, timeout=240)
            line = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            print(f"{spec:16s} FAILED: {e}", flush=True)
            continue
        line["variant"] = spec
        out.write(json.dumps(line) + "\n")
        out.flush()
        k = line.get("kernels", {})
        ph = " ".join(f"{n}={k[n]['ms']:.2f}" for n in ("refit", "traverse", "cull", "roots", "contact", "reduce") if n in k)
        print(f"{spec:16s} {line['ms_per_step']:.2f} ms  e2e {line['e2e']['ms_per_step']:.2f}  {ph}  units={line.get('units')}", flush=True)


if __name__ == "__main__":
    main()
