"""Prints one compact line per bench JSON log given on the command line."""
import json
import sys

for path in sys.argv[1:]:
    try:
        for line in open(path):
            if line.startswith("{"):
                j = json.loads(line)
                r = j.get("roofline", {})
                c = j.get("config", {})
                print(f"{path.split('/')[-1]:28s} {j['value']:>12.1f} q/s  step {j['ms_per_step']:.4f} ms | kernel {r.get('kernel_ms')} ms "
                      f"{r.get('achieved')} {r.get('unit')} frac {r.get('frac')} ({r.get('bound')}) pipeline {r.get('pipeline_ms')} ms | "
                      f"e2e {j.get('e2e', {}).get('value')} | algo {c.get('algo')} fb {c.get('fallback_queries')} ovf {c.get('overflow_queries')} "
                      f"| launches {j.get('gpu_launches')} clocks {j.get('clocks', {}).get('sm_mhz') if j.get('clocks') else None}")
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
