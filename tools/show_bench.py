"""Print the numbers of a bench.py JSON line (and of its `also` records) in a few readable rows."""
import json
import sys


def row(d, tag=""):
    r, e, c = d.get("roofline") or {}, d.get("e2e") or {}, d.get("cpu_baseline") or {}
    print(f"{tag}{d.get('metric')}: value {d.get('value', 0):.4g} {d.get('unit')}  ms/step {d.get('ms_per_step', 0):.3f}  "
          f"frac {r.get('frac', 0):.3f}  e2e {e.get('value', 0):.4g} (wall {e.get('wall_s', 0):.2f} s, ndcg {e.get('ndcg_at_10')})  "
          f"cpu {c.get('value', 0):.4g} [{c.get('kind')}, {c.get('cores')} cores]  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
    if d.get("e2e_error") or d.get("error"):
        print("   ERROR:", d.get("e2e_error") or d.get("error"))


for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as ex:
        print(path, "FAILED", ex)
        continue
    row(d, f"{path} n_gpus={d.get('n_gpus')} ")
    for a in d.get("also", []):
        row(a, "   also ")
