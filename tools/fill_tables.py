"""Writes the measured tables of DESIGN.md §6 and BASELINE.md §5 from the bench lines kept under profiles/
(profiles/r02_bench_n{1,2,4,8}.json = the JSON line `bench.py --gpus N` printed on the GPU box).
usage: python tools/fill_tables.py"""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = 6543.7


def load(n):
    p = os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p) if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


runs = {n: load(n) for n in (1, 2, 4, 8)}
runs = {n: d for n, d in runs.items() if d}
base = runs.get(1)
cpu = base.get("cpu_baseline", {}).get("value") if base else None

meas = ["| N GPUs | ms/step (merge inside) | rows/s | GB/s per GPU | frac of measured | scaling | kernels back to back, no merge | e2e rows/s (host buffers) | e2e PCIe roofline (achieved / plain memcpy, GB/s) |",
        "|---|---|---|---|---|---|---|---|---|"]
for n, d in sorted(runs.items()):
    r, e = d["roofline"], d["e2e"]
    nm = d.get("throughput_no_merge")
    meas.append(f"| {n} | {d['ms_per_step']:.3f} | {d['value']:.3g} | {r['achieved']:.0f} | {r['frac']:.3f} | "
                f"{(d['value'] / base['value']) if base else float('nan'):.2f} | {('%.3f ms' % nm['ms_per_step']) if nm else '—'} | {e['value']:.3g} | "
                f"{e['roofline']['achieved']:.1f} / {e['roofline']['peak']:.1f} |")
if cpu:
    meas.append("")
    meas.append(f"CPU baseline (oracle port of the reference pipeline, 8 partitions on 8 threads, on the GPU box's host): {cpu:.3g} rows/s "
                f"({base['cpu_baseline']['sample'].split(';')[-1].strip()}); best-case fused CPU pass on all cores: "
                f"{base['cpu_baseline'].get('best_case_fused', {}).get('value', float('nan')):.3g} rows/s.")

pq = ["| Query over `numbers_mt(10^10)` | " + " | ".join(f"N={n} mat ms (GB/s per GPU) / gen ms" for n in sorted(runs)) + " | README FuseQuery (8 vCPU) | bit-exact |", "|---|" + "---|" * (len(runs) + 2)]
names = [k for k in (base or {}).get("per_query", {})]
for name in names:
    cells = []
    for n, d in sorted(runs.items()):
        q = d["per_query"].get(name, {})
        m, g = q.get("materialised"), q.get("generated")
        c = ""
        if m:
            c += f"{m['ms']:.3f}" + (f" ({m['gb_per_s'] / n:.0f})" if m.get("gb_per_s") else " (no bytes)")
        if g:
            c += f" / {g['ms']:.3f}"
        cells.append(c or "—")
    q1 = base["per_query"][name]
    any_mode = q1.get("materialised") or q1.get("generated")
    readme = any_mode.get("readme_seconds_8vcpu")
    ok = all((d["per_query"].get(name, {}).get(m) or {"verified": True})["verified"] for d in runs.values() for m in ("materialised", "generated"))
    pq.append(f"| `{name}` | " + " | ".join(cells) + f" | {('%.2f s' % readme) if readme else '—'} | {'Y (asserted on every rank)' if ok else 'N'} |")

gb = ["| GROUP BY `number % k` over 10⁹ rows: sum, count, min, max | " + " | ".join(f"N={n} ms (G rows/s)" for n in sorted(runs)) + " | verified |", "|---|" + "---|" * (len(runs) + 1)]
for key in (base or {}).get("group_by", {}) or {}:
    cells = []
    for n, d in sorted(runs.items()):
        g = (d.get("group_by") or {}).get(key)
        cells.append(f"{g['ms']:.2f} ({g['rows_per_s'] / 1e9:.1f})" if g else "—")
    gb.append(f"| k = {key.split('%')[1].strip()} | " + " | ".join(cells) + " | Y |")

ob = ["| ORDER BY over 10⁹ rows, 1 GPU | digit passes | sort ms (G rows/s) | GB/s over 36 B per row and pass | gather of one 8-byte column, ms | `… LIMIT 10` (radix select), ms | verified |", "|---|---|---|---|---|---|---|"]
for key, o in ((base or {}).get("order_by") or {}).items():
    ob.append(f"| `{key}` | {o['passes']} | {o['sort_ms']:.1f} ({o['rows_per_s'] / 1e9:.2f}) | {o['gb_per_s']:.0f} | {o['take_ms']:.1f} | {o.get('limit_10_ms', float('nan')):.2f} | Y |")

out = {"MEASUREMENT_TABLE": "\n".join(meas), "PER_QUERY_TABLE": "\n".join(pq), "GROUP_BY_TABLE": "\n".join(gb), "ORDER_BY_TABLE": "\n".join(ob)}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_tables.json"), "w"), indent=1)
for fn in ("DESIGN.md", "BASELINE.md"):
    p = os.path.join(ROOT, fn)
    s = open(p).read()
    for k, v in out.items():
        s = re.sub(rf"<!-- {k} -->.*?<!-- /{k} -->", f"<!-- {k} -->\n{v}\n<!-- /{k} -->", s, flags=re.S)
        s = s.replace(f"\n{k}\n", f"\n<!-- {k} -->\n{v}\n<!-- /{k} -->\n")
    open(p, "w").write(s)
print("\n\n".join(out.values()))
