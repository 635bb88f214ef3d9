"""DESIGN.md section 6 table from profiles/r02_bench_n{1,2,4,8}.json (+ the reference arm's lines)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = {}
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")
    if os.path.exists(p):
        d[n] = json.load(open(p))
hdr = ("| GPUs | one query, min_match 2 (19 k hits): cold µs (back to back) | G pairs/s | vs 1 GPU | min_match 5: cold µs (b2b) | "
       "8 queries per pass: µs per pass → G pairs/s | weak, 1 M rows/GPU: cold µs → G pairs/s | fragment, 100 k rows: µs → G pairs/s | "
       "`find_duplicates` e2e (min_match 5) µs | scoring replicas, M frames/s (e2e k frames/s) |")
print(hdr)
print("|" + "---|" * 10)
base = None
for n, x in sorted(d.items()):
    m, f = x["matching"], x["fragment"]
    base = base or m["ms_per_query"]
    w = m.get("weak_scaling")
    m5 = m.get("min_match_5", {})
    print(f"| {n} | {m['ms_per_query'] * 1e3:.1f} ({m['ms_per_query_back_to_back'] * 1e3:.1f}) | {m['value'] / 1e9:.1f} | "
          f"{base / m['ms_per_query']:.2f} | "
          + (f"{m5['ms_per_query'] * 1e3:.1f} ({m5['ms_per_query_back_to_back'] * 1e3:.1f})" if m5 else "—") + " | "
          f"{m['batched']['ms_per_pass'] * 1e3:.0f} → {m['batched']['value'] / 1e9:.0f} | "
          + (f"{w['ms_per_query'] * 1e3:.1f} → {w['value'] / 1e9:.0f}" if w else f"{m['ms_per_query'] * 1e3:.1f} → {m['value'] / 1e9:.1f}") + " | "
          f"{f['ms_per_query'] * 1e3:.1f} → {f['value'] / 1e9:.2f} | {m['e2e']['min_match_5']['ms_per_query'] * 1e3:.0f} | "
          f"{x['value'] / 1e6:.2f} ({x['e2e']['value'] / 1e3:.0f}) |")
print()
print("parity flags:", {n: x.get("parity") for n, x in sorted(d.items())})
