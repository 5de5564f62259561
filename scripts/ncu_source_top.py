"""Top instructions by stall samples / excessive shared wavefronts from `ncu --page source --csv` output.
usage: ncu -i rep --page source --csv --kernel-name regex:NAME --launch-count 1 | python scripts/ncu_source_top.py"""
import csv, sys
rows = [r for r in csv.reader(sys.stdin)]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break                                   # first kernel of the report only
    if len(r) == len(hdr):
        data.append(r)
def num(r, k):
    try:
        return int(float(r[ix[k]] or 0))
    except (KeyError, ValueError):
        return 0
tot = sum(num(r, "# Samples") for r in data)
print("kernel", rows[0][1][:100] if rows[0] else "", "instructions", len(data), "samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(num(r, h) for r in data) for h in stall_cols}
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:n]:
    st = sorted(((num(r, h), h[6:]) for h in stall_cols if num(r, h)), reverse=True)[:3]
    print(f"{num(r, '# Samples'):6d} {num(r, 'Instructions Executed'):>10d} {r[ix['Source']][:64]:64s} {st}")
print("-- excessive shared wavefronts")
for r in sorted(data, key=lambda r: -num(r, "L1 Wavefronts Shared Excessive"))[:6]:
    print(f"{r[ix['Source']][:64]:64s} wf={num(r, 'L1 Wavefronts Shared')} ideal={num(r, 'L1 Wavefronts Shared Ideal')} "
          f"excess={num(r, 'L1 Wavefronts Shared Excessive')}")
