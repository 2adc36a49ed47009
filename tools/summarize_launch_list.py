"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of the LAST step in order and
per kernel (python tools/summarize_launch_list.py launches.csv launches_per_step [first launch]); without a first launch
the last `launches_per_step` launches are taken, with `auto` the second run of launches that starts at a one-pass /
logit kernel of the forward (one full training step after warm-up)."""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
L = []
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    L.append((r[ki].split("(")[0][:64], us))
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(L)
if len(sys.argv) > 3 and sys.argv[3] == "auto":
    starts = [i for i, (k, _) in enumerate(L) if "fused_kernel<0" in k or "ks_kernel<2" in k or "ks_kernel<0" in k]
    first = starts[1] if len(starts) > 1 else starts[0]
    last = L[first:first + n]
elif len(sys.argv) > 3:
    last = L[int(sys.argv[3]):int(sys.argv[3]) + n]
else:
    last = L[-n:]
tot = sum(u for _, u in last)
print(f"{n} launches, {tot:.1f} us summed (cold-cache, serialised: shares, not absolutes)\n\nin launch order:")
for k, u in last:
    print(f"  {u:8.1f} us  {100 * u / tot:5.1f} %  {k}")
agg = {}
for k, u in last:
    a = agg.setdefault(k, [0.0, 0]); a[0] += u; a[1] += 1
print("\nby kernel:")
for k, (u, c) in sorted(agg.items(), key=lambda t: -t[1][0]):
    print(f"  {u:8.1f} us  {100 * u / tot:5.1f} %  x{c}  {k}")
st = sum(u for k, u in last if "fused_kernel" in k or "ks_kernel" in k or "kp_kernel" in k)
print(f"\ntoken-streaming kernels: {st:.1f} us = {100 * st / tot:.1f} % of the step; everything else {tot - st:.1f} us")
