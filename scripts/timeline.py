"""Per-stream timeline of the convolution launches of one traced step (bench.py --trace-layers CSV): where the streams idle, which
launches overlap, what sits on the critical path.   python scripts/timeline.py layers.csv [min_gap_us]"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.DictReader(open(sys.argv[1])) if float(r["ms"]) >= 0 and float(r["t0_ms"]) >= 0]
gap_min = float(sys.argv[2]) / 1e3 if len(sys.argv) > 2 else 0.05
streams = defaultdict(list)
for r in rows:
    streams[r["stream"]].append(r)
t_end = max(float(r["t0_ms"]) + float(r["ms"]) for r in rows)
print(f"{len(rows)} launches on {len(streams)} streams, span {t_end:.3f} ms")
order = sorted(streams, key=lambda s: -sum(float(r['ms']) for r in streams[s]))
for si, s in enumerate(order):
    rs = sorted(streams[s], key=lambda r: float(r["t0_ms"]))
    busy = sum(float(r["ms"]) for r in rs)
    print(f"\n== stream {si} ({s}): {len(rs)} launches, busy {busy:.3f} ms, first {float(rs[0]['t0_ms']):.3f} last end {float(rs[-1]['t0_ms']) + float(rs[-1]['ms']):.3f}")
    prev_end = None
    for r in rs:
        t0, ms = float(r["t0_ms"]), float(r["ms"])
        gap = t0 - prev_end if prev_end is not None else 0.0
        flag = f"   <-- {gap * 1e3:7.0f} us since the previous conv launch of this stream" if gap >= gap_min else ""
        if si < 2 or flag or ms >= 0.1:
            print(f"  {t0:8.3f} +{ms:6.3f}  {r['kind']:11s} {r['kernel']:13s} {r['Cin']:>3s}->{r['Cout']:>3s} {r['Ld']}x{r['Lh']}x{r['Lw']} s{r['sd']}{r['sh']}{r['sw']} T{r['T']}{flag}")
        prev_end = t0 + ms
