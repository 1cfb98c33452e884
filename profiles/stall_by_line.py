"""Stall samples and instruction counts per CUDA source line (see sass_by_line.py).
usage: stall_by_line.py <nvdisasm -g -c out> <kernel substring> <ncu source csv> <units (e.g. particles*steps)>"""
import csv, re, sys, collections
sass, kern, ncu = sys.argv[1:4]; units = float(sys.argv[4])
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l)
seq = []; cur = ('?', 0)
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        if seq: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): seq.append((cur, l.strip()))
rows = list(csv.reader(open(ncu)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
st = collections.Counter(); ins = collections.Counter(); tot = 0; itot = 0; perinst = []
for (k, txt), r in zip(seq, body):
    v = int(r[idx['Warp Stall Sampling (All Samples)']]); st[k] += v; tot += v
    n = int(r[idx['Instructions Executed']]); ins[k] += n; itot += n
    perinst.append((v, k, txt, n))
print(f"warp instructions {itot}; thread instructions per unit {itot*32/units:.1f}")
print("lines by stall samples:")
for k, v in st.most_common(14): print(f"  {k[0]}:{k[1]:4d}  {v/tot*100:5.1f}% stall   {ins[k]*32/units:6.1f} instr/unit")
print("lines by instructions:")
for k, v in ins.most_common(16): print(f"  {k[0]}:{k[1]:4d}  {v*32/units:6.1f} instr/unit   {st[k]/tot*100:5.1f}% stall")
print("top stalled instructions:")
for v, k, txt, n in sorted(perinst, reverse=True)[:8]: print(f"  {v/tot*100:5.1f}%  {k[0]}:{k[1]}  n={n}  {txt[:80]}")
reasons = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
agg = collections.Counter()
for r in body:
    for h in reasons:
        try: agg[h] += int(r[idx[h]])
        except ValueError: pass
s = sum(agg.values())
print("stall reasons:", [(k, round(v / s * 100, 1)) for k, v in agg.most_common(9)])
