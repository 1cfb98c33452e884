"""Attribute ncu per-SASS-instruction counts to CUDA source lines.
usage: sass_by_line.py <nvdisasm -g -c output> <mangled kernel substring> <ncu --page source --csv output>
nvdisasm prints `//## File "...", line N` markers; ncu's SASS page lists the same instructions in the same order."""
import csv, re, sys, collections
sass, kern, ncu = sys.argv[1:4]
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l)
seq = []          # (file, line, inline-chain) per instruction
cur = ('?', 0)
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        if seq: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        seq.append(cur)
rows = list(csv.reader(open(ncu)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
print(len(seq), "sass instrs in cubin;", len(body), "in ncu", file=sys.stderr)
agg = collections.Counter(); st = collections.Counter()
for (f, ln), r in zip(seq, body):
    agg[(f, ln)] += int(r[idx['Instructions Executed']]); st[(f, ln)] += int(r[idx['Warp Stall Sampling (All Samples)']])
tot = sum(agg.values()); stt = sum(st.values())
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:60]:
    print(f"{k[0]:18s}:{k[1]:4d}  {v/tot*100:5.1f}% instr  {v*scale:8.1f}/unit  {st[k]/stt*100:5.1f}% stall")
