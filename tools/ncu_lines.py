"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur_file = None; data = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] == '' : continue
    ii = hdr.index('Instructions Executed'); sa = hdr.index('# Samples')
    try: n = int(r[ii]); s = int(r[sa])
    except ValueError: continue
    data.append((n, s, cur_file, r[0], r[1][:100]))
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print('total warp-inst', tot, 'samples', tots)
for n, s, f, ln, src in sorted(data, reverse=True)[:top]:
    print('%5.2f%% inst %5.2f%% smp  %s:%s | %s' % (100.0*n/tot, 100.0*s/max(tots,1), f, ln, src.strip()))
