"""Aggregate warp-stall samples of an ncu report per source file / source line (needs -lineinfo and
--import-source on).  usage: ncu_source.py file.ncu-rep [topN]"""
import csv, io, subprocess, sys, collections
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = cur = None
agg, inst, lines = collections.Counter(), collections.Counter(), []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1]; continue
    if r[0] == 'Line No':
        hdr = r; si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed'); continue
    if hdr and cur and r[0] not in ('', 'Function Name'):
        try:
            s, ie = int(r[si]), int(r[ii])
        except ValueError:
            continue
        agg[cur] += s; inst[cur] += ie
        lines.append((s, ie, cur.split('/')[-1], r[0], r[1][:100]))
tot = sum(agg.values()); ti = sum(inst.values())
print('total samples', tot, 'warp instructions', ti)
for f, s in agg.most_common():
    print('%6.2f%% samples  %6.2f%% inst  %s' % (100.0 * s / tot, 100.0 * inst[f] / ti, f))
print('--- top lines')
for s, ie, f, ln, src in sorted(lines, reverse=True)[:top]:
    print('%5.2f%% inst=%-9d %s:%s  %s' % (100.0 * s / tot, ie, f, ln, src))
