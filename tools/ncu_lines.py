"""Per-source-line totals of one metric column of the ncu source page (first kernel): 
python tools/ncu_lines.py rep.ncu-rep "L1 Wavefronts Shared" [top]"""
import csv, io, subprocess, sys, collections
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
col = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
# the cuda,sass view lists: source line rows (Line No ...) followed by their SASS rows? fall back to SASS-only aggregation by opcode
hdr = None; agg = collections.Counter(); aggop = collections.Counter(); instr = collections.Counter(); nk = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        nk += 1
        if nk > 1: break
    if r and r[0] in ('Address', 'Line No'):
        hdr = r; continue
    if hdr and len(r) == len(hdr) and col in hdr:
        v = r[hdr.index(col)]
        try: v = float(v)
        except ValueError: continue
        key = r[1].strip()
        if hdr[0] == 'Address':
            op = key.split()[0] if not key.startswith('@') else key.split()[1]
            aggop[op.split('.')[0] + ('.' + op.split('.')[1] if '.' in op else '')] += v
            instr[op.split('.')[0]] += float(r[hdr.index('Instructions Executed')] or 0)
        else:
            agg[(r[0], key[:90])] += v
print('by opcode:')
for k, v in aggop.most_common(12): print('  %-16s %14.0f' % (k, v))
print('instructions executed by opcode:')
for k, v in instr.most_common(25): print('  %-16s %14.0f' % (k, v))
if agg:
    print('by source line:')
    for (ln, src), v in agg.most_common(top): print('  %12.0f  %s: %s' % (v, ln, src))
