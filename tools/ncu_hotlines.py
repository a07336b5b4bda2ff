"""Hottest source lines (warp-stall samples) of the first kernel in an .ncu-rep; needs -lineinfo."""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, data = '', []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        i = [k for k, h in enumerate(hdr) if h == '# Samples'][0]
        data.append((int(r[i]), cur_file, int(r[0]), r[1].strip()[:100]))
tot = sum(d[0] for d in data) or 1
print('total samples', tot)
for d in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print('%6.2f%%  %s:%d  %s' % (100.0 * d[0] / tot, d[1], d[2], d[3]))
