"""Bytes of SASS per source line of one function of an object file, inside an address range (nvdisasm -gi line info).

usage: python tools/sass_code_map.py OBJ FUNCTION_SUBSTRING FILE_SUFFIX LINE_LO LINE_HI [ADDR_LO ADDR_HI]
Every instruction is charged to the OUTERMOST frame of its inline chain that lies in FILE_SUFFIX between LINE_LO and
LINE_HI (the body of the function of interest), e.g. the step loop of xp_diagonal in exact_pruned.cu."""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, pat, suffix, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    alo = int(sys.argv[6], 16) if len(sys.argv) > 6 else 0
    ahi = int(sys.argv[7], 16) if len(sys.argv) > 7 else 1 << 60
    obj = os.path.abspath(obj)
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(['cuobjdump', '-xelf', 'all', obj], cwd=d, capture_output=True)
        dis = subprocess.run(['nvdisasm', '-gi', '-c'] + glob.glob(d + '/*.cubin'), capture_output=True, text=True).stdout
    line_re = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    addr_re = re.compile(r'^\s*/\*([0-9a-f]{4,6})\*/\s+(\S.*?);')
    inside = False
    frames = []           # frames of the current location, innermost first
    fresh = True
    per_line = collections.Counter()
    total = 0
    for l in dis.split('\n'):
        if l.startswith('.text.'):
            inside = pat in l
            continue
        if not inside:
            continue
        m = line_re.search(l)
        if m:
            if fresh:
                frames = []
                fresh = False
            frames.append((m.group(1), int(m.group(2))))
            if m.group(3):
                frames.append((m.group(3), int(m.group(4))))
            continue
        m = addr_re.match(l)
        if m:
            fresh = True
            a = int(m.group(1), 16)
            if not (alo <= a < ahi):
                continue
            total += 16
            key = None
            for f, n in frames:                 # outermost matching frame wins (last in the list)
                if f.endswith(suffix) and lo <= n <= hi:
                    key = n
            per_line[key] += 16
    print('bytes in range:', total)
    for k in sorted(per_line, key=lambda x: (x is None, x)):
        if per_line[k] >= 128:
            print(k, per_line[k])


if __name__ == '__main__':
    main()
