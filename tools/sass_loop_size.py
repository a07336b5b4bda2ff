"""Code size of the loops of a kernel: largest backward branches in its SASS (cuobjdump -sass of an object file).

usage: python tools/sass_loop_size.py pasio_b200/csrc/exact_pruned.o exact_pruned_kernelILb1ELb1E [min_bytes]
Written to test whether the step loop of the exact DP's diagonal CTA (70 KB) suffered from the 32 KB instruction cache: shrinking it
to 48 KB changed nothing (profiles/r02_exact_dp_v9_notes.txt)."""
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    min_bytes = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    fn = None
    ins = {}
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
            ins[fn] = []
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
        if m and fn:
            ins[fn].append((int(m.group(1), 16), m.group(2).strip()))
    for fn, lst in ins.items():
        if pat not in fn:
            continue
        print(fn, 'total', len(lst) * 16, 'bytes')
        loops = []
        for a, t in lst:
            m = re.search(r'\bBRA(?:\.\S+)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)', t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and a - tgt >= min_bytes:
                    loops.append((a - tgt + 16, tgt, a))
        for size, tgt, a in sorted(loops, reverse=True)[:8]:
            print('  loop 0x%x .. 0x%x: %d bytes' % (tgt, a, size))


if __name__ == '__main__':
    main()
