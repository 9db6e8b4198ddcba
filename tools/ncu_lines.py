"""Stall samples and executed instructions of one kernel by CUDA source line: joins the SASS page of
an ncu report with the line table of the cubin (nvdisasm -g).
Usage: python tools/ncu_lines.py report.ncu-rep lib.so 'Li896ELi1ELb0ELb0ELi9E' [--skip K] [--top N]"""
import csv, glob, os, re, subprocess, sys, tempfile
rep, lib, tag = sys.argv[1:4]
skip = int(sys.argv[sys.argv.index('--skip') + 1]) if '--skip' in sys.argv else 0
top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []  # source line of every instruction of the function, in address order
for cubin in glob.glob(os.path.join(tmp, '*sm_100a.cubin')):
    sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
    inside, cur = False, (None, 0)
    for l in sass:
        if l.startswith('.text.'):
            inside = ('pairs_tiled_kernel' in l or tag in l) and tag in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', l):
            lines.append(cur)
    if lines:
        break
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [k for k, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
rows = rows[starts[skip]:starts[skip + 1]]
hdr, data = rows[1], [r for r in rows[2:] if len(r) > 10]
iS, iE = hdr.index('# Samples'), hdr.index('Instructions Executed')
print(rows[0][1][:110], '| sass rows', len(data), '| line table', len(lines))
agg = {}
for k, r in enumerate(data):
    key = lines[k] if k < len(lines) else ('?', 0)
    a = agg.setdefault(key, [0, 0])
    a[0] += int(r[iS] or 0); a[1] += int(r[iE] or 0)
ts = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
src = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), 'csrc', f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ''
    print(f'{f}:{ln:5d} samples {100*a[0]/ts:5.1f}% instr {100*a[1]/te:5.1f}%  {text}')
