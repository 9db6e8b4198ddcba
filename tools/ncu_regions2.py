import csv, glob, os, re, subprocess, sys, tempfile
rep, lib, tag = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for cubin in glob.glob(os.path.join(tmp, '*sm_100a.cubin')):
    sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
    inside, cur = False, (None, 0)
    for l in sass:
        if l.startswith('.text.'):
            inside = tag in l
            continue
        if not inside: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', l): lines.append(cur)
    if lines: break
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [k for k, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = rows[starts[skip]:starts[skip + 1]]
hdr, data = rows[1], [r for r in rows[2:] if len(r) > 10]
iS, iE = hdr.index('# Samples'), hdr.index('Instructions Executed')
regions = [('count.cuh step4 (scatter)', 'icikt_count.cuh', 130, 240), ('count.cuh count_pass rest', 'icikt_count.cuh', 241, 430), ('count.cuh accessors', 'icikt_count.cuh', 1, 129),
           ('count_pass_inplace', 'icikt_pairs.cu', 160, 255),
           ('group_hist', 'icikt_pairs.cu', 256, 514), ('small_groups_direct', 'icikt_pairs.cu', 515, 575), ('small_groups_inplace', 'icikt_pairs.cu', 576, 621),
           ('large_groups_sorted', 'icikt_pairs.cu', 622, 782), ('large_groups_sorted2', 'icikt_pairs.cu', 783, 976), ('staged_gather', 'icikt_pairs.cu', 977, 1030),
           ('kernel body: unit/masks', 'icikt_pairs.cu', 1151, 1283), ('gather', 'icikt_pairs.cu', 1284, 1317), ('tail/reduce', 'icikt_pairs.cu', 1318, 1425)]
agg = {}
ts = te = 0
for k, r in enumerate(data):
    f, ln = lines[k] if k < len(lines) else ('?', 0)
    s, e = int(r[iS] or 0), int(r[iE] or 0)
    ts += s; te += e
    name = 'other: ' + str(f)
    for nm, ff, a, b in regions:
        if f == ff and a <= ln <= b: name = nm; break
    x = agg.setdefault(name, [0, 0]); x[0] += s; x[1] += e
for nm, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{nm:32s} instr {100*e/te:5.1f}%  samples {100*s/ts:5.1f}%')
print('total warp instr', te)
