"""Executed-instruction / shared-memory-wavefront share per barrier-delimited region of a kernel
(ncu source page).  Usage: python tools/ncu_regions.py report.ncu-rep [--ops] [--skip K]
(--skip K: the report holds several launches, read the K-th)"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [k for k, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
which = int(sys.argv[sys.argv.index('--skip') + 1]) if '--skip' in sys.argv else 0
rows = rows[starts[which]:starts[which + 1]]
print(rows[0][1][:100])
hdr, data = rows[1], [r for r in rows[2:] if len(r) > 10]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
iW, iWi = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
def num(x):
    try: return int(x)
    except Exception: return 0
tot = sum(num(r[iE]) for r in data)
totw = sum(num(r[iW]) for r in data)
reg, cur = [], dict(start=0, n=0, ex=0, sm=0, wf=0, wfi=0)
for k, r in enumerate(data):
    cur['n'] += 1; cur['ex'] += num(r[iE]); cur['sm'] += num(r[iSm]); cur['wf'] += num(r[iW]); cur['wfi'] += num(r[iWi])
    if 'BAR.SYNC' in r[iS]:
        cur['end'] = k; reg.append(cur); cur = dict(start=k + 1, n=0, ex=0, sm=0, wf=0, wfi=0)
cur['end'] = len(data) - 1; reg.append(cur)
ts = sum(x['sm'] for x in reg)
print('total warp-instructions', tot, 'shared wavefronts', totw)
for x in reg:
    if x['ex']:
        print(f"instr {x['start']:5d}-{x['end']:5d} n={x['n']:4d} exec={x['ex']:14d} ({100*x['ex']/tot:5.1f}%) "
              f"samples={100*x['sm']/ts:5.1f}% smem_wavefronts={100*x['wf']/max(totw,1):5.1f}% (ideal {100*x['wfi']/max(totw,1):5.1f}%)")
if '--ops' in sys.argv:
    for k, r in enumerate(data):
        if num(r[iW]):
            print(k, r[iS][:60], 'exec', r[iE], 'wf', r[iW], 'ideal', r[iWi], 'per-instr %.2f' % (num(r[iW]) / max(num(r[iE]), 1)))
