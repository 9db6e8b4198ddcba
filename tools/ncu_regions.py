"""Executed-instruction share per barrier-delimited region of a kernel (ncu source page)."""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iE]) for r in data)
reg, cur = [], dict(start=0, n=0, ex=0, sm=0)
for k, r in enumerate(data):
    cur['n'] += 1; cur['ex'] += int(r[iE]); cur['sm'] += int(r[iSm])
    if 'BAR.SYNC' in r[iS]:
        cur['end'] = k; reg.append(cur); cur = dict(start=k + 1, n=0, ex=0, sm=0)
cur['end'] = len(data) - 1; reg.append(cur)
ts = sum(x['sm'] for x in reg)
print('total warp-instructions', tot)
for x in reg:
    if x['ex']:
        print(f"instr {x['start']:5d}-{x['end']:5d} n={x['n']:4d} exec={x['ex']:14d} ({100*x['ex']/tot:5.1f}%) samples={100*x['sm']/ts:5.1f}%")
