"""Dynamic instruction breakdown of one captured launch: opcode mix and straight-line blocks (same execution count) of an .ncu-rep source page."""
import csv, collections, subprocess, sys, io
rep, kern, skip = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '0'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
ends = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
k = int(skip) if len(ends) > 1 and len(ends) > int(skip) else 0
rows = rows[ends[k]: ends[k + 1] if k + 1 < len(ends) else len(rows)]
print(rows[0][1][:100])
H = rows[1]; ie = H.index('Instructions Executed'); isrc = H.index('Source'); isamp = H.index('# Samples')
def opc(src):
    t = src.split()
    return (t[1] if t[0].startswith('@') else t[0])
data = [(int(r[ie]), int(r[isamp]), r[isrc].strip()) for r in rows[2:] if len(r) > ie]
tot = sum(d[0] for d in data); ts = max(1, sum(d[1] for d in data))
print('total warp instructions', tot, 'stall samples', ts)
c = collections.Counter(); s = collections.Counter()
for e, sm, src in data:
    op = opc(src).split('.')[0]; c[op] += e; s[op] += sm
for op, v in c.most_common(14):
    print(f'  {op:10s} {v:12d} {v / tot:.3f}  samples {s[op] / ts:.3f}')
blocks = []
for e, sm, src in data:
    if blocks and blocks[-1][0] == e: blocks[-1][1] += 1; blocks[-1][2] += sm; blocks[-1][3].append(opc(src))
    else: blocks.append([e, 1, sm, [opc(src)]])
for b in blocks:
    if b[0] * b[1] > 0.02 * tot or b[2] > 0.03 * ts:
        cc = collections.Counter(x.split('.')[0] for x in b[3])
        print(f'  exec {b[0]:9d} x {b[1]:4d} = {b[0] * b[1] / tot:.3f} of instr, {b[2] / ts:.3f} of samples', dict(cc.most_common(7)))
