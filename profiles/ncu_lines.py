"""Per-source-line roll-up of an ncu --set full --import-source on report: samples and executed warp instructions per
CUDA source line of one kernel launch.   python profiles/ncu_lines.py report.ncu-rep <launch index> [min share]"""
import csv, subprocess, sys, collections
rep, idx = sys.argv[1], int(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = -1
i = 0
per = collections.OrderedDict()
name = ''
while i < len(rows):
    r = rows[i]
    if r and r[0] == 'Function Name':
        name = r[1]
    if r and r[0] == 'Line No' and '# Samples' in r:
        k += 1
        hdr = r
        ln, src, sa, ex = hdr.index('Line No'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
        i += 1
        while i < len(rows) and not (rows[i] and rows[i][0] in ('Line No', 'File Path', 'Function Name')):
            rr = rows[i]
            if k == idx and len(rr) > ex:
                try:
                    key = (int(rr[ln]), rr[src][:110])
                    a = per.setdefault(key, [0, 0])
                    a[0] += int(rr[sa]); a[1] += int(rr[ex])
                except ValueError:
                    pass
            i += 1
        if k == idx:
            print('kernel', name[:120])
        continue
    i += 1
ts = sum(v[0] for v in per.values()); te = sum(v[1] for v in per.values())
print('samples', ts, 'warp instructions', te)
for (ln, src), (s, e) in per.items():
    if s > ts * thr or e > te * thr:
        print(f"{ln:5d} {100*s/max(ts,1):5.1f}% smp {100*e/max(te,1):5.1f}% inst  {src}")
