"""Aggregates an ncu report's instruction / stall-sample counts by CUDA source line:
    python tools/ncu_lines.py rep.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name',
                      f'regex:{kre}', '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, agg, hdr = None, {}, None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Name':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 4 and r[0] == 'Line No':
        hdr = r
        iI, iS = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None or len(r) <= iI:
        continue
    if r[0].isdigit() and r[iI].isdigit() and r[iS].isdigit():
        a = agg.setdefault((cur, int(r[0]), r[1].strip()[:90]), [0, 0])
        a[0] += int(r[iI])
        a[1] += int(r[iS])
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print(f'total warp instructions {tot}, stall samples {tots}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{str(k[0]):18s}:{k[1]:4d} {k[2]:90s} {100 * a[0] / tot:5.1f}% instr {100 * a[1] / tots:5.1f}% samples')
