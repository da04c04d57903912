"""Compact view of a tools/sweep_r2.py table."""
import re, sys
rows = []
for line in open(sys.argv[1]):
    if line.startswith('| L'):
        f = [x.strip() for x in line.split('|')]
        d = dict(re.findall(r'([A-Za-z0-9 ]+?)=([0-9.]+|inf)', f[10].split(';')[0]))
        d = {k.strip(): float(v) for k, v in d.items()}
        rows.append((f[1], f[2], f[3], d, f[8], f[10]))
    elif line.startswith('##'):
        rows.append((line.strip(),))
    elif line.startswith('<!--'):
        print(line.strip())
for r in rows:
    if len(r) == 1:
        print(r[0]); continue
    blk, D, op, d, roof, allv = r
    if 'transpose-free' in op:
        b = min(d, key=d.get)
        print(f"   {blk[:34]:34s} D={D:>5} scatter  best {b}={d[b]} default={d.get('default')} | {allv.split(';')[-1].strip()}")
    else:
        rs = d.pop('rowsplit'); df = d.pop('default', None)
        srt = sorted(d.items(), key=lambda kv: kv[1])[:3]
        print(f"{blk[:37]:37s} D={D:>5} {op[:11]:11s} roof {roof:>5} rs {rs:6.1f} def {df} | " + ' '.join(f'{k}={v}' for k, v in srt))
