"""Print the hottest SASS instructions (by warp-stall samples) of an ncu report's source page."""
import csv, subprocess, sys, io
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(out))]
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
body = [r for r in rows[hdr_i + 1:] if len(r) > 5 and r[2].isdigit()]
# only the first kernel instance
seen = set(); first = []
for r in body:
    if r[0] in seen: break
    seen.add(r[0]); first.append(r)
tot = sum(int(r[2]) for r in first)
print('total samples', tot, 'instructions', len(first))
top = sorted(enumerate(first), key=lambda x: -int(x[1][2]))[:n]
for i, r in sorted(top):
    print('%5d %6s %5.1f%% exec=%8s  %s' % (i, r[2], 100 * int(r[2]) / tot, r[5], r[1].strip()[:100]))
