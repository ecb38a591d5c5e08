"""ncu-rep (captured with --set full --import-source on) -> the SASS instructions with the most
warp-stall samples of the first kernel in the report, with their dominant stall reasons.
usage: ncu_source_hot.py report.ncu-rep out.txt [n_lines]"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr_i = next(i for i, r in enumerate(rows) if 'Address' in r and 'Source' in r)
hdr = rows[hdr_i]
isrc, ismp = hdr.index('Source'), hdr.index('# Samples')
stall = [(i, c) for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
data = []
for r in rows[hdr_i + 1:]:
  try:
    data.append((float(r[ismp] or 0), r))
  except (ValueError, IndexError):
    pass
tot = sum(d[0] for d in data) or 1.0
out = open(sys.argv[2], 'w')
out.write('%s\n%d stall samples over %d SASS instructions\n' % (rows[0][1][:120], tot, len(data)))
tots = {}
for _, r in data:
  for i, c in stall:
    try:
      tots[c] = tots.get(c, 0.0) + float(r[i] or 0)
    except ValueError:
      pass
out.write('by reason: ' + ', '.join('%s %.1f%%' % (k[6:], 100 * v / tot) for k, v in
                                  sorted(tots.items(), key=lambda kv: -kv[1])[:8]) + '\n')
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
order = sorted(range(len(data)), key=lambda j: -data[j][0])[:n]
for j in sorted(order):
  v, r = data[j]
  top = sorted(((float(r[i] or 0), c[6:]) for i, c in stall), reverse=True)[:2]
  out.write('%5.2f%%  #%-5d %-70s %s\n' % (100 * v / tot, j, r[isrc].strip()[:70],
                                          ' '.join('%s=%d' % (c, x) for x, c in top if x > 0)))
out.close()
