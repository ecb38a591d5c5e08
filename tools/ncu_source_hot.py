"""ncu-rep (captured with --import-source on, -lineinfo build) -> the source lines with the most
warp-stall samples of the first kernel in the report.
usage: ncu_source_hot.py report.ncu-rep out.txt [n_lines]"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr_i = next((i for i, r in enumerate(rows) if 'Source' in r and any('Sampl' in c for c in r)), None)
out = open(sys.argv[2], 'w')
if hdr_i is None:
  out.write(raw[:4000])
  sys.exit(0)
hdr = rows[hdr_i]
isrc = hdr.index('Source')
icol = [i for i, c in enumerate(hdr) if 'Sampl' in c][0]
inst = [i for i, c in enumerate(hdr) if 'Instructions Executed' in c]
data = []
for r in rows[hdr_i + 1:]:
  try:
    data.append((float(r[icol].replace(',', '') or 0), r))
  except (ValueError, IndexError):
    pass
tot = sum(d[0] for d in data) or 1.0
data.sort(key=lambda d: -d[0])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out.write('column: %s; total %d samples\n' % (hdr[icol], tot))
for v, r in data[:n]:
  out.write('%6.2f %%  %s\n' % (100 * v / tot, r[isrc].strip()[:150]))
out.close()
