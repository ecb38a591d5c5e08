"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table of the LAST
training step: name, grid, duration (us), share.  usage: launch_table.py launches.csv [n_last]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if 'Kernel Name' in r)
i0 = rows.index(hdr)
ik, iv, ig = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
out = [(r[ik], r[ig], float(r[iv].replace(',', '')) / 1e3) for r in rows[i0 + 1:] if len(r) > iv]
# one step = from the last split / shadow kernel (first launch of a forward) to the end
starts = [i for i, o in enumerate(out) if 'split_params' in o[0] or 'shadow_params' in o[0]
          or 'group_indices' in o[0]]
step = out[starts[-1]:] if starts else out[-int(sys.argv[2]) if len(sys.argv) > 2 else -20:]
tot = sum(o[2] for o in step)
for name, grid, us in step:
  short = name.split('(')[0].replace('void ', '').replace('tfr::', '')
  print('%-58s %-14s %8.1f us %5.1f %%' % (short[:58], grid, us, 100 * us / tot))
print('%-58s %-14s %8.1f us' % ('TOTAL (kernels of one step)', '', tot))
