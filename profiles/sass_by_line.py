#!/usr/bin/env python
"""Join an ncu source-page CSV (SASS view) with nvdisasm line info and aggregate per CUDA source line.

    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all sqaod_b200/lib/libsqaod_b200.so   (in a scratch directory)
    python profiles/sass_by_line.py src.csv dense_annealer.sm_100a.cubin <mangled kernel name> [top]
Prints the source lines with the most executed warp instructions and the most stall samples."""
import csv, re, subprocess, sys, collections

src_csv, cubin, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur, inside = None, False
for ln in dis:
    if ln.startswith('.section') or ln.startswith('//--------------------- .text.'):
        inside = ('.text.' + kernel) in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(rows[2][ia], 16)
inst, samp = collections.Counter(), collections.Counter()
stall = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    off = int(r[ia], 16) - base
    key = line_of.get(off, ('?', 0))
    inst[key] += int(r[ii]); samp[key] += int(r[isamp])
    for i, h in stall_cols:
        if r[i] not in ('', '0'):
            stall[key][h] += int(r[i])
ti, ts = sum(inst.values()), sum(samp.values())
print('total warp instructions %d, samples %d' % (ti, ts))
print('--- by executed instructions')
for k, v in inst.most_common(top):
    print('%s:%d  inst %.1f%%  samples %.1f%%' % (k[0], k[1], 100. * v / ti, 100. * samp[k] / ts))
print('--- by stall samples')
for k, v in samp.most_common(top):
    print('%s:%d  samples %.1f%%  inst %.1f%%  %s' % (k[0], k[1], 100. * v / ts, 100. * inst[k] / ti,
                                                 ' '.join('%s=%d' % (h[6:], c) for h, c in stall[k].most_common(3))))
if len(sys.argv) > 6:
    lo, hi = int(sys.argv[5]), int(sys.argv[6])
    print('--- lines %d..%d' % (lo, hi))
    tot_i = tot_s = 0
    for k in sorted(inst):
        if k[0].endswith('.cu') and lo <= k[1] <= hi and (inst[k] > ti * 0.0005 or samp[k] > ts * 0.0005):
            print('%s:%d  inst %.2f%%  samples %.2f%%  %s' % (k[0], k[1], 100. * inst[k] / ti, 100. * samp[k] / ts,
                                                         ' '.join('%s=%d' % (h[6:], c) for h, c in stall[k].most_common(3))))
    for k in inst:
        if k[0].endswith('.cu') and lo <= k[1] <= hi:
            tot_i += inst[k]; tot_s += samp[k]
    print('range total: inst %.1f%%  samples %.1f%%' % (100. * tot_i / ti, 100. * tot_s / ts))
