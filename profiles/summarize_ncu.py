"""Turn an .ncu-rep (one kernel launch, `ncu --set full`) into the small JSON / text summaries kept in profiles/.

  python profiles/summarize_ncu.py REPORT.ncu-rep OUT.json [--stalls OUT.txt] [--kernel NAME --workload TEXT --command TEXT
                                                           --algorithmic-bytes N]
Reads the report with `ncu -i ... --page raw --csv` (metrics) and `--page source --csv` (stall samples per SASS line).
"""
import argparse, collections, csv, io, json, subprocess, sys

KEEP = ['dram__bytes_read.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum', 'launch__block_size',
        'launch__grid_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg.per_second', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']
SCALE = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-6, 'us': 1e-3, 'ms': 1., 's': 1e3}


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'] + list(extra), capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report'); ap.add_argument('out')
    ap.add_argument('--stalls'); ap.add_argument('--kernel', default=''); ap.add_argument('--workload', default='')
    ap.add_argument('--command', default=''); ap.add_argument('--algorithmic-bytes', type=float, default=0)
    ap.add_argument('--note', default='')
    a = ap.parse_args()
    rows = ncu_csv(a.report, 'raw')
    hdr = next(r for r in rows if 'Kernel Name' in r)
    i0 = rows.index(hdr)
    units, vals = rows[i0 + 1], rows[i0 + 2]
    metrics = {}
    for k in KEEP:
        if k in hdr:
            j = hdr.index(k)
            metrics[k] = {'unit': units[j], 'value': vals[j]}

    def scaled(name):
        m = metrics.get(name)
        if not m:
            return None
        return float(m['value'].replace(',', '')) * SCALE.get(m['unit'], 1.)
    dram = (scaled('dram__bytes_read.sum') or 0) + (scaled('dram__bytes_write.sum') or 0)
    ms = scaled('gpu__time_duration.sum')
    d = {'kernel': a.kernel or vals[hdr.index('Kernel Name')], 'workload': a.workload, 'command': a.command,
         'dram_bytes_per_launch': dram, 'duration_ms_under_ncu': ms,
         'dram_GBps_under_ncu': dram / (ms * 1e-3) / 1e9 if ms else None, 'note': a.note, 'metrics': metrics}
    if a.algorithmic_bytes:
        d['algorithmic_bytes_per_launch'] = int(a.algorithmic_bytes)
    json.dump(d, open(a.out, 'w'), indent=1)
    if a.stalls:
        rows = ncu_csv(a.report, 'source', ['--print-source', 'sass'])
        hdr = next(r for r in rows if 'Source' in r and 'Address' in r)
        data = rows[rows.index(hdr) + 1:]
        iS, iN = hdr.index('Source'), hdr.index('# Samples')
        cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        tot = collections.Counter(); total = 0; top = []
        for r in data:
            if len(r) <= iN or not r[iN].isdigit():
                continue
            n = int(r[iN]); total += n; top.append((n, r[iS].strip()))
            for i, h in cols:
                tot[h] += int(r[i] or 0)
        with open(a.stalls, 'w') as f:
            f.write('# ncu source-page stall summary, %s\n' % d['kernel'])
            f.write('total samples %d\n' % total)
            for h, v in tot.most_common():
                if v:
                    f.write('%-26s %8d %5.1f%%\n' % (h, v, 100. * v / max(1, sum(tot.values()))))
            f.write('\ntop SASS instructions by samples:\n')
            for n, src in sorted(top, reverse=True)[:16]:
                f.write('%7d  %s\n' % (n, src))


if __name__ == '__main__':
    main()
