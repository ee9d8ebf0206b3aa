"""Turn gpurun_out ncu artefacts into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/summarize_ncu.py report   gpurun_out/prof_gather.ncu-rep profiles/r01_gather_ncu.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(src, dst):
    rows = []
    with open(src) as f:
        lines = [l for l in f if not l.startswith('==')]
    for r in csv.DictReader(io.StringIO(''.join(lines))):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(r['Metric Value'].replace(',', ''))
            unit = r.get('Metric Unit', 'ns')
            if unit in ('us', 'usecond'):
                v *= 1e3
            elif unit in ('ms', 'msecond'):
                v *= 1e6
            rows.append((r['Kernel Name'], v))
    agg = OrderedDict()
    for k, v in rows:
        name = k.split('(')[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(v for _, v in rows)
    with open(dst, 'w') as f:
        f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n')
        f.write('source: %s, %d launches, %.1f us total\n\n| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n' % (src, len(rows), total / 1e3))
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.1f | %.1f%% | %.2f |\n' % (name[:90], n, t / 1e3, 100 * t / total, t / n / 1e3))
    print(open(dst).read())


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def report(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    with open(dst, 'w') as f:
        f.write('# ncu --set full summary of %s\n\n' % src)
        for row in data:
            d = dict(zip(hdr, row))
            f.write('## %s  (grid %s, block %s)\n\n| metric | value | unit |\n|---|---|---|\n' % (
                d.get('Kernel Name', '?')[:100], d.get('Grid Size', '?'), d.get('Block Size', '?')))
            for k in hdr:
                if any(k == key or k.startswith(key) for key in KEYS):
                    f.write('| %s | %s | %s |\n' % (k, d[k], units[hdr.index(k)]))
            f.write('\n')
    print(open(dst).read()[:6000])


if __name__ == '__main__':
    {'launches': launches, 'report': report}[sys.argv[1]](sys.argv[2], sys.argv[3])
