#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/r01c_launches.csv profiles/r01_launches.md
  python tools/summarize_ncu.py kernel   gpurun_out/r01c_prof_mma.ncu-rep profiles/r01_accumulate_mma.md
"""
import collections
import csv
import io
import subprocess
import sys

KEY_METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__cycles_active.avg', 'gpc__cycles_elapsed.max', 'smsp__inst_executed.sum',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
    'lts__t_requests_srcunit_tex_op_red.sum', 'lts__t_sectors.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] in ('ns', 'nsecond') else v          # -> us
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(v for _, v in agg.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list summary\n\nsource: `{src}` (`ncu --metrics gpu__time_duration.sum --clock-control none`; '
                f'per-launch times are cold-cache and serialised: compare shares, not absolutes)\n\n'
                f'{sum(n for n, _ in agg.values())} launches, {total:.1f} us of kernel time in total\n\n'
                '| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n')
        for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| `{name[:110]}` | {n} | {v:.1f} | {v / n:.2f} | {100 * v / total:.1f} % |\n')
    print(open(dst).read()[:3000])


def kernel(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full summary\n\nsource: `{src}` (`ncu --set full --clock-control none --import-source on`, '
                'one GPU; cache control flushes L2 before every replay, so this is the cold-cache, isolated-launch view)\n\n')
        for r in rows[2:]:
            f.write(f'## `{r[hdr.index("Kernel Name")][:100]}` (launch id {r[hdr.index("ID")]})\n\n| metric | value | unit |\n|---|---:|---|\n')
            for m in KEY_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f'| {m} | {r[i]} | {units[i]} |\n')
            f.write('\n')
    print(open(dst).read()[:2500])


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
