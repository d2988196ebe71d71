"""Summarise ncu outputs into the small CSV/JSON files kept under profiles/.

  python scripts/ncu_summary.py launches <launches.csv> <out.csv>
      per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list
  python scripts/ncu_summary.py full <raw.csv> <out.csv> <out.json>
      one row per captured launch of `ncu --set full` (export with --page raw --csv): time, DRAM bytes,
      tensor / TC pipe %, issue %; the JSON holds the per-launch average DRAM traffic of the conv family
      (bench.py reads it for roofline.traffic)
"""
import csv
import json
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::|void ', '', name)
    name = re.sub(r'\(int\)|\(bool\)', '', name)
    return name.split('(')[0].strip()


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r]
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hdr]
    ik, iv, iu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    tot = OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'ns ': 1e-3}.get(r[iu], 1.0)
        k = short(r[ik])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += v
    total = sum(t[1] for t in tot.values())
    with open(dst, 'w') as f:
        f.write('kernel,launches,total_us,share_pct\n')
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.1f,%.2f\n' % (k, n, us, 100 * us / total))
    print('launch list: %d kernels, %.1f us total -> %s' % (len(tot), total, dst))


def full(src, dst, dst_json):
    rows = list(csv.reader(open(src)))
    h, units, data = rows[0], rows[1], rows[2:]

    def col(name):
        return h.index(name) if name in h else None

    def val(r, name, scale=None):
        i = col(name)
        if i is None or r[i] == '':
            return float('nan')
        v = float(r[i].replace(',', ''))
        u = units[i]
        if scale == 'us':
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'msecond': 1e3, 'usecond': 1.0, 'nsecond': 1e-3}.get(u, 1.0)
        if scale == 'MB':
            v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
        return v

    out = []
    for r in data:
        out.append((short(r[col('Kernel Name')]), val(r, 'gpu__time_duration.sum', 'us'),
                    val(r, 'dram__bytes_read.sum', 'MB'), val(r, 'dram__bytes_write.sum', 'MB'),
                    val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                    val(r, 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active'),
                    val(r, 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active'),
                    val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                    val(r, 'launch__registers_per_thread'), val(r, 'launch__grid_size')))
    with open(dst, 'w') as f:
        f.write('kernel,us,dram_read_MB,dram_write_MB,dram_pct,hmma_issue_pct,tc_pipe_pct,issue_pct,regs,grid\n')
        for o in out:
            f.write('"%s",%.1f,%.1f,%.1f,%.1f,%.2f,%.1f,%.1f,%.0f,%.0f\n' % o)
    conv = [o for o in out if o[0].startswith(('conv_tc_kernel', 'conv_xc_kernel', 'conv_qd_kernel', 'conv_qf_kernel',
                                               'conv_qu_kernel', 'first_conv_kernel'))]
    if conv:
        js = {'launches': len(conv),
              'avg_dram_bytes_per_launch': sum((o[2] + o[3]) for o in conv) / len(conv) * 1e6,
              'total_us': sum(o[1] for o in conv),
              'total_dram_bytes': sum((o[2] + o[3]) for o in conv) * 1e6,
              'time_weighted_tc_pipe_pct': sum(o[1] * o[6] for o in conv) / sum(o[1] for o in conv),
              'source': 'ncu --set full --clock-control none, one forward pass'}
        json.dump(js, open(dst_json, 'w'), indent=1)
        print(js)
    print('full capture: %d launches -> %s' % (len(out), dst))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
