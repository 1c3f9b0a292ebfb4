"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
  python tools/ncu_summary.py launches <csv> <out.txt>
  python tools/ncu_summary.py full <rep.ncu-rep> <out.txt>
"""
import collections, csv, re, subprocess, sys

def kname(s):
    m = re.search(r'(hb_[a-z0-9_]+(?:<[^>]*>)?)', s)
    return m.group(1) if m else s[:60]

def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); iu = hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(',', ''))
        u = r[iu]
        v_us = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
        a = agg.setdefault(kname(r[ik]), [0, 0.0]); a[0] += 1; a[1] += v_us
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({path})\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n:50s} launches {c:4d}  total {t/1e3:9.3f} ms  avg {t/c:9.1f} us  share {t/tot*100:5.1f}%\n")
    print(open(out).read())

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sector_hit_rate.pct']

def full(path, out):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write(f"# ncu --set full --clock-control none ({path}); selected raw metrics per captured launch\n")
        for r in rows[2:]:
            f.write(f"\n## {kname(r[hdr.index('Kernel Name')])}\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w); f.write(f"{w:70s} {r[i]:>16s} {units[i]}\n")
    print(open(out).read())

def traffic(path, out):
    """dram__bytes_read.sum + dram__bytes_write.sum per captured launch -> profiles/ncu_traffic.json (bench.py reports it as
    roofline.traffic next to the algorithmic bytes)."""
    import json
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    res = {}
    for r in rows[2:]:
        k = kname(r[hdr.index('Kernel Name')])
        key = 'K1' if 'k1_' in k else 'K3' if 'k3_' in k else 'K2'
        b = float(r[ir].replace(',', '')) * scale[units[ir]] + float(r[iw].replace(',', '')) * scale[units[iw]]
        res[key] = b
        res[key + '_kernel'] = k
    res['_source'] = f'ncu --set full --clock-control none, {path}: dram__bytes_read.sum + dram__bytes_write.sum per launch'
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'traffic': traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
