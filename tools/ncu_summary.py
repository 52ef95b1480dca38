'''
Turns ncu output brought back in gpurun_out/ into the small text / JSON
summaries committed under profiles/:

    python tools/ncu_summary.py launches <launches.csv> <out.txt>
    python tools/ncu_summary.py kernels <report.ncu-rep> <out.json> [capture label]

`kernels` merges into an existing <out.json>: one entry per kernel name (template arguments included), the first
launch of each, tagged with the capture label and the hash of the sources the library was built from
(composer_b200/build.py source_hash), which is what bench.py checks before it quotes `traffic` from an entry.
'''
import collections, csv, json, os, re, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('cb200::', '')
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1000 if row['Metric Unit'] == 'ns' else v * 1000 if row['Metric Unit'] == 'ms' else v
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    with open(out, 'w') as handle:
        handle.write('# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and '
                     'serialised: compare shares)\n')
        handle.write('# %d launches, %.1f us in total\n' % (sum(cnt.values()), total))
        handle.write('%-64s %7s %12s %10s %7s\n' % ('kernel', 'count', 'total us', 'avg us', 'share'))
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            handle.write('%-64s %7d %12.1f %10.1f %6.1f%%\n' % (k[:64], cnt[k], v, v / cnt[k], 100 * v / total))
    print(open(out).read())


def kernels(path, out, capture=None):
    from composer_b200 import build as native_build
    source_hash = native_build.source_hash()
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = {'gpu__time_duration.sum': 'duration', 'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
            'launch__registers_per_thread': 'registers', 'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
            'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct',
            'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'xu_pipe_pct',
            'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active': 'alu_pipe_pct',
            'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active': 'fma_pipe_pct',
            'smsp__inst_executed.sum': 'warp_instructions',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct'}
    result = {}
    units = rows[1]
    for r in rows[2:]:
        name = re.sub(r'\((int|bool|unsigned int)\)', '', r[hdr.index('Kernel Name')])
        name = re.sub(r'\(.*', '', name).replace('void ', '').replace('cb200::', '')
        entry = {}
        for metric, key in want.items():
            if metric in hdr:
                i = hdr.index(metric)
                entry[key] = '%s %s' % (r[i], units[i])
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
        try:
            rd, ru = entry['dram_read'].split()
            wr, wu = entry['dram_write'].split()
            entry['dram_bytes_per_launch'] = float(rd) * scale[ru] + float(wr) * scale[wu]
        except Exception:
            pass
        name = re.sub(r'\((int|bool|unsigned int)\)', '', name)
        entry['capture'] = capture or os.path.basename(path)
        entry['source_hash'] = source_hash
        result.setdefault(name, entry)
    if os.path.exists(out):
        merged = json.load(open(out))
        merged.update(result)
        result = merged
    json.dump(result, open(out, 'w'), indent=1)
    print(json.dumps(result, indent=1))


if __name__ == '__main__':
    {'launches': launches, 'kernels': kernels}[sys.argv[1]](*sys.argv[2:])
