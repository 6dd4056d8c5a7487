"""Turns the ncu artefacts a GPU run left in gpurun_out/ into the tracked summaries under profiles/.

  launches_*.csv      (ncu --metrics gpu__time_duration.sum)        -> profiles/<tag>_launches_<name>.md
  *.ncu-rep           (ncu --set full --import-source on)            -> profiles/<tag>_ncu_<name>.md
Usage: python scripts/summarize_profiles.py r1
"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum']


def launches(path, title):
    rows = list(csv.reader(open(path)))
    hdr, start = None, 0
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i + 1
            break
    idx = {h: i for i, h in enumerate(hdr)}
    a = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        if len(r) < len(hdr) or r[idx['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('se::', '')[:70]
        v = float(r[idx['Metric Value']].replace(',', ''))
        unit = r[idx['Metric Unit']]
        v = v / 1000.0 if unit == 'ns' else (v * 1000.0 if unit == 'ms' else v)
        k = (name, r[idx['Grid Size']])
        a[k][0] += 1
        a[k][1] += v
    tot = sum(v[1] for v in a.values())
    out = ['# %s\n' % title,
           '`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares, not absolutes).\n',
           '| kernel | grid | launches | avg us | share |', '|---|---|---:|---:|---:|']
    for k, v in sorted(a.items(), key=lambda kv: -kv[1][1])[:28]:
        out.append('| `%s` | %s | %d | %.2f | %.3f |' % (k[0], k[1], v[0], v[1] / v[0], v[1] / tot))
    out.append('\ntotal device time of the listed launches: %.1f us over %d launches\n' % (tot, sum(v[0] for v in a.values())))
    return '\n'.join(out)


def ncu_report(path, title):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return '# %s\n\n(no kernels in report)\n' % title
    hdr, units = rows[0], rows[1]
    out = ['# %s\n' % title, '`ncu --set full --clock-control none --import-source on`; one column per captured launch.\n']
    name_i = hdr.index('Kernel Name') if 'Kernel Name' in hdr else None
    seen = {}
    for r in rows[2:]:
        nm = r[name_i].split('(')[0] if name_i is not None else 'kernel'
        grid = r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''
        seen.setdefault((nm, grid), r)
    keys = list(seen.keys())[:12]
    out.append('| metric | unit | ' + ' | '.join('`%s` %s' % (k[0].replace('se::', '')[:28], k[1]) for k in keys) + ' |')
    out.append('|---|---|' + '---:|' * len(keys))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append('| %s | %s | ' % (w, units[i]) + ' | '.join(seen[k][i] for k in keys) + ' |')
    # top stall lines of the first kernel
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 3:
        sh = srows[1]
        si = {h: i for i, h in enumerate(sh)}
        stalls = [h for h in sh if h.startswith('stall_') and 'Not Issued' not in h]
        lines, tot = [], 0.0
        for r in srows[2:]:
            if len(r) < len(sh):
                break
            try:
                v = float(r[si['Warp Stall Sampling (All Samples)']])
            except ValueError:
                continue
            tot += v
            top = sorted(((float(r[si[s]]), s) for s in stalls), reverse=True)[0]
            lines.append((v, r[si['Source']].strip()[:70], top[1]))
        lines.sort(reverse=True)
        out.append('\nTop warp-stall sample locations of the first captured kernel (SASS):\n')
        out.append('| share | instruction | dominant stall |\n|---:|---|---|')
        for v, s, st in lines[:10]:
            out.append('| %.1f%% | `%s` | %s |' % (100 * v / max(tot, 1), s.replace('|', '/'), st))
    return '\n'.join(out) + '\n'


# bench.py op-category -> (ncu report, kernel-name substring, grid) for the `roofline.traffic` field
TRAFFIC_MAP = {
    'pairwise_dist': ('r2_kernels', 'pairwise_tc_kernel<0>', None),
    'conv_wgrad 3x3 s1 16->16 @32x32': ('r2_kernels', 'conv_wgrad_pk_kernel', '148'),
    'conv_fwd 3x3 s1 16->16 @32x32': ('r2_kernels', 'conv_tc_kernel<1, 0>', '148'),
    'conv_dgrad 3x3 s1 16->16 @32x32': ('r2_kernels', 'conv_tc_kernel<1, 0>', '148'),
    'conv_wgrad 3x3 s1 32->32 @16x16': ('r2_kernels', 'conv_wgrad_tc_kernel<1>', None),
    # config 4 (ResNet-50, batch 32): the first conv_tc_kernel<1, 1> launch of scripts/ncu_targets.py is the 1x1 forward
    'conv_fwd 1x1 s1 256->1024 @14x14': ('r2_kernels', 'conv_tc_kernel<1, 1>', '148'),
    'conv_dgrad 1x1 s1 256->1024 @14x14': ('r2_kernels', 'conv_tc_kernel<1, 1>', '98'),
    'conv_wgrad 1x1 s1 256->1024 @14x14': ('r2_kernels', 'conv1x1_wgrad_tc_kernel<1>', None),
    'conv_wgrad 3x3 s1 128->128 @28x28': ('r2_kernels', 'conv_wgrad_tc_kernel<1>', '37'),
}


def traffic_json(tag):
    import json
    out = {}
    for cat, (rep, sub, grid) in TRAFFIC_MAP.items():
        path = os.path.join(OUT, rep + '.ncu-rep')
        if not os.path.exists(path):
            continue
        raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ni, gi = hdr.index('Kernel Name'), hdr.index('Grid Size')
        ri, wi = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        for r in rows[2:]:
            if sub in r[ni] and (grid is None or r[gi].strip('() ').split(',')[0] == grid):
                b = float(r[ri].replace(',', '')) * scale[units[ri]] + float(r[wi].replace(',', '')) * scale[units[wi]]
                out[cat] = {'dram_bytes_per_launch': b, 'kernel': r[ni].split('(')[0], 'source': 'profiles/%s_ncu_%s.md' % (tag, rep),
                            'note': 'ncu replays flush the caches: in a training step these tensors are L2-resident'
                                    if 'pairwise' not in sub else 'N=50000, D=100'}
                break
    with open(os.path.join(PROF, '%s_traffic.json' % tag), 'w') as f:
        json.dump(out, f, indent=1)
    return out


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    os.makedirs(PROF, exist_ok=True)
    for p in sorted(glob.glob(os.path.join(OUT, 'launches*.csv'))):
        name = os.path.splitext(os.path.basename(p))[0]
        with open(os.path.join(PROF, '%s_%s.md' % (tag, name)), 'w') as f:
            f.write(launches(p, 'ncu launch list: %s' % name))
    for p in sorted(glob.glob(os.path.join(OUT, '*.ncu-rep'))):
        name = os.path.splitext(os.path.basename(p))[0]
        with open(os.path.join(PROF, '%s_ncu_%s.md' % (tag, name)), 'w') as f:
            f.write(ncu_report(p, 'ncu full capture: %s' % name))
    print(traffic_json(tag))
    print(sorted(os.listdir(PROF)))
