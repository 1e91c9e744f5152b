"""Turn the raw ncu artefacts of a gpurun profile round into the small tracked summaries under profiles/.
    python tests/tools/summarise_profiles.py <tag>      (reads gpurun_out/<tag>_*, writes profiles/<tag>_*)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tag = sys.argv[1]
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def launches():
    rows = list(csv.reader(open(f'{G}/{tag}_launches.csv')))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi, bi = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Grid Size', 'Block Size'))
    # one complete forward = from one im2col launch up to (not including) the next one
    marks = [i for i, r in enumerate(data) if len(r) > ki and 'im2col_patch16' in r[ki]]
    if len(marks) >= 2:
        data = data[marks[-2]:marks[-1]]
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r'\(.*', '', re.sub(r'^void ', '', r[ki]))[:48]
        a = agg.setdefault((name, r[gi], r[bi]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', ''))
    tot = sum(a[1] for a in agg.values())
    out = ['kernel,grid,block,launches,total_us,avg_us,share']
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f'"{k[0]}","{k[1]}","{k[2]}",{n},{t / 1000:.1f},{t / 1000 / n:.2f},{t / tot:.3f}')
    out.append(f'# total {tot / 1000:.1f} us over {len(data)} launches (one hm_forward of the bench step, eager launches, ncu serialised / cold cache)')
    open(f'{P}/{tag}_launches_summary.csv', 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))


WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors.avg.pct_of_peak_sustained_elapsed']


def full(rep):
    path = f'{G}/{tag}_{rep}.ncu-rep'
    if not os.path.exists(path):
        return
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    stall = [i for i, h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    seen, out = set(), []
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', re.sub(r'^void ', '', r[hdr.index('Kernel Name')]))
        # one entry per distinct launch shape and duration class (proj and fc2 share a kernel and a grid)
        key = (name, r[hdr.index('launch__grid_size')], round(float(r[hdr.index('gpu__time_duration.sum')].replace(',', '')) / 20.0))
        if key in seen:
            continue
        seen.add(key)
        d = {'kernel': name}
        for i in idx:
            d[hdr[i] + ' [' + units[i] + ']'] = r[i]
        top = sorted(((float(r[i].replace(',', '')) if r[i] else 0.0, hdr[i].split('stalled_')[1].split('_per_')[0]) for i in stall), reverse=True)[:4]
        d['top_stalls_per_issue'] = {n: round(v, 2) for v, n in top}
        out.append(d)
    json.dump(out, open(f'{P}/{tag}_ncu_{rep}.json', 'w'), indent=1)
    for d in out:
        print(json.dumps(d)[:400])


if os.path.exists(f'{G}/{tag}_launches.csv'):
    launches()
for rep in ('gemm', 'misc', 'head', 'mano', 'crop'):
    full(rep)
