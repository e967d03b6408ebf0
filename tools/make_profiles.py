"""Turns a `ncu --set full` capture of one document pass and the ncu launch list of bench.py into the files under profiles/.
usage: python tools/make_profiles.py <capture.ncu-rep> <launches.csv> <ms_per_pass_from_bench> [tag, default r2] [git sha of the profiled build]"""
import collections, csv, json, os, subprocess, sys
rep, launches, ms = sys.argv[1], sys.argv[2], sys.argv[3]
tag = sys.argv[4] if len(sys.argv) > 4 else 'r2'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sha = sys.argv[5] if len(sys.argv) > 5 else subprocess.run(['git', '-C', root, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, rows = rows[0], rows[1], rows[2:]
keys = {'gpu__time_duration.sum': 'duration_us', 'launch__grid_size': 'grid', 'launch__block_size': 'block', 'launch__registers_per_thread': 'regs',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct', 'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct', 'smsp__inst_executed.sum': 'warp_instructions',
        'sm__inst_issued.avg.pct_of_peak_sustained_active': 'issue_pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'alu_pipe_pct',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'fma_pipe_pct', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'xu_pipe_pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_active': 'l1tex_pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smem_wavefronts', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct'}
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
res = collections.OrderedDict()
for r in rows:
    name = r[hdr.index('Kernel Name')]
    d = {'kernel': name}
    for k, v in keys.items():
        if k in hdr:
            i = hdr.index(k)
            x = r[i].replace(',', '')
            try:
                x = float(x)
            except ValueError:
                pass
            if v in ('dram_read', 'dram_write') and isinstance(x, float):
                x = int(x * scale.get(units[i], 1))
            d[v] = x
    stalls = {}
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
            try:
                val = float(r[i])
            except ValueError:
                continue
            if val >= 0.3:
                stalls[h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')] = round(val, 2)
    d['stalled_warps_per_issue'] = stalls
    key = name.split('(')[0].replace('void ', '')
    if key in res:   # several passes captured: keep the first launch of every kernel
        continue
    res[key] = d
json.dump({'command': 'ncu --set full --clock-control none --import-source on -k regex:stage1_ python tools/quickbench.py 1024  (KERNELS=stream; 1 GiB synthetic document, one document pass)',
           'git_sha': sha,
           'note': f'Times under ncu are cold-cache and serialised (no programmatic dependent launch overlap); the shares are what to compare with the bench ({ms} ms per pass).',
           'kernels': res}, open(f'profiles/{tag}_stream_ncu_summary.json', 'w'), indent=1)
tot = sum(d['duration_us'] for d in res.values())
dram = sum(d['dram_read'] + d['dram_write'] for d in res.values())
for k, d in res.items():
    print('%-45s %8.1f us %5.1f%%  dram %.3f GB alu %.1f issue %.1f l1 %.1f' % (k, d['duration_us'], 100 * d['duration_us'] / tot, (d['dram_read'] + d['dram_write']) / 1e9, d['alu_pipe_pct'], d['issue_pct'], d['l1tex_pct']))
t = json.load(open('profiles/traffic.json'))
alg = t['stream']['algorithmic_bytes_per_pass']
t['stream'].update({'git_sha': sha, 'source': f'profiles/{tag}_stream_ncu_summary.json', 'kernels': ' -> '.join(res.keys()), 'dram_bytes_per_pass': int(dram),
                    'dram_bytes_by_kernel': {k: int(d['dram_read'] + d['dram_write']) for k, d in res.items()},
                    'kernel_shares': {k: round(d['duration_us'] / tot, 4) for k, d in res.items()},
                    'note': 'traffic is %.2fx the algorithmic bytes: classify writes the two structural mask planes (16 B per 64 input bytes), flatten reads one mask plane back (8 B per 64); input read once, every index written once' % (dram / alg)})
json.dump(t, open('profiles/traffic.json', 'w'), indent=1)
print('total us', tot, 'dram', dram, 'x algorithmic', dram / alg)
rows = list(csv.reader(open(launches)))
with open(f'profiles/{tag}_stream_bench_launches.csv', 'w') as f:
    f.write(f'# git {sha}: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1\n')
    f.write('# (the 64 extra stage1_persistent_kernel launches belong to the chunked host path of the e2e leg; the short ones are the no-op fallback)\n')
    w = csv.writer(f)
    for r in rows:
        if len(r) > 10:
            w.writerow([r[0], r[4], r[7], r[8], r[12], r[13], r[14]])
