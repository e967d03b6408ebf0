"""One line per bench log under gpurun_out/ (or the directory given): the figures DESIGN.md / BASELINE.md quote."""
import glob, json, os, sys
d = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out'
for f in sorted(glob.glob(os.path.join(d, 'bench_final_*.log')) + glob.glob(os.path.join(d, 'bench_final8_*.log'))):
    for line in open(f):
        if not line.startswith('{'):
            continue
        j = json.loads(line)
        r = j.get('roofline') or {}
        print('%-34s value %9.2f %s  ms %.4f  frac %s  e2e %s  launches %s  segs %s  lat %s  cpu %s' % (
            os.path.basename(f), j['value'], j['unit'], j['ms_per_step'], r.get('frac'), (j.get('e2e') or {}).get('value'), j.get('gpu_launches'),
            j.get('segments_per_gpu'), j.get('latency_us'), (j.get('cpu_baseline') or {}).get('value')))
