"""Print a compact summary of a bench.py JSON line: python tools/bench_summary.py <file>"""
import json, sys
for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    print("REF ms/step %.4f value %.3e e2e %.3e (%.4f ms) launches %s clocks %s" % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('ms_per_step', 0), d.get('gpu_launches'), d.get('clocks')))
    r = d['roofline']
    print("  roofline %s achieved %.1f frac %.3f share %.3f" % (r['kernel'], r['achieved'] or 0, r['frac'] or 0, r['share_of_step']))
    for k, v in d.get('kernels', {}).items():
        print('   %-24s %8.1f us x %-3d share %.3f  %s' % (k, v['ms_per_launch'] * 1000, v['launches'], v['share'], ('%.0f GB/s' % v['gbs']) if v.get('gbs') else ''))
    for name, o in (d.get('ns') or {}).items():
        print(' NS %-12s ms/step %.4f value %.3e e2e %.3e (%.4f ms) uniform %s' % (name, o['ms_per_step'], o['value'], o['e2e']['value'], o['e2e']['ms_per_step'],
              ('%.3e' % o['uniform_particles']['evals_per_s_per_gpu']) if 'uniform_particles' in o else None))
        print('    gather', o.get('gather_microbench_reads_per_s'))
        for k, v in o.get('kernels', {}).items():
            print('     %-22s %8.1f us x %-3d share %.3f  %s' % (k, v['ms_per_launch'] * 1000, v['launches'], v['share'], ('%.0f GB/s' % v['gbs']) if v.get('gbs') else ''))
    if 'cpu_baseline' in d:
        print(' cpu', d['cpu_baseline'])
