"""Print a compact summary of a bench.py JSON line: python tools/bench_summary.py <file>"""
import json, sys


def ktable(kernels, ind="   "):
    for k, v in kernels.items():
        print(ind + '%-24s %8.1f us x %-3d share %.3f  %s' % (k, v['ms_per_launch'] * 1000, v['launches'], v['share'], ('%.0f GB/s' % v['gbs']) if v.get('gbs') else ''))


def leg(name, o):
    print(' %-16s ms/step %.4f value %.3e e2e %.3e (%.4f ms) launches/step %s %s' % (name, o['ms_per_step'], o['value'], o['e2e']['value'], o['e2e']['ms_per_step'],
          o.get('launches_per_step'), ('uniform %.3e' % o['uniform_particles']['evals_per_s_per_gpu']) if 'uniform_particles' in o else ''))
    if 'roofline' in o:
        r = o['roofline']
        print('    roofline %s bound %s frac %s hbm_frac %s share %.3f' % (r['kernel'], r['bound'], r['frac'], r.get('hbm_frac'), r['share_of_step']))
    if 'update_skew' in o:
        print('    skew', o['update_skew'], o.get('kernel_us_per_rank'))
    if 'kernels' in o:
        ktable(o['kernels'], '     ')


for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    print("N=%d %s\n  ms/step %.4f value %.3e e2e %.3e (%.4f ms) launches/step %s gate %s clocks %s" % (
        d['n_gpus'], d['config']['workload'][:60], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('ms_per_step', 0), d.get('launches_per_step'),
        d.get('parity_gate'), d.get('clocks')))
    r = d['roofline']
    print("  roofline %s bound %s frac %s hbm_frac %s share %.3f" % (r['kernel'], r['bound'], r['frac'], r.get('hbm_frac'), r['share_of_step']))
    ktable(d.get('kernels', {}))
    for g, v in (d.get('parity_gates') or {}).items():
        print("  gate %-16s %s %s" % (g, v['status'], v.get('hash', '')))
    for r in d.get('reference_sized', []):
        print("  ref-sized N=%d %2d beams: GPU %.1f us/tick (%.1f launches)  CPU %.3f ms/tick  x%.1f" % (r['particles'], r['scored_beams'], r['gpu_us_per_tick'], r['gpu_launches_per_tick'],
              r['cpu_ms_per_tick'], r['gpu_over_cpu']))
    for name, o in (d.get('ns') or {}).items():
        leg(name, o)
    for name, o in (d.get('ns_weak') or {}).items():
        leg(name, o)
    if 'ref_replicas' in d:
        o = d['ref_replicas']
        print(' ref_replicas ms/step %.4f value %.3e e2e %.4f ms' % (o['ms_per_step'], o['value'], o['e2e']['ms_per_step']))
    if 'cpu_baseline' in d:
        print(' cpu', d['cpu_baseline'])
