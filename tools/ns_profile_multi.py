"""Run the sensor-model kernel of every NS bench leg once, warm (for one ncu capture of all field forms):
map.txt 1M particles (field in shared memory via TMA), 1025^2 10M (fp32 through L1/L2), 8193^2 12.5M as fp32 and as one-byte
codes, converged-like (after three filter steps) and freshly uniform.   python tools/ns_profile_multi.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bench import ns_workload
from montecarlolocalisation_b200 import NsShard

CASES = [("map_txt_1M", 6, 1_000_000, 360, 0, 4), ("configs2", 128, 10_000_000, 720, 0, 3), ("grid8192_f32", 1024, 12_500_000, 1080, 16, 5), ("grid8192_u8", 1024, 12_500_000, 1080, 8, 5)]
for name, cells, n, beams, force, seed in CASES:
    if cells == 6:
        from scenario import Scenario
        sc = Scenario(2, n_beams=beams); occ, scans = sc.occ, sc.scans
    else:
        occ, scans = ns_workload(cells, beams, 2, seed)
    s = NsShard(0, 1, n)
    s.pf.setMap(occ, np.float32(0.1))
    s.pf.sampleParticles(n)
    if force:
        s.pf.forceSequential(force)
    sca = scans[0]
    s.pf.stageScan(0, sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
    for i in range(4):          # the 4th update is the converged-like one
        s.pf.updateParticlePos(0.01, 0.02, 0.0)
        mx = s.update_local_staged(0)
        t = s.weights_local(mx)
        s.resample_local(0, t, s.u0())
        s.end_step()
    print(name, s.field_form(), flush=True)
    del s
