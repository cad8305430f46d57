"""Run one NS case for a few steps (for ncu / quick timing): python tools/ns_profile.py <cells> <particles> <beams> <uniform 0|1> [steps]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bench import ns_workload
from montecarlolocalisation_b200 import NsShard

cells, n, beams, uniform = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
if cells == 6:
    from scenario import Scenario
    sc = Scenario(2, n_beams=beams); occ, scans = sc.occ, sc.scans
else:
    occ, scans = ns_workload(cells, beams, 2, 4)
s = NsShard(0, 1, n)
s.pf.setMap(occ, np.float32(0.1))
s.pf.sampleParticles(n)
if os.environ.get("NS_FORCE"):
    s.pf.forceSequential(int(os.environ["NS_FORCE"]))      # 8: coded field, 16: fp32 global field, 32: scalar form
sca = scans[0]
s.pf.stageScan(0, sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
for i in range(steps):
    s.pf.profileEnable(True)
    if uniform:
        s.pf.sampleParticles(n)
    s.pf.updateParticlePos(0.01, 0.02, 0.0)
    mx = s.update_local_staged(0)
    t = s.weights_local(mx)
    s.resample_local(0, t, s.u0())
    s.end_step()
    print("step %d [%s]: " % (i, s.field_form()) + "  ".join("%s %.1f us" % (k.replace("k_ns_", ""), 1e3 * v[0] / v[1]) for k, v in s.pf.profileRead().items()))
    if os.environ.get("NS_PROFILE_ANC"):
        anc = s.pf.ancestors()
        print("   distinct ancestors: %d of %d" % (len(np.unique(anc)), n))
fb = occ.size * 4
print("gather bench: smem 9.6KB %.3e reads/s; global %d KB %.3e reads/s; global 64 MiB %.3e; global 1 GiB %.3e" % (
    s.pf.benchGather(0, 9604), fb // 1024, s.pf.benchGather(1, fb), s.pf.benchGather(1, 64 << 20), s.pf.benchGather(1, 1 << 30)))
