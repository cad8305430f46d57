"""Run the reference-parity loop for a few steps (for ncu / quick timing): python tools/ref_profile.py <particles> [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import montecarlolocalisation_b200 as m
from scenario import RES, Scenario

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = Scenario(steps, n_beams=360, seed=1)
pf = m.ParticleFilter(max_particles=n, seed=0x1234)
pf.setMap(sc.occ, RES)
pf.sampleParticles(n)
for s in range(steps):
    sca = sc.scans[s]
    pf.stageScan(s, sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
for s in range(steps):
    pf.profileEnable(True)
    pf.diffDriveModel(sc.enc_left[s], sc.enc_right[s])
    pf.computeWeightStaged(s)
    pf.resampleParticles(1)
    pose = pf.estimateWeightedPose()
    print("step %d: " % s + "  ".join("%s %.1f" % (k.replace("k_ref_", "").replace("k_", ""), 1e3 * v[0] / v[1]) for k, v in pf.profileRead().items()), "(us)")
print("pose", pose)
