"""Known-answer tests for the CPU oracle: the facts SURVEY.md App. B extracted from the reference, re-verified."""
import hashlib

import numpy as np

from oracle.pyoracle import Oracle, Scan


def test_map_txt_rasterisation_kat(map_txt):
    occ = Oracle.rasterise_map_txt(map_txt)
    assert occ.shape == (49, 49)                       # 6*8+1 (publish_map_rviz.cpp:330-331)
    assert int((occ == 100).sum()) == 375
    assert set(np.unique(occ)) == {0, 100}
    assert hashlib.sha256(occ.tobytes()).hexdigest() == "9d700e0d21c8b669621222f4c2514d9d80a7e502fc476f77120c348c4849c275"


def test_ragged_and_bad_maps():
    # a short second row: the missing cell is filled solid (publish_map_rviz.cpp:395-408)
    occ = Oracle.rasterise_map_txt("[[[T,L],[T,R]],[[L,B]]]")
    assert occ.shape == (17, 17)
    assert (occ[8:16, 8:17] == 100).all()
    for bad in ("", "[[[X]]]", "[[T]]", "[[[T]]"):
        try:
            Oracle.rasterise_map_txt(bad)
            assert False, bad
        except ValueError:
            pass


def test_ray_direction_prefill_bug_keys(map_txt):
    o = Oracle()
    o.precompute_ray_directions(-120.0, 120.0, 0.1)
    k_all, _, _ = o.ray_lut()
    assert len(k_all) == 2401                                   # 2401 iterations, 2401 distinct keys
    k, dx, dy = o.ray_lut(-300, 300)
    expect = list(range(-300, 0, 10)) + [0] + list(range(9, 300, 10))
    assert list(k) == expect and len(k) == 61                   # Q9: keys are (int)(a*100), looked up at 1 degree
    i = list(k).index(-90)                                      # key -90 holds the direction of -0.9 degrees
    assert abs(dx[i] - np.cos(np.deg2rad(-0.9))) < 1e-12 and abs(dy[i] - np.sin(np.deg2rad(-0.9))) < 1e-12


def test_gauss_lut_constants():
    o = Oracle()
    t = o.gauss_table()
    assert len(t) == 11001
    assert t.max() == 3.9894228040143269 == t[0]
    assert o.gauss_get(-1e-9) == 0.0 and o.gauss_get(1.2) == 0.0
    assert o.gauss_get(1.1000000238418579) == t[-1]            # max_diff itself is inside
    # linear interpolation between entries, evaluated without FMA
    d = 0.123456
    res = float(np.float32(0.0001))
    idxf = d / res
    i = int(idxf)
    w = idxf - i
    assert o.gauss_get(d) == (1.0 - w) * t[i] + w * t[i + 1]


def test_beam_counts_after_fov_and_stride():
    for B, kept, used in ((360, 240, 12), (720, 480, 24), (1080, 720, 36)):
        sc = Scan(np.full(B, 1.0, np.float32), -np.pi, 2 * np.pi / B, 0.02, 5.6)
        r, a = Oracle.filter_scan(sc)
        assert len(r) == kept
        assert len(range(0, kept, 20)) == used
    # robot-like scan: 683 beams at 0.352 degrees from -120 degrees
    inc = np.deg2rad(0.352)
    sc = Scan(np.full(683, 1.0, np.float32), np.deg2rad(-120.0), inc, 0.02, 5.6)
    r, a = Oracle.filter_scan(sc)
    assert len(r) == 681 and len(range(0, 681, 20)) == 35


def test_scan_filter_semantics():
    ranges = np.array([0.5, np.nan, np.inf, 0.01, 6.0, 1.0], np.float32)
    sc = Scan(ranges, 0.0, 0.01, 0.02, 5.6)
    r, a = Oracle.filter_scan(sc)
    # NaN/Inf -> 1.05; out-of-range finite readings are dropped so later beams shift (Q5)
    assert list(r) == [0.5, 1.05, 1.05, 1.0]
    assert np.allclose(a, np.float64(np.float32(0.01)) * np.array([0, 1, 2, 5]))


def test_ray_march_radii_and_truncation_quirk(map_txt):
    o = Oracle()
    occ = Oracle.rasterise_map_txt(map_txt)
    o.set_map(occ)
    # Q7: a point up to one cell below the low edge still maps to cell 0
    assert o.is_occupied(-0.05, -0.05) is True and o.is_occupied(-0.11, 0.0) is False
    # Q10: resolution is float32(0.1): 0.8/res truncates to 7, and column 7 of row 1 is free, column 8 is a wall
    assert occ[1, 7] == 0 and occ[1, 8] == 100
    assert o.is_occupied(0.8, 0.15) is False
    # ray along +x from an open spot: the hit distance is one of the 11 accumulated radii or max_range
    radii = [0.0]
    while radii[-1] + 0.1 < 1.0 or len(radii) < 11:
        radii.append(radii[-1] + 0.1)
    assert radii[10] == 0.99999999999999989 and len(radii) == 11
    o.precompute_ray_directions()
    got = {o.raycast(x, 2.05, 0.3, 0.0) for x in np.linspace(0.3, 4.5, 60)}
    assert got <= set(radii) | {1.0}


def test_resample_edge_total_weight_zero(map_txt):
    """total weight 0 -> NaN weights -> every lower_bound returns index 0, p_inject = max(0, NaN) = 0."""
    o = Oracle()
    o.set_map(Oracle.rasterise_map_txt(map_txt))
    o.precompute_ray_directions()
    n = 64
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = -5.0            # all outside the map -> weight 0
    sc = Scan(np.full(360, 0.7, np.float32), -np.pi, 2 * np.pi / 360, 0.02, 5.6)
    u = np.linspace(0.01, 0.99, n)
    out, idx, cdf, st = o.resample(P, 1, sc, u, np.full(3 * n, 0.5))
    assert st["total_weight"] == 0.0 and st["p_inject"] == 0.0 and st["injected"] == 0
    assert np.isnan(cdf).all() and (idx == 0).all()
