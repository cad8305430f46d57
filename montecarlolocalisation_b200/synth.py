"""Deterministic synthetic workloads for the configs in BASELINE.json / SURVEY.md §8d.

Host-side numpy only (no oracle, no CUDA): occupancy grids, 360/720/1080-beam LIDAR scans ray-cast from a
ground-truth pose, and wheel-encoder traces. The PRNG is SplitMix64 so inputs are identical on every machine.
"""
import numpy as np

_M64 = (1 << 64) - 1


class SplitMix64:
    """Vectorised SplitMix64: state advances by the golden-gamma per draw."""

    def __init__(self, seed):
        self.state = int(seed) & _M64

    def u64(self, n):
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = np.uint64(self.state) + idx * np.uint64(0x9E3779B97F4A7C15)
            self.state = int(z[-1]) if n else self.state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n):
        """n canonical doubles in [0,1) with 53 random bits."""
        return (self.u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def normal(self, n):
        m = (n + 1) // 2
        u1 = 1.0 - self.uniform(m)
        u2 = self.uniform(m)
        r = np.sqrt(-2.0 * np.log(u1))
        z = np.concatenate([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)])
        return z[:n]

    def integers(self, n, hi):
        return (self.u64(n) % np.uint64(hi)).astype(np.int64)


def maze_occupancy(cells, seed, cell_px=8, p_wall=0.35):
    """cells x cells maze, each cell cell_px pixels, walls one pixel thick on cell boundaries; every interior
    wall segment present with probability p_wall, border closed. Returns int8 (cells*cell_px+1)^2, 100 = wall."""
    rng = SplitMix64(seed)
    n = cells * cell_px + 1
    occ = np.zeros((n, n), np.int8)
    horiz = rng.uniform((cells + 1) * cells).reshape(cells + 1, cells) < p_wall   # wall above cell (r,c)
    vert = rng.uniform(cells * (cells + 1)).reshape(cells, cells + 1) < p_wall    # wall left of cell (r,c)
    horiz[0, :] = horiz[-1, :] = True
    vert[:, 0] = vert[:, -1] = True
    hrows = np.repeat(horiz, cell_px, axis=1)            # (cells+1, cells*cell_px)
    occ[::cell_px, :-1] |= np.where(hrows, 100, 0).astype(np.int8)
    occ[::cell_px, 1:] |= np.where(hrows, 100, 0).astype(np.int8)
    vcols = np.repeat(vert, cell_px, axis=0)             # (cells*cell_px, cells+1)
    occ[:-1, ::cell_px] |= np.where(vcols, 100, 0).astype(np.int8)
    occ[1:, ::cell_px] |= np.where(vcols, 100, 0).astype(np.int8)
    return occ


def raycast_true(occ, res, x, y, angles, max_range):
    """Ground-truth ranges: march each ray at res/8 steps until an occupied cell (>50) or leaving the grid."""
    h, w = occ.shape
    step = res / 8.0
    n_steps = int(max_range / step) + 1
    r = np.arange(n_steps, dtype=np.float64) * step
    px = x + np.cos(angles)[:, None] * r[None, :]
    py = y + np.sin(angles)[:, None] * r[None, :]
    mx = np.floor(px / res).astype(np.int64)
    my = np.floor(py / res).astype(np.int64)
    inside = (mx >= 0) & (my >= 0) & (mx < w) & (my < h)
    hit = np.zeros_like(inside)
    hit[inside] = occ[my[inside], mx[inside]] > 50
    stop = hit | ~inside
    first = np.where(stop.any(axis=1), stop.argmax(axis=1), n_steps - 1)
    out = r[first]
    left = ~inside[np.arange(len(angles)), first]
    out[left] = np.inf
    return out


def make_scan(occ, res, pose, n_beams, seed, range_min=0.02, range_max=5.6, noise_sigma=0.01, p_nan=0.05,
              laser_offset=0.1, angle_min=None, angle_inc=None):
    """A scan seen from `pose`: full circle by default (angle_min=-pi, inc=2pi/B as float32), or the given float32
    angle_min / angle_inc (e.g. the robot's own 683 beams at 0.352 degrees from -120 degrees, comment MC:638-640). The sensor
    sits laser_offset ahead of the robot and its beam angles are mirrored, matching how the reference projects them
    (MC:644,653)."""
    rng = SplitMix64(seed)
    angle_min = np.float32(-np.pi) if angle_min is None else np.float32(angle_min)
    angle_inc = np.float32(2 * np.pi / n_beams) if angle_inc is None else np.float32(angle_inc)
    beam = np.float64(angle_min) + np.arange(n_beams) * np.float64(angle_inc)
    x, y, th = pose
    lx, ly = x + laser_offset * np.cos(th), y + laser_offset * np.sin(th)
    true = raycast_true(occ, res, lx, ly, th - beam, range_max + 1.0)
    ranges = true + noise_sigma * rng.normal(n_beams)
    ranges = np.where(np.isfinite(ranges), ranges, np.inf).astype(np.float32)
    ranges[rng.uniform(n_beams) < p_nan] = np.nan
    return dict(ranges=ranges, angle_min=angle_min, angle_inc=angle_inc, range_min=np.float32(range_min),
                range_max=np.float32(range_max))


WHEEL_SIZE = 0.0620    # PID_lib.hpp:20
WHEEL_SPACE = 0.265    # PID_lib.hpp:19


def encoder_trace(n_steps, straight=0.5, turn=0.3, turn_every=10):
    """Cumulative wheel-encoder angles (rad): equal increments when driving straight, a differential every
    turn_every-th step. Returns (left[n_steps], right[n_steps])."""
    dl = np.full(n_steps, straight)
    dr = np.full(n_steps, straight)
    k = np.arange(n_steps)
    t = (k % turn_every) == turn_every - 1
    sign = np.where((k // turn_every) % 2 == 0, 1.0, -1.0)
    dl[t] += turn * sign[t]
    dr[t] -= turn * sign[t]
    return np.cumsum(dl), np.cumsum(dr)


def integrate_odometry(left, right, start_pose):
    """Ground-truth poses from an encoder trace with the reference's midpoint model (MC:719-731)."""
    x, y, th = start_pose
    poses = []
    pl = pr = 0.0
    for l, r in zip(left, right):
        d_l = (l - pl) * WHEEL_SIZE * 0.5
        d_r = (r - pr) * WHEEL_SIZE * 0.5
        d_c = 0.5 * (d_l + d_r)
        dth = (d_l - d_r) / WHEEL_SPACE
        x += d_c * np.cos(th + 0.5 * dth)
        y += d_c * np.sin(th + 0.5 * dth)
        th = np.arctan2(np.sin(th + dth), np.cos(th + dth))
        pl, pr = l, r
        poses.append((x, y, th))
    return poses


def uniform_particles(n, extent_x, extent_y, seed):
    """n particles uniform over the map extent, theta uniform in [-pi,pi), weight 1. float32 [n,4]."""
    rng = SplitMix64(seed)
    P = np.empty((n, 4), np.float32)
    P[:, 0] = rng.uniform(n) * extent_x
    P[:, 1] = rng.uniform(n) * extent_y
    P[:, 2] = rng.uniform(n) * (2 * np.pi) - np.pi
    P[:, 3] = 1.0
    return P
