"""Deterministic synthetic "64-beam lidar in a box" sweeps: the benchmark's input.

Sensor pose of sweep k is (0.5 k, 0, 1.8) m; 64 beams with elevation -24.8 deg .. +2.0 deg, 2048 azimuth steps;
the scene is the axis-aligned room x in [-40, 40 + 0.5 K], y in [-30, 30], z in [0, 12]; range noise N(0, 0.02 m)
from MT19937(1153297050); returns beyond 60 m are dropped.  Rays are stored in firing order
(index = azimuth * 64 + beam), laid out exactly as RayMapper::integrateRays expects: [origin0, sample0, ...] f64.
"""
import numpy as np

BEAMS = 64
AZIMUTHS = 2048
RAYS_PER_SWEEP = BEAMS * AZIMUTHS
SEED = 1153297050
MAX_RANGE = 60.0


def _directions():
    elev = np.deg2rad(-24.8 + np.arange(BEAMS) * (26.8 / 63.0))
    azim = 2.0 * np.pi * np.arange(AZIMUTHS) / AZIMUTHS
    ce, se = np.cos(elev), np.sin(elev)
    d = np.empty((AZIMUTHS, BEAMS, 3))
    d[..., 0] = np.cos(azim)[:, None] * ce[None, :]
    d[..., 1] = np.sin(azim)[:, None] * ce[None, :]
    d[..., 2] = se[None, :]
    return d.reshape(-1, 3)


class LidarBox:
    def __init__(self, sweeps=1, seed=SEED):
        self.sweeps = sweeps
        self.rng = np.random.RandomState(seed % (2 ** 32))
        self.dirs = _directions()
        self.lo = np.array([-40.0, -30.0, 0.0])
        self.hi = np.array([40.0 + 0.5 * sweeps, 30.0, 12.0])
        self._next = 0

    def sweep(self, k=None):
        """Returns (rays (2n,3) f64, intensities (n,) f32, timestamps (n,) f64) of sweep k (sweeps must be drawn
        in order for the noise stream to be reproducible)."""
        if k is None:
            k = self._next
        if k != self._next:
            raise ValueError("sweeps must be generated in order")
        self._next += 1
        origin = np.array([0.5 * k, 0.0, 1.8])
        d = self.dirs
        with np.errstate(divide="ignore"):
            t_hi = (self.hi - origin) / d
            t_lo = (self.lo - origin) / d
        t = np.where(d > 0, t_hi, np.where(d < 0, t_lo, np.inf)).min(axis=1)
        t = t + self.rng.normal(0.0, 0.02, size=t.shape)
        intensities = self.rng.uniform(0.0, 255.0, size=t.shape).astype(np.float32)
        idx = np.arange(RAYS_PER_SWEEP)
        timestamps = 0.1 * k + 0.1 * idx / RAYS_PER_SWEEP
        keep = t <= MAX_RANGE
        samples = origin[None, :] + d[keep] * t[keep, None]
        rays = np.empty((2 * samples.shape[0], 3))
        rays[0::2] = origin
        rays[1::2] = samples
        return rays, intensities[keep], timestamps[keep]


def cube_rays(count=10000, half_extent=3.1, origin=(0.05, 0.05, 0.05), seed=5489):
    """BASELINE config 1: rays from a fixed sensor to uniform samples inside a cube (single region at 0.2 m)."""
    rng = np.random.RandomState(seed)
    rays = np.empty((2 * count, 3))
    rays[0::2] = np.asarray(origin)
    rays[1::2] = rng.uniform(-half_extent, half_extent, size=(count, 3))
    return rays
