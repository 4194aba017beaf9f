"""Synthetic CARMEN-format laser logs.

The generator is ours; the *format* is the one the reference's log reader parses
(reference: openslam_gmapping/log/sensorstream.cpp:68-110 for FLASER lines,
log/carmenconfiguration.cpp:17-95,190-236 for PARAM lines and the beam table).

Two worlds:
  * ``room``     -- 18 x 12 m rectangle with 5 square pillars (1.2 m), robot on an ellipse
                    (6.5 x 4 m semi-axes).  Used by configs C1..C4.
  * ``corridor`` -- square corridor loop (3 m wide), used by the C5 sizing stress.

Both the reference driver (oracle/_ref) and our code read the *text* log, so both
sides see bit-identical doubles (correctly-rounded decimal parsing on both sides).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


# --------------------------------------------------------------------------- worlds
def _rect(cx, cy, w, h):
    x0, x1, y0, y1 = cx - w / 2, cx + w / 2, cy - h / 2, cy + h / 2
    return [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]


def room_segments():
    segs = _rect(0.0, 0.0, 18.0, 12.0)
    for (px, py) in [(0.0, 0.0), (-4.5, 2.2), (4.5, 2.2), (-4.5, -2.2), (4.5, -2.2)]:
        segs += _rect(px, py, 1.2, 1.2)
    return np.asarray(segs, dtype=np.float64)


def corridor_segments(outer=380.0, width=3.0):
    segs = _rect(0.0, 0.0, outer, outer) + _rect(0.0, 0.0, outer - 2 * width, outer - 2 * width)
    # door-frame like ribs every 7.5 m so the corridor is not featureless along its axis
    half = outer / 2
    rib = 0.4
    s = -half + 5.0
    while s < half - 5.0:
        for sign in (-1.0, 1.0):
            segs.append((s, sign * half, s, sign * (half - rib)))
            segs.append((sign * half, s, sign * (half - rib), s))
        s += 7.5
    return np.asarray(segs, dtype=np.float64)


def raycast(segs, ox, oy, angles, max_range):
    """Distance from (ox,oy) along each angle to the nearest segment (numpy, vectorised)."""
    dx = np.cos(angles)[:, None]
    dy = np.sin(angles)[:, None]
    x1, y1, x2, y2 = segs[:, 0][None, :], segs[:, 1][None, :], segs[:, 2][None, :], segs[:, 3][None, :]
    ex, ey = x2 - x1, y2 - y1
    den = dx * ey - dy * ex
    with np.errstate(divide="ignore", invalid="ignore"):
        t = ((x1 - ox) * ey - (y1 - oy) * ex) / den
        u = ((x1 - ox) * dy - (y1 - oy) * dx) / den
    ok = (np.abs(den) > 1e-12) & (t > 1e-9) & (u >= 0.0) & (u <= 1.0)
    t = np.where(ok, t, np.inf)
    r = t.min(axis=1)
    return np.where(np.isfinite(r), np.minimum(r, max_range), max_range)


# --------------------------------------------------------------------------- log
@dataclass
class LaserSpec:
    beams: int
    res_deg: float
    max_range: float  # sensor max range (a reading at max range means "no return")

    def angles(self):
        """Beam table exactly as carmenconfiguration.cpp:218-236 builds it (cumulative +=)."""
        n = self.beams
        out = np.zeros(n, dtype=np.float64)
        center = n / 2.0
        low = int(math.floor(center))
        up = int(math.ceil(center))
        step = self.res_deg * math.pi / 180.0
        angle = 0.0 if n % 2 else step
        i = 0 if n % 2 else 1
        while i < low + 1:
            out[low - i] = -angle
            out[up + i - 1] = angle
            i += 1
            angle += step
        return out


LMS200 = LaserSpec(181, 1.0, 81.0)
UTM30 = LaserSpec(1081, 0.25, 30.0)


@dataclass
class SynthLog:
    laser: LaserSpec
    ranges: np.ndarray  # [S, B] float64 (already rounded to the printed precision)
    odom: np.ndarray  # [S, 3]
    truth: np.ndarray  # [S, 3]
    stamps: np.ndarray  # [S]
    meta: dict = field(default_factory=dict)


def make_log(world="room", laser=LMS200, scans=100, step=0.25, seed=1, noise=0.01,
             drift=(2e-4, -1e-4, 2e-4)):
    """Seeded synthetic trajectory + scans.  Ranges/poses are rounded to the precision the
    text log prints so that write->parse is the identity."""
    rng = np.random.default_rng(seed)
    if world == "room":
        segs = room_segments()
        a, b = 6.5, 4.0
        # ellipse parameter advanced so that arc length per step ~= `step`
        ts = np.zeros(scans)
        t = 0.0
        for k in range(scans):
            ts[k] = t
            ds = math.hypot(a * math.sin(t), b * math.cos(t))
            t += step / ds
        x = a * np.cos(ts)
        y = b * np.sin(ts)
        th = np.arctan2(b * np.cos(ts), -a * np.sin(ts))
    elif world == "corridor":
        segs = corridor_segments()
        half = 380.0 / 2 - 1.5
        per = 8 * half
        s = (np.arange(scans) * step) % per
        x = np.zeros(scans); y = np.zeros(scans); th = np.zeros(scans)
        for k, sk in enumerate(s):
            side = int(sk // (2 * half)); off = sk - side * 2 * half - half
            if side == 0: x[k], y[k], th[k] = off, -half, 0.0
            elif side == 1: x[k], y[k], th[k] = half, off, math.pi / 2
            elif side == 2: x[k], y[k], th[k] = -off, half, math.pi
            else: x[k], y[k], th[k] = -half, -off, -math.pi / 2
    else:
        raise ValueError(world)
    truth = np.stack([x, y, th], axis=1)
    # start the odometric frame at the first true pose => particles start at initialPose (0,0,0)
    # the filter's map frame equals the odom frame at t0; keep absolute coords (gfs drivers do too).
    k = np.arange(scans, dtype=np.float64)[:, None]
    odom = truth + k * np.asarray(drift)[None, :]
    odom[:, 2] = np.arctan2(np.sin(odom[:, 2]), np.cos(odom[:, 2]))
    angles = laser.angles()
    ranges = np.zeros((scans, laser.beams))
    for i in range(scans):
        r = raycast(segs, x[i], y[i], th[i] + angles, laser.max_range)
        r = r + rng.normal(0.0, noise, size=r.shape) * (r < laser.max_range)
        ranges[i] = np.clip(r, 0.02, laser.max_range)
    ranges = np.round(ranges, 3)
    odom = np.round(odom, 6)
    stamps = np.round(1000.0 + 0.5 * np.arange(scans), 3)
    return SynthLog(laser, ranges, odom, truth, stamps,
                    dict(world=world, step=step, seed=seed, noise=noise, drift=list(drift)))


def write_carmen(log: SynthLog, path):
    """CARMEN text log.  FLASER layout: `FLASER B r0..rB-1 lx ly lth ox oy oth t host t`
    (second pose triple becomes reading.getPose(), sensorstream.cpp:96-99)."""
    with open(path, "w") as f:
        f.write("# synthetic CARMEN log (gm_b200 synth.py)\n")
        f.write("PARAM robot_frontlaser_offset 0.0\n")
        f.write("PARAM laser_front_laser_resolution %.6f\n" % log.laser.res_deg)
        for i in range(log.ranges.shape[0]):
            o = log.odom[i]
            t = log.stamps[i]
            f.write("ODOM %.6f %.6f %.6f 0.0 0.0 0.0 %.3f synth %.3f\n" % (o[0], o[1], o[2], t, t))
            rs = " ".join("%.3f" % v for v in log.ranges[i])
            line = "FLASER %d %s %.6f %.6f %.6f %.6f %.6f %.6f %.3f synth %.3f\n" % (
                log.laser.beams, rs, o[0], o[1], o[2], o[0], o[1], o[2], t, t)
            assert len(line) < 8191, "FLASER line exceeds the reference parser's 8191-char buffer"
            f.write(line)


def read_carmen(path):
    """Parse the subset of CARMEN we write (FLASER + resolution PARAM)."""
    ranges, odom, stamps = [], [], []
    res = 1.0
    beams = None
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "PARAM" and tok[1] == "laser_front_laser_resolution":
                res = float(tok[2])
            elif tok[0] == "FLASER":
                n = int(tok[1])
                beams = n
                ranges.append([float(v) for v in tok[2:2 + n]])
                p = tok[2 + n + 3:2 + n + 6]
                odom.append([float(v) for v in p])
                stamps.append(float(tok[2 + n + 6]))
    if beams in (180, 181):
        res = 1.0
    elif beams in (360, 361, 540, 541):
        res = 0.5
    laser = LaserSpec(beams, res, float(np.max(ranges)))
    return SynthLog(laser, np.asarray(ranges), np.asarray(odom), np.asarray(odom), np.asarray(stamps), {})
