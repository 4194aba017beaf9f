"""Reader for the binary traces written by oracle/_ref/ref_driver (-mode trace) -- TEST INFRASTRUCTURE.

Record = 4-byte tag, u64 byte count, payload.  Returns a list of per-reading dicts.
"""
from __future__ import annotations

import struct

import numpy as np

from .orc import DIGEST_DTYPE


def read_trace(path):
    with open(path, "rb") as f:
        buf = f.read()
    off = 0
    head = None
    angles = None
    readings = []
    cur = None
    final = {}
    while off < len(buf):
        tag = buf[off:off + 4].decode()
        n = struct.unpack_from("<Q", buf, off + 4)[0]
        p = buf[off + 12:off + 12 + n]
        off += 12 + n
        if tag == "HEAD":
            h = np.frombuffer(p, dtype="<f8")
            head = dict(particles=int(h[0]), beams=int(h[1]), init=h[2:5].copy(), delta=float(h[5]), seed=int(h[6]))
        elif tag == "ANGL":
            angles = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "READ":
            r = np.frombuffer(p, dtype="<f8")
            cur = dict(odom=r[:3].copy(), stamp=float(r[3]))
            readings.append(cur)
        elif tag == "RNGS":
            cur["ranges"] = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "ODOM":
            cur["pre_pose"] = np.frombuffer(p, dtype="<f8").reshape(-1, 3).copy()
        elif tag == "PROC":
            fl = np.frombuffer(p, dtype="<i4")
            cur["processed"] = bool(fl[0]); cur["resampled"] = bool(fl[1])
        elif tag == "SMAT":
            # init xyz, opt score, corrected xyz, s, l, c, pose xyz, weight, weightSum
            cur["smat"] = np.frombuffer(p, dtype="<f8").reshape(-1, 15).copy()
        elif tag == "PROB":
            cur["probes"] = np.frombuffer(p, dtype="<f8").reshape(-1, 8).copy()
        elif tag == "LIKE":
            cur["like"] = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "RNEF":
            cur["neff_at_resample"] = float(np.frombuffer(p, dtype="<f8")[0])
        elif tag == "RWGT":
            cur["w_at_resample"] = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "RIDX":
            cur["indexes"] = np.frombuffer(p, dtype="<u4").copy()
        elif tag == "STAT":
            cur["state"] = np.frombuffer(p, dtype="<f8").reshape(-1, 6).copy()
        elif tag == "NEFF":
            cur["neff"] = float(np.frombuffer(p, dtype="<f8")[0])
        elif tag == "WGTS":
            cur["weights"] = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "TREE":
            # per particle: chain length, leaf accWeight, sum accWeight, position-weighted sum, sum visitCounter, sum childs
            cur["tree"] = np.frombuffer(p, dtype="<f8").reshape(-1, 6).copy()
        elif tag == "MDIG":
            d = np.frombuffer(p, dtype=DIGEST_DTYPE).copy()
            (cur if cur is not None else final)["digests"] = d
            final["digests"] = d
        elif tag == "MGEO":
            g = np.frombuffer(p, dtype="<i4").reshape(-1, 6).copy()
            cur["geometry"] = g
            final["geometry"] = g
        elif tag == "MCEL":
            a = np.frombuffer(p, dtype="<i4")
            final["best"] = int(a[0])
            final["best_cells"] = a[1:].reshape(-1, 6).copy()  # wx wy n visits accx_bits accy_bits
        elif tag == "TRAJ":  # getTrajectories(): per leaf depth, sum x y theta, sum accWeight, sum childs, readings, shared, aliasing
            final["trajectories"] = np.frombuffer(p, dtype="<f8").reshape(-1, 9).copy()
        elif tag == "UMAP":  # the ROS node's map rebuild from the best trajectory (ref_driver.cpp dumpRosMap)
            final["ros_map_digest"] = np.frombuffer(p, dtype=DIGEST_DTYPE).copy()
        elif tag == "UINF":  # nodes, map size x y, world min x y, world max x y, unknown / free / occupied cells, grid hash hi lo
            final["ros_map_info"] = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "SMPP":  # ScanMatcherProcessor: pose after every reading
            final["smp_poses"] = np.frombuffer(p, dtype="<f8").reshape(-1, 3).copy()
        elif tag == "SMPD":
            final["smp_digest"] = np.frombuffer(p, dtype=DIGEST_DTYPE).copy()
        elif tag == "SMPG":
            final["smp_geometry"] = np.frombuffer(p, dtype="<i4").copy()
        elif tag == "BBOX":  # SensorLog: readings, boundingBox xmin ymin xmax ymax, first range pose
            final["sensorlog_bbox"] = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "ICPO":  # ScanMatcher::icpOptimize / icpStep from two poses: (pose xyz, score) of each
            final["icp"] = np.frombuffer(p, dtype="<f8").reshape(-1, 8).copy()
        elif tag == "RSEN":  # ScanMatcher::registerScan return values (entropy change) of two standalone registrations
            final["register_gain"] = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "ISEQ":  # integrateScanSequence on a second filter: pose xyz, weight, weightSum, depth, leaf pose xyz, map geometry (6), path nodes
            final["integrate"] = np.frombuffer(p, dtype="<f8").reshape(-1, 16).copy()
        elif tag == "DONE":
            pass
        else:
            raise ValueError("unknown tag %r" % tag)
    return dict(head=head, angles=angles, readings=readings, final=final)
