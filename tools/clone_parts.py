"""Where GridSlamProcessor::clone() spends its time at 4096 particles: the host copy of the trajectory tree (getTrajectories) against the
whole call (tree + particle vector + device-side clone: directory buffers, pinned staging, one reference per patch)."""
import importlib, json, sys, time
sys.path.insert(0, ".")
hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi"); synth = importlib.import_module("slam-gmapping-learn_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
log = synth.make_log("room", synth.UTM30, scans=12, step=0.25, seed=1)
f = hostapi.Processor(log.laser.angles(), n, log.odom[0], maxUrange=29.0, maxrange=30.0, linearUpdate=0.2, angularUpdate=0.1, srr=0.1, srt=0.2, str=0.1, stt=0.2, ogain=90.0 / n)
for i in range(8):
    f.process_scan(log.ranges[i], log.odom[i], log.stamps[i])
f.check_maps()
def ms(fn, k=4):
    out = []
    for _ in range(k):
        t0 = time.perf_counter(); r = fn(); out.append(round((time.perf_counter() - t0) * 1e3, 3))
    return out, r
tree, leaves = ms(lambda: f.copy_trajectories(False))
nodes = f.copy_trajectories(True)
clones = []
cl, _ = ms(lambda: clones.append(f.clone()))
t0 = time.perf_counter()
for c in clones: c.close()
close = (time.perf_counter() - t0) * 1e3 / len(clones)
print(json.dumps({"particles": n, "tree_nodes": nodes, "tree_copy_and_release_ms": tree, "clone_ms": cl, "close_ms_per_clone": round(close, 3)}))
f.close()
