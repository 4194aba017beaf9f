"""Cost of GridSlamProcessor::clone() at 4096 particles x 1081 beams (gfs-carmen clones the filter after every processed scan,
gfs-carmen.cpp:231-236): the clone is a second set of patch directories over the same patch pool, no patch is copied."""
import importlib, json, sys, time
import numpy as np
sys.path.insert(0, ".")
hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi"); synth = importlib.import_module("slam-gmapping-learn_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
log = synth.make_log("room", synth.UTM30, scans=12, step=0.25, seed=1)
f = hostapi.Processor(log.laser.angles(), n, log.odom[0], maxUrange=29.0, maxrange=30.0, linearUpdate=0.2, angularUpdate=0.1, srr=0.1, srt=0.2, str=0.1, stt=0.2, ogain=90.0 / n)
for i in range(8):
    f.process_scan(log.ranges[i], log.odom[i], log.stamps[i])
used = f.check_maps()
ms = []
clones = []
for k in range(4):
    t0 = time.perf_counter(); c = f.clone(); ms.append((time.perf_counter() - t0) * 1e3); clones.append(c)
assert f.check_maps() == used  # four clones later: the same patches, more owners
t0 = time.perf_counter(); f.process_scan(log.ranges[8], log.odom[8], log.stamps[8]); f.stats(); step_ms = (time.perf_counter() - t0) * 1e3
after = f.check_maps()
t0 = time.perf_counter()
for c in clones:
    c.close()
close_ms = (time.perf_counter() - t0) * 1e3 / len(clones)
print(json.dumps({"particles": n, "clone_ms": [round(x, 3) for x in ms], "patches_referenced_before": used, "after_one_scan_with_4_clones_alive": after,
                  "after_closing_the_clones": f.check_maps(), "scan_ms_with_clones_alive_incl_copy_on_write": round(step_ms, 3), "close_ms_per_clone": round(close_ms, 3),
                  "pool_in_use": f.stats()["pool_in_use"]}))
f.close()
