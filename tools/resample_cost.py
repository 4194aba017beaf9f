"""Cost of a resampling step at 4096 x 1081: patch-reference remap + the registration that follows (copy-on-write of the
active patches of particles that now share their parent's map)."""
import importlib, sys, numpy as np
sys.path.insert(0, ".")
from oracle import orc
gm = importlib.import_module("slam-gmapping-learn_b200"); synth = importlib.import_module("slam-gmapping-learn_b200.synth")
n = 4096
log = synth.make_log("room", synth.UTM30, scans=12, step=0.25, seed=1)
ctx = gm.GmContext(n, pool_patches=n * 300); ctx.set_laser(log.laser.angles()); ctx.set_matching(urange=29., range_=30.); ctx.init_particles(n)
rng = np.random.default_rng(0)
for k in range(10):
    poses = log.odom[k][None, :] + rng.normal(0, [0.01, 0.01, 0.005], size=(n, 3))
    ctx.register(log.ranges[k], poses)
base = ctx.stats()
w, _ = orc.normalize(rng.normal(0, 600.0, size=n), 1.0)  # a spread that leaves ~1/3 of the particles alive
idx, _ = orc.resample_indexes(w, 0.5)
poses = log.odom[10][None, :] + rng.normal(0, [0.01, 0.01, 0.005], size=(n, 3))
ctx.register(log.ranges[10], poses[idx], idx)
st = ctx.stats()
print({"distinct_parents": int(len(set(idx.tolist()))), "ms_remap": st["ms_remap"], "ms_register_after_resample": st["ms_register"],
       "ms_register_steady": base["ms_register"], "cow_copies": st["cow_copies"],
       "phase_kcycles_per_cta_steady": [round(c / n / 1e3, 1) for c in base["register_phase_cycles"]],
       "phase_kcycles_per_cta_after": [round(c / n / 1e3, 1) for c in st["register_phase_cycles"]], "pool_in_use_before": base["pool_in_use"], "pool_in_use_after": st["pool_in_use"]})
ctx.register(log.ranges[11], poses[idx] + 0.01)
print({"ms_register_next": ctx.stats()["ms_register"], "cow_copies_next": ctx.stats()["cow_copies"]})
