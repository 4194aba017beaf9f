import importlib, sys, numpy as np
sys.path.insert(0, "/root/repo")
gm = importlib.import_module("slam-gmapping-learn_b200"); synth = importlib.import_module("slam-gmapping-learn_b200.synth")
n = 4096
log = synth.make_log("room", synth.UTM30, scans=8, step=0.25, seed=1)
ctx = gm.GmContext(n, pool_patches=n*256); ctx.set_laser(log.laser.angles()); ctx.set_matching(urange=29., range_=30.); ctx.init_particles(n)
rng = np.random.default_rng(0)
for k in range(8):
    poses = log.odom[k][None, :] + rng.normal(0, [0.01, 0.01, 0.005], size=(n, 3))
    ctx.register(log.ranges[k], poses)
st = ctx.stats()
ph = np.array(st["register_phase_cycles"], dtype=float)
print("register ms", st["ms_register"], "phases %", np.round(100*ph/ph.sum(),1), "cycles/CTA", np.round(ph/n))
