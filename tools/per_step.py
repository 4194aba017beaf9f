import importlib, sys, numpy as np
sys.path.insert(0, "/root/repo")
gm = importlib.import_module("slam-gmapping-learn_b200"); synth = importlib.import_module("slam-gmapping-learn_b200.synth")
n = 4096
log = synth.make_log("room", synth.UTM30, scans=26, step=0.25, seed=1)
ctx = gm.GmContext(n, pool_patches=n*256); ctx.set_laser(log.laser.angles()); ctx.set_matching(urange=29., range_=30.); ctx.init_particles(n)
rng = np.random.default_rng(0)
poses = np.repeat(log.odom[0][None,:], n, 0)
ctx.register(log.ranges[0], poses)
for k in range(1, 26):
    init = log.odom[k][None, :] + rng.normal(0, [0.01, 0.01, 0.005], size=(n, 3))
    poses, sc, s, l, c, ev = ctx.scanmatch(log.ranges[k], init)
    st1 = ctx.stats()
    ctx.register(log.ranges[k], poses)
    st = ctx.stats()
    ph = np.array(st["register_phase_cycles"], dtype=float)/n/1000
    print(k, "sm %.2f ms evals %.1f | reg %.2f ms active %d tiles %d phases(kcyc/CTA) %s" % (st1["ms_scanmatch"], ev.mean(), st["ms_register"], st["max_active_patches"], st["register_tile_cap"], np.round(ph,1)))
