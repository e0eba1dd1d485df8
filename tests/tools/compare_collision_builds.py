import sys, os, subprocess, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
if len(sys.argv) > 1:
    from junction_mpc import synth
    from junction_mpc.batched import BatchedMPC
    from oracle import collision_oracle as C
    res = {}
    for cfg in (3, 4):
        w = synth.make_workload(cfg)
        mpc = BatchedMPC(w["courses"], dl=w["dl"], T=13, max_batch=w["B"])
        margin = C.cutoff_margin(C.CarGeometry(), w["dl"])
        flag, clen = mpc.collision_host(w["agent_idx"], w["state"][:, 2], w["obstacles"], frame_window=w["frame_window"], margin=margin)
        res[f"flag{cfg}"] = flag; res[f"clen{cfg}"] = clen
    np.savez(sys.argv[1], **res)
else:
    # default: arc-length table off vs on; with an argument "head": previous build vs this one
    pairs = (("", "/tmp/coll_old.npz", {"JMPC_NO_ARC_TABLE": "1"}), ("", "/tmp/coll_new.npz", {}))
    for lib, out, extra in pairs:
        env = dict(os.environ); env.update(extra)
        if lib: env["JMPC_LIB"] = lib
        subprocess.check_call([sys.executable, __file__, out], env=env)
    a, b = np.load("/tmp/coll_old.npz"), np.load("/tmp/coll_new.npz")
    for k in a.files: print(k, "identical" if np.array_equal(a[k], b[k]) else "DIFFERENT", a[k].shape, int(a[k].sum()))
