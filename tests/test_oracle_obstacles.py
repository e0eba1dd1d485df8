"""The scripted-obstacle restatement (oracle/obstacle_oracle.py) against tracks recorded from the reference's own
classes (tests/golden/scripted_obstacles.npz, made by tests/golden/make_golden.py --obstacles)."""
import os

import numpy as np

from oracle import obstacle_oracle as OB


def test_scripted_obstacles_reproduce_the_reference_tracks(golden_dir):
    z = np.load(os.path.join(golden_dir, "scripted_obstacles.npz"))
    assert len(z["specs"]) == 51
    for spec, track in zip(z["specs"], z["tracks"]):
        o = OB.from_spec(spec)
        for row in track:
            assert np.array_equal(np.array(o.get(), float), row), spec
            o.step()


def test_device_script_rows_describe_the_same_obstacles(golden_dir):
    """junction_mpc.episodes.scripted_obstacles (host-side packing, no GPU needed) against the fixture's specs."""
    from junction_mpc.episodes import scripted_obstacles
    z = np.load(os.path.join(golden_dir, "scripted_obstacles.npz"))
    kinds = {1: "t_intersection", 2: "roundabout", 3: "arterial"}
    rows = [[dict(kind=kinds[int(s[0])], direction=int(s[1]), turning=bool(s[2]), speed=float(s[3]),
                  offset=None if s[4] < 0 else float(s[4]), dt=float(s[5]), x_init=float(s[6]), y_init=float(s[7]),
                  initial_speed=float(s[8]))] for s in z["specs"]]
    script, model = scripted_obstacles(rows)
    assert script.shape == (51, 1, 8) and model.shape == (51, 1, 4)
    # initial pose = first recorded get() tuple
    assert np.array_equal(model[:, 0, :2], z["tracks"][:, 0, :2]) and np.array_equal(model[:, 0, 2], z["tracks"][:, 0, 3])
    assert (script[z["specs"][:, 0] == 2, 0, 6] == 0.2).all()          # the roundabout's dt quirk
