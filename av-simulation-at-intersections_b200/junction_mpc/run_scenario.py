"""Run one of the reference's scenario scripts, unmodified, with the B200 controller swapped in.

    python -m junction_mpc.run_scenario /path/to/reference/main/scenarios/mpc_intersection.py

The script's own `from lib.mpc import MPC, MAX_ACCEL` then resolves to junction_mpc.mpc (see install()); every
other import (`lib.simulation`, `lib.collision_avoidance`, `envs.*`, matplotlib, ...) stays the reference's.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    main_dir = os.path.dirname(os.path.dirname(script))         # .../main/scenarios/x.py -> .../main
    from .mpc import install
    install(main_dir)
    os.chdir(os.path.dirname(script))                           # the scripts use sys.path.append('..')
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
