"""`python -m sim_juncs_b200` -- the reference's command line (src/main.cpp:13-67) on the B200 engine,
all in Python: argparse.h-compatible flags and params.conf (settings.py), the own CGS reader (cgs.py),
BoundGeom on the C ABI, monitor series written by output.py.  Same phase timers on stdout.
Engine-only extras: --fp32, --real-fields, --gpus N (z-slabs over N GPUs of this box, one process), --current-sources
(src_time::is_integrated = false: the drive enters as a current on D; see scripts/compare_with_meep.py)."""
import sys
import time

from .bound_geom import BoundGeom
from .cgs import CgsError
from .settings import ParseSettings


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    precision = "f32" if "--fp32" in argv else "f64"
    n_sets = 1 if "--real-fields" in argv else 2
    integrated = "--current-sources" not in argv
    argv = [a for a in argv if a not in ("--fp32", "--real-fields", "--current-sources")]
    gpus = None
    if "--gpus" in argv:
        i = argv.index("--gpus")
        gpus = int(argv[i + 1])
        del argv[i:i + 2]
    args = ParseSettings()
    ret = args.parse_args(argv)
    if ret:
        return ret
    ret = args.parse_conf_file(args.conf_fname or "params.conf")
    if ret:
        return ret
    args.correct_defaults()
    start = time.monotonic()
    try:
        geom = BoundGeom(args, None, precision=precision, n_sets=n_sets, gpus=gpus, integrated=integrated)
    except CgsError as e:
        print(e)
        return e.code
    end_init = time.monotonic()
    print("initialization completed in %d ms" % int(1e3 * (end_init - start)))
    geom.run(args.out_dir)
    end_run = time.monotonic()
    print("simulation completed in %g s (%d steps)" % (end_run - end_init, geom.n_t_pts))
    geom.save_field_times(args.out_dir)
    end_write = time.monotonic()
    print("saving timeseries completed in %d ms" % int(1e3 * (end_write - end_run)))
    print("total time: %g s" % (end_write - start))
    return 0


if __name__ == "__main__":
    sys.exit(main())
