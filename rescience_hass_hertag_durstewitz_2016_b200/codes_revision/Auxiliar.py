"""`set_simulation_dir` and the `time_report` decorator of codes_revision/Auxiliar.py (used by Tests.py and by the setup modules)."""
import functools
import os
import time


def set_simulation_dir(name=None):
    """Create and return a fresh run directory: `name`, or the first free `Simulation_<k>` (codes_revision/Auxiliar.py:4-17)."""
    if name is None:
        k = 0
        while os.path.isdir("Simulation_%d" % k):
            k += 1
        name = "Simulation_%d" % k
    elif os.path.exists(name):
        raise NameError("Path already exists.")
    os.mkdir(name)
    print("REPORT: simulation_dir created:", name, end="\n\n")
    return name


def _span(seconds):
    h, rest = divmod(seconds, 3600)
    m, s = divmod(rest, 60)
    parts = (["%d h" % h] if h else []) + (["%d min" % m] if h or m else []) + ["%.2f s" % s]
    return ", ".join(parts[:-1]) + (" and " if len(parts) > 1 else "") + parts[-1]


def time_report(text=None):
    """Decorator: announce the call and report its wall time (codes_revision/Auxiliar.py:19-46)."""
    def wrap(f):
        @functools.wraps(f)
        def timed(*args, **kwargs):
            what = text if text is not None else f.__name__
            print("|   TIME REPORT:   | Starting %s ..." % what, end="\n\n")
            t0 = time.time()
            out = f(*args, **kwargs)
            print("||   TIME REPORT: %s completed in %s.   ||" % (what, _span(time.time() - t0)), end="\n\n")
            return out
        return timed
    return wrap
