"""src/helpers.py: the terminal progress bar barrier_method prints (src/barrier_method.py:178-180).
Kept call-compatible; silent unless WHITOMO_PROGRESS=1 (a per-iteration print would serialise the GPU)."""
import os
import sys


def printProgressBar(iteration, total, prefix='', suffix='', decimals=1, length=100, fill='#'):
    if os.environ.get("WHITOMO_PROGRESS") != "1":
        return
    frac = iteration / float(total) if total else 1.0
    filled = int(length * frac)
    sys.stdout.write('\r%s |%s| %.*f%% %s' % (prefix, fill * filled + '-' * (length - filled), decimals, 100 * frac, suffix))
    if iteration == total:
        sys.stdout.write('\n')
    sys.stdout.flush()
