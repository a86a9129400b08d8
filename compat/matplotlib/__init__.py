"""Stand-in for `matplotlib` (train.py:8,21-97; render_rollout.py:6,122-143): the plotting calls of the reference's
scripts become no-ops and `savefig` leaves a 1x1 PNG where the plot would be, so the scripts run unchanged on a
headless box without the package.  A real matplotlib found elsewhere on sys.path takes this module's place."""
import os
import sys

from _shim import prefer_real  # noqa: E402

if not prefer_real(__name__):
    __version__ = "0.0+cgnn-shim"

    def use(*_, **__):
        pass
