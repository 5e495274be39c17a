"""Console progress bar used by every ``fit`` when ``verbose`` is true.

Behavioural twin of the reference's SimRank/Helper.py:3-19 (same module-level names, same
exceptions, byte-identical text on stdout); the implementation is independent.
"""
import sys

BAR_LENGTH = 30

_PREFIX = "\rPercent: "


def _render(fraction, suffix):
    filled = int(round(BAR_LENGTH * fraction))
    bar = "#" * filled + "-" * (BAR_LENGTH - filled)
    return f"{_PREFIX}[{bar}] {round(fraction * 100, 1)}% {suffix}"


def update_progress(progress):
    """Draw the bar for ``progress`` in [0, 1).  Integers are accepted as floats, anything else
    that is not a float raises ``ValueError('Progress must be float')``, negatives raise
    ``ValueError('Progress below 0')``, and values >= 1 clamp to a full bar followed by
    ``'Done...\\r\\n'`` (Helper.py:7-15)."""
    if isinstance(progress, int):
        progress = float(progress)
    if not isinstance(progress, float):
        raise ValueError("Progress must be float")
    if progress < 0:
        raise ValueError("Progress below 0")
    done = progress >= 1
    sys.stdout.write(_render(1 if done else progress, "Done...\r\n" if done else ""))
    sys.stdout.flush()
