"""Switch for the side-stream overlap inside a step.  bench.py turns it off for its roofline pass so that a
bracketed launch's duration is the kernel's own, not the kernel sharing its SMs with a concurrent glue kernel.
``P2I_OVERLAP_MASK`` (bit 0: discriminator 3-D branch, 1: bias column sums, 2: InputBlock, 3: DO-Conv composition backward, 4: generator
weight gradients on the side stream, 5: discriminator weight gradients on the aux stream, 6: the fake / real discriminator calls
of the D update on two lanes, 7: the stem's weight gradient on the side stream, 8: level-3 weight gradients on a high-priority stream)
selects individual overlaps for A/B measurements; default all on."""
import os

ENABLED = True
MASK = int(os.environ.get("P2I_OVERLAP_MASK", "511"))
D_BRANCH, D_COLSUM, G_INPUT, G_DOCONV, G_WGRAD, D_WGRAD, D_PAIR, G_STEMDW, G_WGRAD3 = 1, 2, 4, 8, 16, 32, 64, 128, 256


def set_stream_overlap(flag: bool) -> None:
    global ENABLED
    ENABLED = bool(flag)


def pick(side, main, which: int = 511):
    """The stream a forked branch should run on."""
    return side if (ENABLED and (MASK & which)) else main
