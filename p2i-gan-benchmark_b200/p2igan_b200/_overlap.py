"""Switch for the side-stream overlap inside a step.  bench.py turns it off for its roofline pass so that a
bracketed launch's duration is the kernel's own, not the kernel sharing its SMs with a concurrent glue kernel."""
ENABLED = True


def set_stream_overlap(flag: bool) -> None:
    global ENABLED
    ENABLED = bool(flag)


def pick(side, main):
    """The stream a forked branch should run on."""
    return side if ENABLED else main
