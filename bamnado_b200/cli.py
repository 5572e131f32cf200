"""Console entry point forwarding to the native CLI (mirrors bamnado-python/python/bamnado/cli.py)."""
import contextlib
import signal
import sys
import threading

from . import _native


@contextlib.contextmanager
def _default_sigint():
    """Restore OS-default SIGINT while the native run holds the thread (cli.py:11-22 of the reference)."""
    if threading.current_thread() is not threading.main_thread():
        yield
        return
    previous = signal.getsignal(signal.SIGINT)
    signal.signal(signal.SIGINT, signal.SIG_DFL)
    try:
        yield
    finally:
        signal.signal(signal.SIGINT, previous)


def main(argv: list[str] | None = None) -> int:
    """Run the CLI and return its exit code."""
    args = sys.argv[1:] if argv is None else list(argv)
    with _default_sigint():
        return _native.lib().cli_main(["bamnado", *args])


if __name__ == "__main__":
    sys.exit(main())
