"""Wall-clock context manager feeding the ``*_epoch_time`` / ``*_total_time``
attributes (same role as the reference's ``numbskull/timer.py``)."""
import time


class Timer(object):
    def __enter__(self):
        self.start = time.perf_counter()
        return self

    def __exit__(self, *exc):
        self.end = time.perf_counter()
        self.interval = self.end - self.start
        return False
