"""Wall-clock progress line shared by the example cases."""
import time


class Progress:
    def __init__(self):
        self.start = self.last = time.time()

    @staticmethod
    def _hms(seconds):
        m, s = divmod(int(seconds), 60)
        h, m = divmod(m, 60)
        return "%dh %dm %ds" % (h, m, s)

    def report(self, step, **quantities):
        now = time.time()
        since_last, elapsed = now - self.last, now - self.start
        self.last = now
        tail = ", ".join("%s = %g" % kv for kv in quantities.items())
        print("step %d: %s since the last report, %s elapsed%s" % (step, self._hms(since_last), self._hms(elapsed),
                                                                   "; " + tail if tail else ""))
