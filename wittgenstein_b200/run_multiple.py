"""RunMultipleTimes (core/RunMultipleTimes.java) and the StatsHelper pieces it uses (core/utils/StatsHelper.java).

The reference runs its `runCount` seeds one after the other (:44-48).  Seeded runs are independent, so here up to
`concurrency` of them are in flight at once, each on its own engine instance / CUDA stream (one host thread per run; the
C ABI releases the GIL while the device works).  For small networks, where a single run is bound by kernel-launch
latency, this is what fills the GPU.  Results are identical to running the seeds sequentially: every run owns its state.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .network import set_thread_device


class SimpleStats:
    """StatsHelper.SimpleStats (:83-120): min / max / avg as Java longs (avg = total / count, truncating)."""

    fields = ("min", "max", "avg")

    def __init__(self, mn, mx, avg):
        self.min, self.max, self.avg = int(mn), int(mx), int(avg)

    def get(self, f):
        return getattr(self, f)

    def __eq__(self, o):
        return (self.min, self.max, self.avg) == (o.min, o.max, o.avg)

    def __repr__(self):
        return f"min: {self.min}, max:{self.max}, avg:{self.avg}"


def get_stats_on(values):
    """StatsHelper.getStatsOn (:122-134) over the values of the live nodes."""
    v = np.asarray(values, np.int64)
    if v.size == 0:
        raise ZeroDivisionError("no live node")
    tot = int(v.sum())
    q = abs(tot) // v.size
    return SimpleStats(v.min(), v.max(), q if tot >= 0 else -q)


class DoneAtStatGetter:
    """StatsHelper.DoneAtStatGetter (:158-163)"""

    def get(self, protocol):
        net = protocol.network()
        live = net.attrs()["down"] == 0
        return get_stats_on(net.counters()[4][live])


class MsgReceivedStatGetter:
    """StatsHelper.MsgReceivedStatGetter (:165-170)"""

    def get(self, protocol):
        net = protocol.network()
        live = net.attrs()["down"] == 0
        return get_stats_on(net.counters()[0][live])


def avg(stats):
    """StatsHelper.avg (:32-52): field-wise sum / count (truncating long division)."""
    if not stats:
        raise ValueError("no stats")
    if len(stats) == 1:
        return stats[0]
    out = []
    for f in SimpleStats.fields:
        tot = sum(s.get(f) for s in stats)
        q = abs(tot) // len(stats)
        out.append(q if tot >= 0 else -q)
    return SimpleStats(*out)


def cont_until_done(protocol):
    """RunMultipleTimes.contUntilDone (:87-97): some live node has doneAt == 0."""
    net = protocol.network()
    live = net.attrs()["down"] == 0
    return bool((net.counters()[4][live] == 0).any())


class RunMultipleTimes:
    """RunMultipleTimes(p, runCount, maxTime, statsGetters, finalCheck).run(contIf) -> one averaged Stat per getter."""

    def __init__(self, p, run_count, max_time, stats_getters, final_check=None):
        self.p = p
        self.run_count = run_count
        self.max_time = max_time
        self.stats_getters = list(stats_getters)
        self.final_check = final_check

    def _one(self, i, cont_if):
        if self.devices:
            set_thread_device(self.devices[i % len(self.devices)])
        c = self.p.copy()
        c.network().set_seed(i)  # :47
        c.init()
        net = c.network()
        try:
            while True:
                did = net.run_ms(10)
                if not ((self.max_time == 0 or net.time < self.max_time) and ((not did) or (cont_if is not None and cont_if(c)))):
                    break
        except Exception as e:  # noqa: BLE001  (:52-60)
            raise RuntimeError(f"Failed execution for random seed of {i}, time={net.time}") from e
        if self.final_check is not None and not self.final_check(c):
            raise RuntimeError(f"Failed execution for random seed of {i}")
        res = [sg.get(c) for sg in self.stats_getters]
        t = net.time
        net.close()
        return res, t

    def run(self, cont_if, concurrency=8, devices=None):
        """`devices`: CUDA device ids to spread the runs over (run i goes to devices[i % len]); None = the default device."""
        self.devices = list(devices) if devices else None
        if concurrency <= 1:
            per = [self._one(i, cont_if) for i in range(self.run_count)]
        else:
            with ThreadPoolExecutor(max_workers=concurrency) as ex:
                per = list(ex.map(lambda i: self._one(i, cont_if), range(self.run_count)))
        self.end_times = [t for _, t in per]
        return [avg([r[k] for r, _ in per]) for k in range(len(self.stats_getters))]


class ProgressPerTime:
    """ProgressPerTime (core/ProgressPerTime.java:16-140) without the graph output (tools/Graph is out of scope): `roundCount`
    seeded rounds, a stat sampled every `statEachXms` ms of each round, the per-round and averaged node counters.  Rounds are
    independent, so up to `concurrency` of them run at once, each on its own engine instance (see RunMultipleTimes above)."""

    def __init__(self, template, stats_getter, round_count, end_callback=None, stat_each_x_ms=10):
        if round_count <= 0:
            raise ValueError(f"roundCount must be greater than 0. roundCount={round_count}")  # :36-39
        self.protocol = template.copy()
        self.stats_getter = stats_getter
        self.round_count = round_count
        self.end_callback = end_callback
        self.stat_each_x_ms = stat_each_x_ms

    def _round(self, r, cont_if):
        p = self.protocol.copy()
        p.network().set_seed(r)  # :71
        p.init()
        net = p.network()
        series = []  # (time, Stat) lines of Graph.ReportLine
        while True:
            net.run_ms(self.stat_each_x_ms)
            s = self.stats_getter.get(p)
            series.append((net.time, s))
            if not cont_if(p):
                break
        if self.end_callback is not None:
            self.end_callback(p)
        live = net.attrs()["down"] == 0
        c = net.counters()
        summary = {"bytes_sent": get_stats_on(c[2][live]), "bytes_rcvd": get_stats_on(c[3][live]), "msg_sent": get_stats_on(c[1][live]),
                   "msg_rcvd": get_stats_on(c[0][live]), "done_at": get_stats_on(c[4][live])}
        net.close()
        return series, summary

    def run(self, cont_if, concurrency=8):
        if concurrency <= 1:
            rounds = [self._round(r, cont_if) for r in range(self.round_count)]
        else:
            with ThreadPoolExecutor(max_workers=concurrency) as ex:
                rounds = list(ex.map(lambda r: self._round(r, cont_if), range(self.round_count)))
        self.series = [s for s, _ in rounds]
        self.summaries = [m for _, m in rounds]
        # "Average on the N rounds" (:122-129): truncating long division of the sums of the per-round averages
        self.average = {k: sum(m[k].avg for m in self.summaries) // self.round_count for k in self.summaries[0]}
        return self.series
