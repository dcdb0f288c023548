"""StatsHelper mirror (wittgenstein_b200/run_multiple.py) against the reference's own test, CT/StatsTest.java:10-22, plus the
integer semantics of getStatsOn (C/utils/StatsHelper.java:122-134)."""
import numpy as np
import pytest

from wittgenstein_b200.run_multiple import SimpleStats, avg, get_stats_on


def test_avg_like_the_reference():
    a = avg([SimpleStats(10, 20, 30), SimpleStats(16, 26, 36)])
    assert isinstance(a, SimpleStats)
    assert (a.min, a.max, a.avg) == (13, 23, 33)


def test_avg_of_one_is_identity_and_empty_throws():
    s = SimpleStats(1, 2, 3)
    assert avg([s]) is s  # StatsHelper.java:36-38
    with pytest.raises(ValueError):
        avg([])  # IllegalStateException :33-35


def test_get_stats_on_truncates_like_java_longs():
    s = get_stats_on(np.array([1, 2, 4], np.int64))
    assert (s.min, s.max, s.avg) == (1, 4, 2)  # 7 / 3 == 2
    s = get_stats_on(np.array([-1, -2, -4], np.int64))
    assert s.avg == -2  # Java: -7 / 3 == -2 (towards zero), not floor
