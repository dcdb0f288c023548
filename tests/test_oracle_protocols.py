"""Restates the reference's protocol tests against the CPU oracle:
PT/GSFSignatureTest.java, PT/PingPongTest.java (PT = protocols/src/test/java/.../protocols)."""
import numpy as np
import pytest

from tests.oracle_lib import OracleGSF, OraclePingPong

NL = "NetworkLatencyByDistanceWJitter"
NB = "RANDOM_SPEED=CONSTANT_TOR=0.00"


def popcount_rows(a):
    return np.array([sum(bin(int(w)).count("1") for w in row) for row in a])


def gsf32():
    p = OracleGSF(32, 1, 3, 20, 10, 10, 0, NB, NL)
    p.init()
    return p


def test_gsf_init_structure():  # GSFSignatureTest.testInit :22-43, testMaxSigInLevel :45-53
    p = gsf32()
    assert p.L == 6
    assert [len(p.peers(0, l)) for l in range(6)] == [0, 1, 2, 4, 8, 16]
    assert p.peers(0, 1)[0] == 1
    ls = p.level_scalars()
    assert ls["card"][0].tolist() == [1, 0, 0, 0, 0, 0]
    waited = p.level_rows(3)
    assert popcount_rows(waited).tolist() == [32] * 32  # levels' waited sets partition [0,32)
    for l in range(1, 6):
        pe = p.peers(5, l)
        lo = (5 >> l) << l
        own = set(range(lo, lo + (1 << l)))
        half = set(range(lo, lo + (1 << l))) - set(range((5 >> (l - 1)) << (l - 1), ((5 >> (l - 1)) << (l - 1)) + (1 << (l - 1))))
        assert set(pe.tolist()) == half and half <= own


def test_gsf_send_after_one_ms():  # testSend :55-60
    p = gsf32()
    p.run_ms(1)
    assert p.msgs_size() == 64
    assert p.msgs_live() == 64


def test_gsf_dead_nodes():  # testDeadNodes :72-80
    p = OracleGSF(32, int(0.8 * 32), 3, 20, 10, 10, int(0.1 * 32), NB, NL)
    p.init()
    assert int(p.attrs()["down"].sum()) == 3
    assert p.attrs()["down"][1] == 0


def test_gsf_last_finished_level():  # testGetLastFinishedLevel :82-93 -> 1,2,2,4
    p = gsf32()
    assert p.lib.wo_gsf_test_last_finished(p.h) == 1224


def test_gsf_simple_run():  # testSimpleRun :95-105
    p = gsf32()
    p.run_ms(10_000)
    assert (p.scalars()["card"] == 32).all()
    assert (popcount_rows(p.verified()) == 32).all()


def test_gsf_simple_threshold():  # testSimpleThreshold :107-124
    p = OracleGSF(64, int(.50 * 64), 3, 20, 10, 10, int(.2 * 64), NB, NL)
    p.init()
    p.run_ms(10_000)
    card = p.scalars()["card"]
    down = p.attrs()["down"]
    assert (card[down == 1] == 1).all()
    assert ((card[down == 0] >= 32) & (card[down == 0] <= 64)).all()


def test_gsf_copy_determinism():  # testCopy :126-147
    mk = lambda: OracleGSF(128, int(.75 * 128), 6, 10, 5, 10, int(.2 * 128), NB, NL)
    p1, p2 = mk(), mk()
    p1.init(); p2.init()
    while p1.time < 2000:
        p1.run_ms(1); p2.run_ms(1)
        assert p1.msgs_live() == p2.msgs_live()
        if p1.time % 50 == 0:
            assert p1.msgs_size() == p2.msgs_size() == p1.msgs_live()
            assert (p1.counters() == p2.counters()).all()
            assert (p1.verified() == p2.verified()).all()
            assert (p1.scalars()["to_verify"] == p2.scalars()["to_verify"]).all()


def test_gsf_runms_slicing_end_state():
    """Same seed, different runMs slicing: the end state agrees once the run is over (busy ticks),
    and verified sets only ever grow."""
    a = OracleGSF(64, 64, 3, 20, 10, 10, 0, NB, NL); a.init()
    b = OracleGSF(64, 64, 3, 20, 10, 10, 0, NB, NL); b.init()
    prev = popcount_rows(a.verified())
    for _ in range(300):
        a.run_ms(10)
        cur = popcount_rows(a.verified())
        assert (cur >= prev).all()
        prev = cur
    b.run_ms(3000)
    assert (popcount_rows(a.verified()) == 64).all() and (popcount_rows(b.verified()) == 64).all()


def test_gsf_params_validation():
    with pytest.raises(ValueError):
        OracleGSF(32, 33, 3, 20, 10, 10, 0, NB, NL)
    with pytest.raises(ValueError):
        OracleGSF(32, 30, 3, 20, 10, 10, 3, NB, NL)
    with pytest.raises(ValueError):
        OracleGSF(32, 1, 3, 20, 10, 10, 0, "NOPE_SPEED=CONSTANT_TOR=0.00", NL)


def test_gsf_aws_tor_builder():
    p = OracleGSF(256, 200, 4, 50, 20, 10, 25, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency")
    p.init()
    at = p.attrs()
    assert set(np.unique(at["extra"]).tolist()) <= {0, 500}
    assert 0.15 < (at["extra"] == 500).mean() < 0.5
    assert at["city"].min() >= 0 and at["city"].max() <= 10
    assert at["speed"].min() >= 0.33 and at["speed"].max() <= 2.99
    assert int(at["down"].sum()) == 25
    pairing = p.scalars()["pairing"]
    assert (pairing == np.maximum(1, (4 * at["speed"])).astype(np.int32)).all()
    p.run_ms(3000)
    card = p.scalars()["card"]
    assert (card[at["down"] == 0] >= 200).all()


def test_pingpong_simple():  # PingPongTest.testSimple :8-19
    p = OraclePingPong()
    p.init()
    p.run_ms(10_000)
    pongs = p.pongs()
    assert pongs[0] == 1000 and (pongs[1:] == 0).all()
    c = p.counters()
    assert c[1, 0] == 1001 and c[0, 0] == 1001  # node 0: 1000 pings (incl. to itself) + its own pong; got 1 ping + 1000 pongs
    assert (c[0, 1:] == 1).all() and (c[1, 1:] == 1).all()
    assert p.msgs_size() == 0


def test_pingpong_copy():  # PingPongTest.testCopy :21-36
    p1, p2 = OraclePingPong(), OraclePingPong()
    p1.init(); p1.run_ms(200)
    p2.init(); p2.run_ms(200)
    assert (p1.pongs() == p2.pongs()).all()


def test_pingpong_trajectory_is_monotone_and_soft_matches_readme():
    """README.md:123-134 prints 0,38,184,420,765,969,998,1000,1000,1000 for a latency class
    (NetworkLatencyByDistance) that no longer exists at this commit; with ByDistanceWJitter the curve is
    a *soft* vector only: monotone, complete by 1 s, same order of magnitude mid-run."""
    p = OraclePingPong()
    p.init()
    traj = []
    for _ in range(10):
        traj.append(int(p.pongs()[0]))
        p.run_ms(100)
    assert traj == sorted(traj) and traj[0] == 0
    assert int(p.pongs()[0]) == 1000
    # one-way latency <= (1144 pts * 10.87 mi * 0.022 + 4.862)/2 ~ 139 ms + jitter: ping+pong done by ~300 ms
    assert 0 < traj[1] < traj[2] < 1000 and traj[4] == 1000


def test_gsf_init_fast_equals_init():
    """initFast (threaded, used by the bench's CPU legs at large N) must reproduce init() exactly."""
    for n, dead in ((256, 25), (1024, 100), (4096, 409)):
        a = OracleGSF(n, int(.8 * n), 4, 50, 20, 10, dead, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"); a.init()
        b = OracleGSF(n, int(.8 * n), 4, 50, 20, 10, dead, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"); b.init_fast(4)
        assert a.rng_state() == b.rng_state()
        for node in (0, 1, n // 2 + 1, n - 1):
            for l in range(a.L):
                assert (a.peers(node, l) == b.peers(node, l)).all()
        for _ in range(20):
            a.run_ms(10); b.run_ms(10)
        assert (a.verified() == b.verified()).all() and (a.counters() == b.counters()).all() and a.rng_state() == b.rng_state()


def test_sanfermin_helper_kats():
    """PT/SanFerminTest.java: candidate sets of node 1 of 8 (:26-46) and pickNextNodes (:48-59)."""
    import ctypes as C
    from tests.oracle_lib import load

    lib = load()
    o = np.zeros(4, np.int32)
    sets = {}
    for lvl in (2, 1, 0):
        lib.wo_sf_helper_sets(1, 8, lvl, o.ctypes.data_as(C.POINTER(C.c_int32)))
        sets[lvl] = range(int(o[0]), int(o[1]))
    assert 0 in sets[2]
    assert 3 in sets[1] and 0 not in sets[1]
    assert 4 in sets[0] and 0 not in sets[0] and 3 not in sets[0]
    lib.wo_sf_helper_sets(4, 8, 0, o.ctypes.data_as(C.POINTER(C.c_int32)))
    assert 1 in range(int(o[0]), int(o[1]))  # helper4.isCandidate(n1, 0)
    out = np.zeros(16, np.int32)
    k = lib.wo_sf_helper_pick(1, 8, 2, 10, 1, out.ctypes.data_as(C.POINTER(C.c_int32)), 16)
    assert k == 1 and out[0] == 0
    assert lib.wo_sf_helper_pick(1, 8, 2, 10, 2, out.ctypes.data_as(C.POINTER(C.c_int32)), 16) == 0


def test_sanfermin_run_and_determinism():
    from tests.oracle_lib import OracleSanFermin

    a = OracleSanFermin(256, 256, 2, 48, 300, 1, None, None); a.init()
    b = OracleSanFermin(256, 256, 2, 48, 300, 1, None, None); b.init()
    for _ in range(400):
        a.run_ms(10); b.run_ms(10)
    sa, sb = a.scalars(), b.scalars()
    for k in sa:
        assert (sa[k] == sb[k]).all()
    assert (a.counters() == b.counters()).all()
    assert sa["done"].sum() > 200 and sa["agg"].max() == 256
    # a finished node aggregated the whole tree; doneAt = time + 2 * pairing (SanFerminSignature.java:396)
    done = sa["done"] == 1
    assert (sa["cpl"][done] == 0).all()
    assert (a.counters()[4][done] > 0).all()


def test_handel_liveness_and_determinism():
    """PT/HandelTest.java: testRun (:36-49, done within 20 s) and testCopy (:13-34, lock-step every ms)."""
    from tests.oracle_lib import OracleHandel

    mk = lambda: OracleHandel(64, 60, 6, 10, 5, 5, 10, 2, NB, NL, 100, False)
    p = mk(); p.init()
    while p.time < 20000 and (p.counters()[4][p.attrs()["down"] == 0] == 0).any():
        p.run_ms(1000)
    assert (p.counters()[4][p.attrs()["down"] == 0] > 0).all()
    a, b = mk(), mk()
    a.init(); b.init()
    while a.time < 2000:
        a.run_ms(1); b.run_ms(1)
        assert a.msgs_live() == b.msgs_live()
        if a.time % 100 == 0:
            assert (a.counters()[4] == b.counters()[4]).all()
            assert (a.scalars()["total_sig_size"] == b.scalars()["total_sig_size"]).all()


def test_handel_byzantine_suicide_blacklists():
    from tests.oracle_lib import OracleHandel

    q = OracleHandel(256, 180, 4, 50, 10, 20, 10, 64, "AWS_SPEED=GAUSSIAN_TOR=0.00", "AwsRegionNetworkLatency", 0, True)
    q.init()
    while q.continue_if() and q.time < 10000:
        q.run_ms(100)
    assert not q.continue_if()
    down = q.attrs()["down"] == 1
    bl = q.rows(5)
    ids = np.arange(256)
    # only byzantine (down) nodes are ever blacklisted, and some are
    listed = np.zeros(256, bool)
    for n in range(256):
        listed |= ((bl[n][ids // 64] >> (ids % 64).astype(np.uint64)) & np.uint64(1)).astype(bool)
    assert listed.any() and not (listed & ~down).any()
    with pytest.raises(ValueError):
        OracleHandel(100, 90, 4, 50, 10, 20, 10, 5, NB, NL)  # power of two only (Handel.java:118-120)


def test_cappos_oracle_liveness_and_threshold():
    """SanFerminCappos.sigsPerTime parameters scaled down (SanFerminCappos.java:465-471): everybody finishes, the threshold is
    reached before the end, and a finished node holds at least the threshold."""
    from tests.oracle_lib import OracleCappos

    o = OracleCappos(1024, 512, 2, 48, 150, 50, "RANDOM_SPEED=CONSTANT_TOR=0.00", "NetworkLatencyByDistanceWJitter")
    o.init()
    for _ in range(600):
        o.run_ms(10)
    s = o.scalars()
    assert s["done"].sum() >= 1000
    done = s["done"] == 1
    assert (s["sigs"][done] >= 512).all()
    assert (s["threshold_at"][done] > 0).all() and (s["threshold_at"][done] <= o.counters()[4][done]).all()
