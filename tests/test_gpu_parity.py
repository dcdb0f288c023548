"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle on the same
seeded inputs — bit-exact (integer / bitmap work).  Sizes are chosen so the oracle finishes in seconds;
full-size properties live in test_gpu_properties.py."""
import numpy as np
import pytest

from tests.oracle_lib import OracleGSF, OraclePingPong
from tests.parity import compare_gsf, compare_init, run_lockstep
from wittgenstein_b200 import GSFSignature, GSFSignatureParameters, PingPong, PingPongParameters  # noqa: E402

pytestmark = pytest.mark.gpu

NL = "NetworkLatencyByDistanceWJitter"
NB = "RANDOM_SPEED=CONSTANT_TOR=0.00"
AWS_NB, AWS_NL = "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"


def mk(n, thr, pairing, timeout, period, acc, dead, nb, nl, seed=None):
    from wittgenstein_b200 import GSFSignature, GSFSignatureParameters

    prm = GSFSignatureParameters(n, thr, pairing, timeout, period, acc, dead, nb, nl)
    p = GSFSignature(prm)
    o = OracleGSF(n, prm.threshold, pairing, timeout, period, acc, prm.nodes_down, nb, nl)
    if seed is not None:
        p.network().set_seed(seed)
        o.lib.wo_gsf_set_seed(o.h, seed)
    p.init()
    o.init()
    bad = compare_init(p, o) + compare_gsf(p, o, "init")
    assert not bad, bad
    return p, o


@pytest.mark.parametrize("latency", [None, "NetworkNoLatency", "IC3NetworkLatency", "NetworkFixedLatency(100)",
                                     "NetworkUniformLatency(200)"])
def test_pingpong_config1(latency):
    """BASELINE config #1: PingPong 1000 nodes, 1000 ms, 10 x runMs(100)."""
    from wittgenstein_b200 import PingPong, PingPongParameters

    p = PingPong(PingPongParameters(1000, None, latency))
    o = OraclePingPong(1000, None, latency)
    p.init(); o.init()
    a, b = p.network().attrs(), o.attrs()
    for k in ("x", "y", "extra", "down"):
        assert (a[k] == b[k]).all()
    for _ in range(10):
        assert p.network().run_ms(100) == o.run_ms(100)
        assert (p.pongs() == o.pongs()).all()
        assert (p.network().counters() == o.counters()).all()
        assert p.network().msgs_size() == o.msgs_size()
    assert p.pongs()[0] == 1000 and p.network().msgs_size() == 0
    assert p.network().stats()["kernel_launches"] > 0


def test_pingpong_aws_tor_and_seeds():
    from wittgenstein_b200 import PingPong, PingPongParameters

    for seed in (0, 1, 2, -7):
        p = PingPong(PingPongParameters(777, AWS_NB, AWS_NL))
        o = OraclePingPong(777, AWS_NB, AWS_NL, seed=seed)
        p.network().set_seed(seed)
        p.init(); o.init()
        for ms in (1, 2, 7, 90, 400, 1500, 3000):
            assert p.network().run_ms(ms) == o.run_ms(ms)
            assert (p.pongs() == o.pongs()).all() and (p.network().counters() == o.counters()).all()
            assert p.network().msgs_size() == o.msgs_size()
        assert p.pongs()[0] == 777


def test_gsf_32_every_ms():
    """Reference test size (PT/GSFSignatureTest): 32 nodes, compared after every single ms."""
    p, o = mk(32, 1.0, 3, 20, 10, 10, 0, NB, NL)
    assert p.network().run_ms(1) == o.run_ms(1)
    assert p.network().msgs_size() == 64  # GSFSignatureTest.testSend
    run_lockstep(p, o, 1, 500)
    assert (p.scalars()["card"] == 32).all()


def test_gsf_128_copy_params_dead_nodes():
    """GSFSignatureTest.testCopy parameters: 128 nodes, 20 % dead, pairing 6, timeout 10, period 5."""
    p, o = mk(128, .75, 6, 10, 5, 10, .2, NB, NL)
    run_lockstep(p, o, 1, 300)
    run_lockstep(p, o, 13, 2000, full_every=4)


@pytest.mark.parametrize("seed", [0, 3])
def test_gsf_256_aws_tor_slicing(seed):
    """AWS regions + Tor + uniform speed; odd runMs slicing exercises the end-of-window conditional pass."""
    p, o = mk(256, 0.8, 4, 50, 20, 10, 0.1, AWS_NB, AWS_NL, seed=seed)
    for step in (1, 2, 3, 5, 7, 11, 1, 1, 29):
        run_lockstep(p, o, step, p.network().time + step)
    run_lockstep(p, o, 7, 2100, full_every=8)
    assert not p.continue_if() and not o.continue_if()


def test_gsf_accel_off_and_small_accel():
    for acc in (0, 1, 3):
        p, o = mk(64, 1.0, 2, 30, 10, acc, 0, NB, "NetworkFixedLatency(100)")
        run_lockstep(p, o, 10, 1500, full_every=5)


def test_gsf_4096_config2():
    """BASELINE config #2: GSFSignature 4096 nodes = GSFSignature.newProtocol() (GSFSignature.java:684-697),
    runMs(10) until every live node reached the threshold; full state compared every 100 ms."""
    p, o = mk(4096, 0.85, 4, 50, 20, 10, 0.10, AWS_NB, AWS_NL)
    steps = 0
    while o.continue_if() and steps < 400:
        assert p.network().run_ms(10) == o.run_ms(10)
        steps += 1
        bad = compare_gsf(p, o, f"t={o.time}", full=(steps % 10 == 0))
        assert not bad, bad
    assert not p.continue_if()
    bad = compare_gsf(p, o, "end", full=True)
    assert not bad, bad
    st = p.network().stats()
    assert st["kernel_launches"] > 1000 and st["min_pool_free"] > 0


def test_gsf_4096_random_positions_no_tor():
    p, o = mk(4096, 0.99, 3, 50, 10, 10, 0, NB, NL)
    for _ in range(12):
        assert p.network().run_ms(50) == o.run_ms(50)
    bad = compare_gsf(p, o, "t=600", full=True)
    assert not bad, bad


def _sf_compare(p, o, tag):
    bad = []
    if p.network().rng_state() != o.rng_state():
        bad.append(f"{tag}: rd state")
    if p.network().msgs_size() != o.msgs_live():
        bad.append(f"{tag}: msgs.size()")
    if not (p.network().counters() == o.counters()).all():
        bad.append(f"{tag}: counters")
    a, b = p.scalars(), o.scalars()
    for k in a:
        if not (a[k] == b[k]).all():
            bad.append(f"{tag}: {k}")
    return bad


@pytest.mark.parametrize("n,nb,nl,seed,step", [(64, None, None, None, 1), (1024, None, None, None, 7),
                                                 (1024, AWS_NB, AWS_NL, 5, 10), (4096, None, "IC3NetworkLatency", 2, 25)])
def test_sanfermin_parity(n, nb, nl, seed, step):
    """SanFerminSignature (swap requests / replies, timeouts, 2-candidate shuffled sends) vs the oracle."""
    from tests.oracle_lib import OracleSanFermin
    from wittgenstein_b200 import SanFerminSignature, SanFerminSignatureParameters

    p = SanFerminSignature(SanFerminSignatureParameters(n, n, 2, 48, 300, 1, False, nb, nl))
    o = OracleSanFermin(n, n, 2, 48, 300, 1, nb, nl)
    if seed is not None:  # after construction: node positions were already drawn (SanFerminSignature.java:112-129)
        p.network().set_seed(seed)
        o.set_seed(seed)
    p.init(); o.init()
    assert not _sf_compare(p, o, "init")
    while o.time < 3500:
        assert p.network().run_ms(step) == o.run_ms(step)
        bad = _sf_compare(p, o, f"t={o.time}")
        assert not bad, bad
    assert p.scalars()["done"].sum() > n // 2


@pytest.mark.parametrize("n,k,seed,force", [(256, 3, None, False), (1024, 10, 7, False), (512, 50, 2, True)])
def test_sanfermin_several_candidates(n, k, seed, force):
    """candidateCount > 1: pickNextNodes hands out k + 1 nodes, shuffled on the network RNG before the request is sent
    (SanFerminHelper.java:123-157); `force` re-derives the draw indices serially on every tick."""
    from tests.oracle_lib import OracleSanFermin
    from wittgenstein_b200 import SanFerminSignature, SanFerminSignatureParameters

    p = SanFerminSignature(SanFerminSignatureParameters(n, n, 2, 48, 300, k, False, None, None))
    o = OracleSanFermin(n, n, 2, 48, 300, k, None, None)
    if force:
        p.network().set_tunable("force_shuffle_serial", 1)
    if seed is not None:
        p.network().set_seed(seed)
        o.set_seed(seed)
    p.init(); o.init()
    while o.time < 3500:
        assert p.network().run_ms(10) == o.run_ms(10)
        bad = _sf_compare(p, o, f"t={o.time}")
        assert not bad, bad
    assert p.scalars()["done"].sum() > n // 2


def test_sanfermin_16384_shipped_scenario():
    """SanFerminSignature.sigsPerTime() parameters (SanFerminSignature.java:566-571): 16 384 nodes, runMs(10) to 6 s."""
    from tests.oracle_lib import OracleSanFermin
    from wittgenstein_b200 import SanFerminSignature, SanFerminSignatureParameters

    n = 16384
    p = SanFerminSignature(SanFerminSignatureParameters(n, n, 2, 48, 300, 1, False, None, None))
    o = OracleSanFermin(n, n, 2, 48, 300, 1, None, None)
    p.init(); o.init()
    for i in range(600):
        assert p.network().run_ms(10) == o.run_ms(10)
        if i % 50 == 49:
            bad = _sf_compare(p, o, f"t={o.time}")
            assert not bad, bad
    assert not _sf_compare(p, o, "end")


def _casper_pair(cyc, bpc, apr, nb, nl, seed, delay, until, kind="WF"):
    from tests.oracle_lib import OracleCasper
    from wittgenstein_b200 import CasperIMD, CasperParemeters

    p = CasperIMD(CasperParemeters(cyc, False, bpc, apr, 1000, 1, nb, nl))
    o = OracleCasper(cyc, False, bpc, apr, 1000, 1, nb, nl)
    if seed is not None:  # after construction: the observer was already drawn (CasperIMD.java:87)
        p.network().set_seed(seed)
        o.set_seed(seed)
    p.network().set_tunable("casper_votes", until // (8000 * cyc) + 3)
    p.init(delay, kind); o.init(delay, kind)
    return p, o


@pytest.mark.gpu
@pytest.mark.parametrize("cyc,bpc,apr,nb,nl,seed,delay,step,until", [
    (5, 5, 80, None, None, None, 0, 1000, 60000),            # the fixture of PT/CasperIMDTest.java:11-13 through init()
    (2, 2, 6, None, None, None, 0, 1, 20000),                # runMs(1) slicing
    (3, 3, 20, None, None, None, 9000, 500, 200000),         # the Byzantine block arrives after its successor: forks + vote counting
    (3, 3, 20, None, None, None, -7500, 500, 120000),        # ByzBlockProducerWF "late" path
    (4, 2, 16, AWS_NB, AWS_NL, 7, 3000, 1000, 300000),       # AWS regions, Tor, re-seeded, several cycles
    (1, 2, 2, None, "NetworkNoLatency", None, 0, 1000, 40000),  # PT/CasperByzantineTest.java:8-15
])
def test_casper_parity(cyc, bpc, apr, nb, nl, seed, delay, step, until):
    """CasperIMD (blocks, attestations, fork choice, ByzBlockProducerWF) vs the oracle, every step."""
    from tests.parity import compare_casper

    p, o = _casper_pair(cyc, bpc, apr, nb, nl, seed, delay, until)
    assert not compare_casper(p, o, "init")
    for t in (8000 + delay, 12000, 20000):
        assert p.network().msgs_size_at(t) == o.msgs_size_at(t)
    while o.time < until:
        assert p.network().run_ms(step) == o.run_ms(step)
        bad = compare_casper(p, o, f"t={o.time}")
        assert not bad, bad
    assert not compare_casper(p, o, "end", atts=True)
    assert len(p.blocks()["height"]) > 2


@pytest.mark.gpu
@pytest.mark.parametrize("kind,cyc,bpc,apr,delay,until", [("plain", 3, 3, 20, 9000, 150000), ("SF", 4, 3, 12, 3000, 250000),
                                                         ("NS", 2, 4, 10, -3000, 250000)])
def test_casper_other_byzantine_producers(kind, cyc, bpc, apr, delay, until):
    """ByzBlockProducer / ByzBlockProducerSF / ByzBlockProducerNS (P/CasperIMD.java:511-640) vs the oracle."""
    from tests.parity import compare_casper

    p, o = _casper_pair(cyc, bpc, apr, None, None, None, delay, until, kind)
    while o.time < until:
        assert p.network().run_ms(1000) == o.run_ms(1000)
        bad = compare_casper(p, o, f"t={o.time}")
        assert not bad, bad
    assert not compare_casper(p, o, "end", atts=True)
    assert p.byz()["on_direct_father"] + p.byz()["on_older_ancestor"] + p.byz()["skipped"] > 0 or kind == "NS"


@pytest.mark.gpu
def test_casper_byzantine_wf_schedule():
    """PT/CasperByzantineTest.java:12-36 through the engine: observer head heights at 9 / 10 / 18 / 26 s."""
    from wittgenstein_b200 import CasperIMD, CasperParemeters

    p = CasperIMD(CasperParemeters(1, False, 2, 2, 1000, 1, None, "NetworkNoLatency"))
    p.init(0)
    net = p.network()
    net.run_ms(9000)
    assert p.node_state()["head"][0] == 0
    net.run_ms(1000)
    b = p.blocks()
    assert b["height"][p.node_state()["head"][0]] == 1 and b["producer"][p.node_state()["head"][0]] == 1
    net.run_ms(8000)
    b = p.blocks()
    assert b["height"][p.node_state()["head"][0]] == 2 and b["producer"][p.node_state()["head"][0]] != 1
    net.run_ms(8000)
    b = p.blocks()
    assert b["height"][p.node_state()["head"][0]] == 3 and b["producer"][p.node_state()["head"][0]] == 1


@pytest.mark.gpu
def test_casper_stopped_nodes_and_partition():
    """Stopped attesters neither vote nor receive; a partition drops cross-partition deliveries (Network.java:478-486, 606)."""
    from tests.parity import compare_casper

    p, o = _casper_pair(3, 3, 24, None, None, 4, 0, 120000)
    for i in (10, 11, 40):
        p.network().stop_node(i); o.stop_node(i)
    for k in range(60):
        assert p.network().run_ms(1000) == o.run_ms(1000)
        if k == 20:
            p.network().partition(0.4); o.partition(0.4)
        if k == 30:
            p.network().start_node(11); o.start_node(11)
        bad = compare_casper(p, o, f"t={o.time}")
        assert not bad, bad


@pytest.mark.gpu
def test_casper_errors():
    from wittgenstein_b200 import CasperIMD, CasperParemeters, WtgError

    with pytest.raises(WtgError):
        CasperIMD(CasperParemeters(0, False, 2, 2, 1000, 1))
    p = CasperIMD(CasperParemeters(4, False, 3, 200, 1000, 1))
    p.network().set_tunable("rec_cap", 100)
    p.init(0)
    p.network().partition(0.5)
    with pytest.raises(WtgError):
        p.network().end_partition()  # BlockChainNetwork.endPartition re-sends every head: needs one record slot per node
    with pytest.raises(WtgError):
        p.network().run_ms(0)  # Network.java:319-321


def _handel_compare(p, o, tag, full=True):
    bad = []
    if p.network().rng_state() != o.rng_state():
        bad.append(f"{tag}: rd state")
    if p.network().msgs_size() != o.msgs_live():
        bad.append(f"{tag}: msgs.size()")
    if not (p.network().counters() == o.counters()).all():
        bad.append(f"{tag}: counters")
    a, b = p.scalars(), o.scalars()
    for k in a:
        if not (a[k] == b[k]).all():
            bad.append(f"{tag}: {k}")
    if full:
        for w in range(6):
            if not (p.rows(w) == o.rows(w)).all():
                bad.append(f"{tag}: rows {w}")
        l1, l2 = p.level_scalars(), o.level_scalars()
        for k in l1:
            if not (l1[k] == l2[k]).all():
                bad.append(f"{tag}: level {k}")
    return bad


def _handel_pair(n, thr, pairing, lw, extra, period, fp, down, nb, nl, desync, byz, seed=None, hidden=False):
    from tests.oracle_lib import OracleHandel
    from wittgenstein_b200 import Handel, HandelParameters

    p = Handel(HandelParameters(n, thr, pairing, lw, extra, period, fp, down, nb, nl, desync, byz, hidden))
    o = OracleHandel(n, thr, pairing, lw, extra, period, fp, down, nb, nl, desync, byz, seed=seed, hidden_byzantine=hidden)
    if seed is not None:
        p.network().set_seed(seed)
    p.init(); o.init()
    for node in (0, 1, n // 2, n - 1):
        assert (p.ranks(node) == o.ranks(node)).all()
        for l in range(p.levels):
            assert (p.peers(node, l) == o.peers(node, l)).all()
    a, b = p.network().attrs(), o.attrs()
    for k in a:
        assert (a[k] == b[k]).all()
    assert not _handel_compare(p, o, "init")
    return p, o


def test_handel_64_test_copy_params_every_ms():
    """PT/HandelTest parameters (64 nodes, 2 dead, desynchronised start 100 ms), compared after every ms."""
    p, o = _handel_pair(64, 60, 6, 10, 5, 5, 10, 2, NB, NL, 100, False)
    while o.time < 1200:
        assert p.network().run_ms(1) == o.run_ms(1)
        bad = _handel_compare(p, o, f"t={o.time}")
        assert not bad, bad
    assert not p.continue_if() and not o.continue_if()


@pytest.mark.parametrize("seed", [None, 3])
def test_handel_1024_byzantine_suicide(seed):
    """BASELINE config #3 scaled down: 25 % Byzantine (suicide attack), AWS regions, fast path 10."""
    p, o = _handel_pair(1024, 760, 4, 50, 10, 20, 10, 256, "AWS_SPEED=GAUSSIAN_TOR=0.00", AWS_NL, 0, True, seed=seed)
    steps = 0
    while o.continue_if() and o.time < 6000:
        assert p.network().run_ms(10) == o.run_ms(10)
        steps += 1
        bad = _handel_compare(p, o, f"t={o.time}", full=(steps % 10 == 0))
        assert not bad, bad
    assert not p.continue_if()
    assert not _handel_compare(p, o, "end")
    assert (p.rows(5) != 0).any()  # somebody got blacklisted


@pytest.mark.parametrize("n,thr,down,nb,nl,desync,seed,step,until", [
    (64, 40, 16, NB, NL, 0, None, 5, 2500),
    (128, 100, 20, NB, NL, 50, 9, 1, 1500),
    (1024, 700, 256, "AWS_SPEED=GAUSSIAN_TOR=0.00", AWS_NL, 0, 3, 10, 3000),  # last level is pooled: the injected signature gets a slab
])
def test_handel_hidden_byzantine(n, thr, down, nb, nl, desync, seed, step, until):
    """HiddenByzantine.attack (P/Handel.java:840-917): low-value signatures of down peers injected at the last level."""
    p, o = _handel_pair(n, thr, 4, 50, 10, 20, 10, down, nb, nl, desync, False, seed=seed, hidden=True)
    k = 0
    while o.time < until:
        assert p.network().run_ms(step) == o.run_ms(step)
        k += 1
        bad = _handel_compare(p, o, f"t={o.time}", full=(k % 5 == 0))
        assert not bad, bad
    assert not _handel_compare(p, o, "end")
    # the attack leaves its trace: some honest node counts a down node's signature in its last level
    down_ids = np.flatnonzero(p.network().attrs()["down"])
    inc = p.rows(0)
    assert any(((inc[:, d >> 6] >> np.uint64(d & 63)) & np.uint64(1)).any() for d in down_ids)


def test_handel_512_tor_desync_plain_dead():
    p, o = _handel_pair(512, 450, 4, 50, 10, 20, 10, 51, AWS_NB, AWS_NL, 50, False, seed=1)
    for step in (1, 3, 7, 20, 50, 100, 100, 300, 500, 1000):
        assert p.network().run_ms(step) == o.run_ms(step)
        bad = _handel_compare(p, o, f"t={o.time}")
        assert not bad, bad


def test_error_paths():
    from wittgenstein_b200 import GSFSignature, GSFSignatureParameters, Network, PingPong, PingPongParameters, WtgError

    with pytest.raises(WtgError):
        GSFSignatureParameters(32, 33, 3, 20, 10, 10, 0, NB, NL)
    with pytest.raises(WtgError):
        GSFSignature(GSFSignatureParameters(32, 30, 3, 20, 10, 10, 0, "NOPE_SPEED=CONSTANT_TOR=0.00", NL))
    with pytest.raises(WtgError):
        PingPong(PingPongParameters(10, None, "NoSuchLatency"))
    p = PingPong(PingPongParameters(10, None, None))
    with pytest.raises(WtgError):
        p.network().run_ms(5)  # not initialised
    p.init()
    with pytest.raises(WtgError):
        p.network().run_ms(0)  # Network.java:319-321
    with pytest.raises(WtgError):
        p.network().set_seed(1)  # after init
    g = GSFSignature(GSFSignatureParameters(48, 40, 3, 20, 10, 10, 0, NB, NL))
    with pytest.raises(WtgError):
        g.init()  # power-of-two node counts only on the device engine
    g = GSFSignature(GSFSignatureParameters(64, 60, 3, 20, 10, 10, 0, NB, AWS_NL))
    with pytest.raises(WtgError):
        g.init()  # AWS latency with non-AWS builder (NetworkLatency.java:146-148)
    from wittgenstein_b200 import SanFerminSignature, SanFerminSignatureParameters
    with pytest.raises(WtgError):
        SanFerminSignature(SanFerminSignatureParameters(1000, 1000, 2, 48, 300, 1))  # power of two only on the device
    with pytest.raises(WtgError):
        SanFerminSignature(SanFerminSignatureParameters(1024, 1024, 2, 48, 300, 64))  # candidateCount + 1 destinations per request: at most 64
    from wittgenstein_b200 import Handel, HandelParameters
    with pytest.raises(WtgError):
        HandelParameters(100, 90)  # Handel.java:118-120
    # a capacity that is too small fails loudly instead of dropping events
    g = GSFSignature(GSFSignatureParameters(256, 250, 3, 20, 10, 10, 0, NB, NL), tunables={"qcap": 32})
    g.init()
    with pytest.raises(WtgError):
        for _ in range(100):
            g.network().run_ms(10)


@pytest.mark.gpu
def test_gsf_stop_start_partition_discard_midrun():
    """Node.stop/start (Node.java:120-127), Network.partition/endPartition (Network.java:693-707) and setMsgDiscardTime
    (:103-106) applied between runMs windows: dropped sends still count as sent (:476-486), stopped nodes neither receive
    (:606) nor re-arm their periodic task, partitions cut deliveries both at send and at arrival time."""
    args = (128, 100, 3, 20, 10, 10, 8, AWS_NB, AWS_NL)
    p = GSFSignature(GSFSignatureParameters(*args))
    o = OracleGSF(*args, seed=5)
    p.network().set_seed(5)
    p.network().set_msg_discard_time(180)
    o.set_msg_discard_time(180)
    p.init(); o.init()
    plan = {5: ("stop", 7), 8: ("stop", 64), 12: ("partition", 0.3), 20: ("partition", 0.7), 30: ("start", 7), 45: ("end_partition", None),
            60: ("start", 64)}
    for k in range(120):
        if k in plan:
            op, arg = plan[k]
            for tgt in (p.network(), o):
                fn = getattr(tgt, {"stop": "stop_node", "start": "start_node"}.get(op, op))
                fn() if arg is None else fn(arg)
        assert p.network().run_ms(10) == o.run_ms(10)
        bad = compare_gsf(p, o, f"t={o.time}", full=(k % 4 == 0))
        assert not bad, bad
    assert not compare_gsf(p, o, "end")


@pytest.mark.gpu
def test_pingpong_discard_and_partition():
    p = PingPong(PingPongParameters(1000, AWS_NB, AWS_NL))
    o = OraclePingPong(1000, AWS_NB, AWS_NL)
    for tgt in (p.network(), o):
        tgt.set_msg_discard_time(120)
    p.init(); o.init()
    for k in range(12):
        if k == 1:
            p.network().partition(0.5); o.partition(0.5)
        if k == 3:
            p.network().stop_node(0); o.stop_node(0)  # the pinger itself
        if k == 5:
            p.network().start_node(0); o.start_node(0)
            p.network().end_partition(); o.end_partition()
        assert p.network().run_ms(50) == o.run_ms(50)
        assert (p.pongs() == o.pongs()).all()
        assert (p.network().counters() == o.counters()).all()
        assert p.network().msgs_size() == o.msgs_size()


@pytest.mark.gpu
def test_sanfermin_with_stopped_nodes_and_partition():
    from tests.oracle_lib import OracleSanFermin
    from wittgenstein_b200 import SanFerminSignature, SanFerminSignatureParameters

    p = SanFerminSignature(SanFerminSignatureParameters(256, 256, 2, 48, 300, 1, False, None, None))
    o = OracleSanFermin(256, 256, 2, 48, 300, 1, None, None)
    p.init(); o.init()
    for k in range(150):
        if k == 2:
            for i in (3, 77, 200):
                p.network().stop_node(i); o.stop_node(i)
        if k == 10:
            p.network().partition(0.4); o.partition(0.4)
        if k == 40:
            p.network().end_partition(); o.end_partition()
        assert p.network().run_ms(20) == o.run_ms(20)
        bad = _sf_compare(p, o, f"t={o.time}")
        assert not bad, bad


def _cappos_compare(p, o, tag):
    bad = []
    if p.network().rng_state() != o.rng_state():
        bad.append(f"{tag}: rd state")
    if p.network().msgs_size() != o.msgs_live():
        bad.append(f"{tag}: msgs.size()")
    if not (p.network().counters() == o.counters()).all():
        bad.append(f"{tag}: counters")
    a, b = p.scalars(), o.scalars()
    for k in a:
        if not (a[k] == b[k]).all():
            bad.append(f"{tag}: {k}")
    return bad


@pytest.mark.gpu
@pytest.mark.parametrize("n,k,nb,nl,seed,step,until,force", [
    (64, 3, None, None, None, 1, 2500, False),
    (1024, 50, AWS_NB, AWS_NL, 4, 10, 4000, False),          # the shipped candidateCount (SanFerminCappos.java:79, 471)
    (512, 7, None, None, 2, 10, 4000, True),                   # draw indices re-derived serially on every tick
    (4096, 50, None, None, None, 20, 5000, False),
])
def test_cappos_parity(n, k, nb, nl, seed, step, until, force):
    """SanFerminCappos (swaps with / without reply, cached levels, timeouts, candidateCount-wide shuffled requests) vs the oracle."""
    from tests.oracle_lib import OracleCappos
    from wittgenstein_b200 import SanFerminCappos, SanFerminCapposParameters

    p = SanFerminCappos(SanFerminCapposParameters(n, n // 2, 2, 48, 150, k, nb, nl), tunables={"force_shuffle_serial": 1} if force else None)
    o = OracleCappos(n, n // 2, 2, 48, 150, k, nb, nl, seed=seed)
    if seed is not None:
        p.network().set_seed(seed)
    p.init(); o.init()
    assert not _cappos_compare(p, o, "init")
    while o.time < until:
        assert p.network().run_ms(step) == o.run_ms(step)
        bad = _cappos_compare(p, o, f"t={o.time}")
        assert not bad, bad
    assert p.scalars()["done"].sum() > n // 2


@pytest.mark.gpu
def test_casper_end_partition_resends_heads():
    """BlockChainNetwork.endPartition (C/BlockChainNetwork.java:46-54): after the partition every node sends its head to
    everybody — N sendAll calls issued by the caller, one draw each, in node order."""
    from tests.parity import compare_casper

    p, o = _casper_pair(3, 3, 16, None, None, 2, 0, 160000)
    for k in range(160):
        if k == 30:
            p.network().partition(0.45); o.partition(0.45)
        if k == 70:
            p.network().end_partition(); o.end_partition()
        if k == 110:
            p.network().partition(0.2); o.partition(0.2)
        if k == 115:
            p.network().end_partition(); o.end_partition()
        assert p.network().run_ms(1000) == o.run_ms(1000)
        bad = compare_casper(p, o, f"t={o.time}")
        assert not bad, bad
    assert not compare_casper(p, o, "end", atts=True)


@pytest.mark.gpu
def test_pingpong_sends_from_the_host():
    """network.send(msg, from, to) / send(msg, from, dests) called between two windows (C/Network.java:353-366), like the
    reference's NetworkTest drives its network: one draw per call, arrival through the latency model, LIFO insertion."""
    p = PingPong(PingPongParameters(200, AWS_NB, AWS_NL))
    o = OraclePingPong(200, AWS_NB, AWS_NL)
    p.init(); o.init()
    plan = {0: [(1, 5, [7])], 2: [(1, 9, [3, 4, 5, 150, 9]), (2, 11, [12])], 3: [(1, 0, list(range(20, 36)))], 7: [(2, 199, [0, 1])]}
    for k in range(15):
        for (typ, frm, to) in plan.get(k, []):
            p.network().send(typ, frm, to if len(to) > 1 else to[0])
            o.send(typ, frm, to)
            assert p.network().rng_state() == o.rng_state()
            assert p.network().msgs_size() == o.msgs_size()
        assert p.network().run_ms(40) == o.run_ms(40)
        assert (p.pongs() == o.pongs()).all()
        assert (p.network().counters() == o.counters()).all()
        assert p.network().msgs_size() == o.msgs_size()


@pytest.mark.gpu
def test_pingpong_delayed_multi_sends_from_the_host():
    """send(msg, sendTime, from, dests, delaysBetweenMessage) — C/Network.java:420-467, the MultipleDestWithDelayEnvelope case of
    CT/NetworkTest.java:122-188: destination i is sent delay + 1 ms after destination i - 1, arrivals sorted stably."""
    p = PingPong(PingPongParameters(300, None, None))
    o = OraclePingPong(300, None, None)
    p.init(); o.init()
    for k in range(12):
        if k == 1:
            for tgt in (p.network(), o):
                tgt.send(1, 4, [10, 11, 12, 13, 200, 7], send_time=tgt.time + 5, delay_between=10)
        if k == 2:
            for tgt in (p.network(), o):
                tgt.send(2, 9, [1, 2, 3], send_time=tgt.time + 1, delay_between=49)   # crosses several windows
                tgt.send(1, 17, 18, send_time=tgt.time + 30)
        if k == 4:
            for tgt in (p.network(), o):
                tgt.send(1, 0, list(range(100, 116)), send_time=tgt.time + 2, delay_between=3)
        assert p.network().rng_state() == o.rng_state()
        assert p.network().run_ms(50) == o.run_ms(50)
        assert (p.pongs() == o.pongs()).all()
        assert (p.network().counters() == o.counters()).all()
        assert p.network().msgs_size() == o.msgs_size()
    with pytest.raises(Exception):
        p.network().send(1, 0, [1, 2], send_time=p.network().time)  # sendTime <= time (Network.java:470-473)


@pytest.mark.parametrize("latency,measured", [("EthScanNetworkLatency", False), (None, True), ("NetworkFixedLatency(8000)", False),
                                              ("NetworkUniformLatency(8000)", False)])
def test_pingpong_far_latencies(latency, measured):
    """NetworkLatency.java:277-383 (Measured, EthScan with its doubled extraLatency) and Fixed / Uniform(8000) on the device:
    arrivals >= 2 048 ms ahead go through the far-future calendar (time ring fixed at 4 096 buckets)."""
    from tests.test_far_latency_emu import check_pingpong, pingpong_pair

    p, o = pingpong_pair(latency, measured, None)
    assert p.network().stats()["ring"] == 4096
    check_pingpong(p, o, 30, 1000)
    assert o.pongs()[0] > 0


def test_gsf_256_ethscan_and_measured():
    """GSFSignature (conditional tasks: ticks every millisecond) over EthScan / Measured latencies: far-future calendar under GSF"""
    from tests.test_far_latency_emu import MEASURED

    for latency, measured in (("EthScanNetworkLatency", False), (None, True)):
        args = (256, 204, 4, 50, 20, 10, 25, "RANDOM_SPEED=CONSTANT_TOR=0.00", latency)
        p = GSFSignature(GSFSignatureParameters(*args))
        o = OracleGSF(*args)
        if measured:
            p.network().set_network_latency_measured(*MEASURED)
            o.set_network_latency_measured(*MEASURED)
        p.init(); o.init()
        for i in range(80):
            assert p.network().run_ms(100) == o.run_ms(100)
            bad = compare_gsf(p, o, f"t={o.time}", full=(i % 8 == 0))
            assert not bad, bad


@pytest.mark.parametrize("nb,nl", [("CITIES_SPEED=CONSTANT_TOR=0.00", "NetworkLatencyByCityWJitter"), ("CITIES_SPEED=GAUSSIAN_TOR=0.33", "NetworkLatencyByCity")])
def test_pingpong_cities(nb, nl):
    """CITIES node builder + NetworkLatencyByCity[WJitter] (NetworkLatency.java:159-233, NodeBuilder.java:98-148) on the device"""
    from wittgenstein_b200 import PingPong, PingPongParameters

    p = PingPong(PingPongParameters(1000, nb, nl))
    o = OraclePingPong(1000, nb, nl)
    p.init(); o.init()
    a, b = p.network().attrs(), o.attrs()
    for k in ("x", "y", "extra", "city", "down"):
        assert (a[k] == b[k]).all(), k
    for _ in range(20):
        assert p.network().run_ms(100) == o.run_ms(100)
        assert (p.pongs() == o.pongs()).all()
        assert (p.network().counters() == o.counters()).all()
        assert p.network().msgs_size() == o.msgs_size()
    assert p.pongs()[0] > 0


def test_handel_2048_default_scenario_cities():
    """HandelScenarios.defaultParams (HandelScenarios.java:65-120): 2 048 nodes, CITIES builder, NetworkLatencyByCityWJitter —
    to completion (Handel.newContIf), every 10 ms against the oracle"""
    from tests.test_cities_emu import handel_default_params

    args = handel_default_params(2048)
    p, o = _handel_pair(*args)
    steps = 0
    while o.continue_if() and steps < 600:
        assert p.network().run_ms(10) == o.run_ms(10)
        steps += 1
        bad = _handel_compare(p, o, f"t={o.time}", full=(steps % 20 == 0))
        assert not bad, bad
    assert not o.continue_if()
    assert not _handel_compare(p, o, "end")
