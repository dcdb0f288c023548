"""CPU tests of the CasperIMD restatement (oracle/casper.hpp): the reference's own test assertions, and basic sanity of
the C interface the parity tests use."""
import os
import subprocess

import numpy as np

from tests import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_casper_kat_binary():
    """PT/CasperIMDTest.java (11 cases) + PT/CasperByzantineTest.java (2 cases), restated in oracle/test_casper_kat.cpp."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "kat_casper"])
    out = subprocess.run([os.path.join(ROOT, "oracle", "kat_casper")], capture_output=True, text=True)
    assert out.returncode == 0 and "CASPER KAT OK" in out.stdout, out.stdout + out.stderr


def test_casper_schedule_through_capi():
    """PT/CasperIMDTest.java:24-44 (testInit) through the C interface."""
    o = oracle_lib.OracleCasper(5, False, 5, 80, 1000, 1, None, None)
    o.init(0)
    assert o.n == 1 + 5 + 400
    assert o.msgs_size_at(1) == 0
    for t in (8000, 16000, 24000, 32000, 40000):
        assert o.msgs_size_at(t) == 1
    assert o.msgs_size_at(48000) == 0
    for t in (12000, 20000, 28000, 36000, 44000):
        assert o.msgs_size_at(t) == 80
    assert o.msgs_size_at(52000) == 0


def test_casper_chain_grows_and_everyone_follows():
    o = oracle_lib.OracleCasper(3, False, 3, 10, 1000, 1, None, None)
    o.init(0)
    for _ in range(100):
        o.run_ms(1000)
    b = o.blocks()
    assert len(b["height"]) >= 11 and (np.diff(b["height"][1:]) == 1).all()  # one block per slot, no fork
    st = o.node_state()
    assert (st["head"] >= len(b["height"]) - 2).all()
    assert b["included"][1:].sum() > 0
    assert o.byz()["on_time"] >= 3 and o.byz()["late"] == 0


def test_casper_late_byzantine_block_forks():
    """delay 9000 ms: the Byzantine block of slot k is built after the block of slot k+1 -> two branches, vote counting."""
    o = oracle_lib.OracleCasper(3, False, 3, 20, 1000, 1, None, None)
    o.init(9000)
    for _ in range(200):
        o.run_ms(1000)
    b = o.blocks()
    parents = b["parent"][1:]
    assert len(set(parents.tolist())) < len(parents)  # some block has two children
