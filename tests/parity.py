"""Shared bit-exact comparison of an engine-side protocol (C ABI) with the CPU oracle."""
import numpy as np


def compare_gsf(p, o, tag, full=True):
    """p: wittgenstein_b200.GSFSignature, o: tests.oracle_lib.OracleGSF.  Returns a list of mismatches."""
    bad = []
    net = p.network()
    if net.time != o.time:
        bad.append(f"{tag}: time {net.time} vs {o.time}")
    if net.rng_state() != o.rng_state():
        bad.append(f"{tag}: rd state differs")
    if net.msgs_size() != o.msgs_live():
        bad.append(f"{tag}: msgs.size() {net.msgs_size()} vs {o.msgs_live()}")
    if not (net.counters() == o.counters()).all():
        d = net.counters() != o.counters()
        bad.append(f"{tag}: counters differ rows={np.argwhere(d.any(axis=1)).ravel().tolist()} nodes={np.argwhere(d.any(axis=0))[:5].ravel().tolist()}")
    s1, s2 = p.scalars(), o.scalars()
    for k in s1:
        if not (s1[k] == s2[k]).all():
            bad.append(f"{tag}: {k} differs at nodes {np.argwhere(s1[k] != s2[k])[:5].ravel().tolist()}")
    if full:
        if not (p.verified() == o.verified()).all():
            bad.append(f"{tag}: verifiedSignatures differ")
        l1, l2 = p.level_scalars(), o.level_scalars()
        for k in l1:
            if not (l1[k] == l2[k]).all():
                bad.append(f"{tag}: level {k} differs")
        for w in (1, 2):
            if not (p.rows(w) == o.level_rows(w)).all():
                bad.append(f"{tag}: level rows {w} differ")
    return bad


def compare_init(p, o):
    bad = []
    a, b = p.network().attrs(), o.attrs()
    for k in a:
        if not (a[k] == b[k]).all():
            bad.append(f"init: attr {k} differs")
    n = p.params.node_count
    for node in sorted(set([0, 1, 2, n // 3, n // 2, n - 2, n - 1])):
        for l in range(p.levels):
            if not (p.peers(node, l) == o.peers(node, l)).all():
                bad.append(f"init: peers of node {node} level {l} differ")
    return bad


def run_lockstep(p, o, step, until, full_every=1):
    """runMs(step) on both until `until`; identical slicing on both sides (SURVEY.md A.1 rule 2)."""
    i = 0
    while p.network().time < until:
        r1 = p.network().run_ms(step)
        r2 = o.run_ms(step)
        assert r1 == r2, f"runMs return differs at t={o.time}"
        i += 1
        bad = compare_gsf(p, o, f"t={o.time}", full=(i % full_every == 0))
        assert not bad, bad


def compare_casper(p, o, tag, atts=False):
    """p: wittgenstein_b200.CasperIMD, o: tests.oracle_lib.OracleCasper.  Returns a list of mismatches."""
    bad = []
    net = p.network()
    if net.time != o.time:
        bad.append(f"{tag}: time {net.time} vs {o.time}")
    if net.rng_state() != o.rng_state():
        bad.append(f"{tag}: rd state differs")
    if net.msgs_size() != o.msgs_live():
        bad.append(f"{tag}: msgs.size() {net.msgs_size()} vs {o.msgs_live()}")
    if not (net.counters() == o.counters()).all():
        d = net.counters() != o.counters()
        bad.append(f"{tag}: counters differ rows={np.argwhere(d.any(axis=1)).ravel().tolist()} nodes={np.argwhere(d.any(axis=0))[:5].ravel().tolist()}")
    a, b = p.node_state(), o.node_state()
    for k in a:
        if not (a[k] == b[k]).all():
            bad.append(f"{tag}: node {k} differs at {np.argwhere(a[k] != b[k])[:5].ravel().tolist()}")
    a, b = p.blocks(), o.blocks()
    for k in a:
        if len(a[k]) != len(b[k]) or not (a[k] == b[k]).all():
            bad.append(f"{tag}: blocks {k} differ")
    if p.byz() != o.byz():
        bad.append(f"{tag}: byzantine producer {p.byz()} vs {o.byz()}")
    if atts and not bad:
        for i in range(1, len(a["height"])):
            if p.block_attestations(i) != o.block_attestations(i):
                bad.append(f"{tag}: attestations of block {i} differ")
    return bad


def state_digest(arrays):
    """blake2b over a list of numpy arrays (shape, dtype and bytes)"""
    import hashlib

    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode() + str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def handel_digests(x, is_oracle):
    """digests of a Handel run's state: x is wittgenstein_b200.Handel or tests.oracle_lib.OracleHandel"""
    net = x if is_oracle else x.network()
    out = {"rng": int(net.rng_state()), "msgs": int(net.msgs_live() if is_oracle else net.msgs_size()),
           "counters": state_digest([net.counters()])}
    sc = x.scalars()
    out["scalars"] = state_digest([sc[k] for k in sorted(sc)])
    out["rows"] = state_digest([x.rows(w) for w in range(6)])
    lv = x.level_scalars()
    out["levels"] = state_digest([lv[k] for k in ("pos", "outgoing_finished", "suicide_biz_after")])
    return out
