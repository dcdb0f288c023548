"""ctypes binding of the CPU oracle (oracle/libwtg_oracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (wittgenstein_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(ORACLE_DIR, "libwtg_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".hpp", ".cpp"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        build()
    lib = C.CDLL(so)
    lib.wo_last_error.restype = C.c_char_p
    lib.wo_lcg_advance.restype = C.c_uint64
    lib.wo_lcg_advance.argtypes = [C.c_uint64, C.c_uint64]
    lib.wo_random_next_double.restype = C.c_double
    lib.wo_random_next_double.argtypes = [C.c_int64, C.c_int]
    lib.wo_gpd_inverse.restype = C.c_double
    lib.wo_gpd_inverse.argtypes = [C.c_double] * 4
    lib.wo_latency.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.wo_handel_create.restype = C.c_void_p
    lib.wo_handel_create.argtypes = [C.POINTER(C.c_int), C.c_char_p, C.c_char_p]
    lib.wo_handel_rng_state.restype = C.c_uint64
    lib.wo_handel_msgs_live.restype = C.c_int64
    lib.wo_handel_run_timed.restype = C.c_double
    lib.wo_handel_run_timed.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.wo_casper_create.restype = C.c_void_p
    lib.wo_casper_create.argtypes = [C.POINTER(C.c_int), C.c_char_p, C.c_char_p]
    lib.wo_casper_rng_state.restype = C.c_uint64
    lib.wo_casper_msgs_live.restype = C.c_int64
    lib.wo_casper_deliveries.restype = C.c_int64
    lib.wo_casper_run_timed.restype = C.c_double
    lib.wo_casper_run_timed.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.wo_casper_partition.argtypes = [C.c_void_p, C.c_float]
    lib.wo_cappos_create.restype = C.c_void_p
    lib.wo_cappos_create.argtypes = [C.POINTER(C.c_int), C.c_char_p, C.c_char_p]
    lib.wo_cappos_rng_state.restype = C.c_uint64
    lib.wo_cappos_msgs_live.restype = C.c_int64
    lib.wo_sf_create.restype = C.c_void_p
    lib.wo_sf_create.argtypes = [C.c_int] * 6 + [C.c_char_p, C.c_char_p]
    lib.wo_sf_rng_state.restype = C.c_uint64
    lib.wo_sf_msgs_live.restype = C.c_int64
    for name in ("wo_pp_create", "wo_gsf_create"):
        getattr(lib, name).restype = C.c_void_p
    lib.wo_pp_create.argtypes = [C.c_int, C.c_char_p, C.c_char_p]
    lib.wo_pp_rng_state.restype = C.c_uint64
    lib.wo_gsf_create.argtypes = [C.c_int] * 7 + [C.c_char_p, C.c_char_p]
    lib.wo_gsf_run_timed.restype = C.c_double
    lib.wo_gsf_run_timed.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.wo_gsf_rng_state.restype = C.c_uint64
    lib.wo_gsf_msgs_live.restype = C.c_int64
    _lib = lib
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _b(s):
    return None if s is None else s.encode()


class _NetCtl:
    """Node.stop/start, Network.partition/endPartition/setMsgDiscardTime on the oracle's network (prefix set by the subclass)."""

    _ctl = None

    def set_network_latency_measured(self, proportions, values):
        """network.setNetworkLatency(int[] distribProp, int[] distribVal) (Network.java:665-667), before init()"""
        p = np.asarray(proportions, np.int32)
        v = np.asarray(values, np.int32)
        fn = getattr(self.lib, self._ctl.replace("net_ctl", "set_latency_measured"))
        if fn(self.h, _p(p, C.c_int), _p(v, C.c_int), len(p)) != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())

    def _c(self, op, arg=0):
        if getattr(self.lib, self._ctl)(self.h, op, int(arg)) != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())

    def peek_messages(self, cap=1 << 16):
        """network.msgs.peekMessages() (Network.java:279-286): (total, dict of arrays), rows sorted by (arrivingAt, from, to)"""
        a = [np.zeros(cap, np.int32) for _ in range(5)]
        fn = getattr(self.lib, self._ctl.replace("net_ctl", "peek_messages"))
        total = fn(self.h, *[_p(x, C.c_int32) for x in a], int(cap))
        k = min(total, cap)
        return total, dict(zip(["from", "to", "sent_at", "arriving_at", "is_task"], [x[:k] for x in a]))

    def stop_node(self, i):
        self._c(0, i)

    def start_node(self, i):
        self._c(1, i)

    def partition(self, part):
        self._c(2, round(part * 10000))

    def end_partition(self):
        self._c(3)

    def set_msg_discard_time(self, ms):
        self._c(4, ms)


class OraclePingPong(_NetCtl):
    _ctl = "wo_pp_net_ctl"

    """protocols/PingPong.java through the oracle."""

    def __init__(self, node_ct=1000, node_builder=None, latency=None, seed=None):
        self.lib = load()
        self.n = node_ct
        self.h = C.c_void_p(self.lib.wo_pp_create(node_ct, _b(node_builder), _b(latency)))
        if not self.h:
            raise ValueError(self.lib.wo_last_error().decode())
        if seed is not None:
            self.lib.wo_pp_set_seed(self.h, C.c_int64(seed))

    def init(self):
        if self.lib.wo_pp_init(self.h) != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())

    def run_ms(self, ms):
        r = self.lib.wo_pp_run_ms(self.h, ms)
        if r < 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        return bool(r)

    @property
    def time(self):
        return self.lib.wo_pp_time(self.h)

    def msgs_size(self):
        return self.lib.wo_pp_msgs_size(self.h)

    def rng_state(self):
        return int(self.lib.wo_pp_rng_state(self.h))

    def send(self, msg_type, from_id, to, send_time=None, delay_between=0):
        dests = np.asarray([to] if np.isscalar(to) else list(to), np.int32)
        if send_time is None:
            rc = self.lib.wo_pp_send(self.h, int(msg_type), int(from_id), _p(dests, C.c_int32), len(dests))
        else:
            rc = self.lib.wo_pp_send_at(self.h, int(msg_type), int(from_id), _p(dests, C.c_int32), len(dests), int(send_time), int(delay_between))
        if rc != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())

    def pongs(self):
        out = np.zeros(self.n, np.int32)
        self.lib.wo_pp_pongs(self.h, _p(out, C.c_int32))
        return out

    def counters(self):
        out = np.zeros((5, self.n), np.int64)
        self.lib.wo_pp_node_counters(self.h, _p(out, C.c_int64))
        return out

    def attrs(self):
        x = np.zeros(self.n, np.int32); y = np.zeros(self.n, np.int32); e = np.zeros(self.n, np.int32)
        c = np.zeros(self.n, np.int32); s = np.zeros(self.n, np.float64); d = np.zeros(self.n, np.uint8)
        self.lib.wo_pp_node_attrs(self.h, _p(x, C.c_int32), _p(y, C.c_int32), _p(e, C.c_int32), _p(c, C.c_int32), _p(s, C.c_double), _p(d, C.c_uint8))
        return dict(x=x, y=y, extra=e, city=c, speed=s, down=d)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.wo_pp_destroy(self.h)
            self.h = None


class OracleGSF(_NetCtl):
    _ctl = "wo_gsf_net_ctl"

    """protocols/GSFSignature.java through the oracle."""

    def __init__(self, node_count, threshold, pairing_time, timeout_per_level_ms, period_ms, accelerated_calls, nodes_down,
                 node_builder, latency, seed=None):
        self.lib = load()
        self.n = node_count
        self.params = (node_count, threshold, pairing_time, timeout_per_level_ms, period_ms, accelerated_calls, nodes_down)
        self.h = C.c_void_p(self.lib.wo_gsf_create(*self.params, _b(node_builder), _b(latency)))
        if not self.h:
            raise ValueError(self.lib.wo_last_error().decode())
        if seed is not None:
            self.lib.wo_gsf_set_seed(self.h, C.c_int64(seed))
        self.words = (node_count + 63) // 64

    def init(self):
        if self.lib.wo_gsf_init(self.h) != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        self.L = self.lib.wo_gsf_levels(self.h, 1)

    def init_fast(self, threads):
        if self.lib.wo_gsf_init_fast(self.h, int(threads)) != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        self.L = self.lib.wo_gsf_levels(self.h, 1)

    def run_ms(self, ms):
        r = self.lib.wo_gsf_run_ms(self.h, ms)
        if r < 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        return bool(r)

    def run_timed(self, ms, steps):
        t = self.lib.wo_gsf_run_timed(self.h, ms, steps)
        if t < 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        return t

    @property
    def time(self):
        return self.lib.wo_gsf_time(self.h)

    def msgs_size(self):
        return self.lib.wo_gsf_msgs_size(self.h)

    def msgs_live(self):
        return self.lib.wo_gsf_msgs_live(self.h)

    def continue_if(self):
        return bool(self.lib.wo_gsf_continue_if(self.h))

    def counters(self):
        out = np.zeros((5, self.n), np.int64)
        self.lib.wo_gsf_node_counters(self.h, _p(out, C.c_int64))
        return out

    def attrs(self):
        x = np.zeros(self.n, np.int32); y = np.zeros(self.n, np.int32); e = np.zeros(self.n, np.int32)
        c = np.zeros(self.n, np.int32); s = np.zeros(self.n, np.float64); d = np.zeros(self.n, np.uint8)
        self.lib.wo_gsf_node_attrs(self.h, _p(x, C.c_int32), _p(y, C.c_int32), _p(e, C.c_int32), _p(c, C.c_int32), _p(s, C.c_double), _p(d, C.c_uint8))
        return dict(x=x, y=y, extra=e, city=c, speed=s, down=d)

    def scalars(self):
        a = [np.zeros(self.n, np.int32) for _ in range(5)]
        self.lib.wo_gsf_node_scalars(self.h, *[_p(v, C.c_int32) for v in a])
        return dict(pairing=a[0], sig_checked=a[1], sig_queue_size=a[2], to_verify=a[3], card=a[4])

    def verified(self):
        out = np.zeros((self.n, self.words), np.uint64)
        self.lib.wo_gsf_verified(self.h, _p(out, C.c_uint64), self.words)
        return out

    def level_rows(self, which):
        out = np.zeros((self.n, self.words), np.uint64)
        self.lib.wo_gsf_level_rows(self.h, which, _p(out, C.c_uint64), self.words)
        return out

    def level_scalars(self):
        L = self.L
        a = [np.zeros((self.n, L), np.int32) for _ in range(3)]
        self.lib.wo_gsf_level_scalars(self.h, L, *[_p(v, C.c_int32) for v in a])
        return dict(pos=a[0], remaining=a[1], card=a[2])

    def peers(self, node, level):
        cap = max(1, self.n)
        out = np.zeros(cap, np.int32)
        k = self.lib.wo_gsf_peers(self.h, node, level, _p(out, C.c_int32), cap)
        return out[:k].copy()

    def rng_state(self):
        return int(self.lib.wo_gsf_rng_state(self.h))

    def stats(self):
        out = np.zeros(12, np.int64)
        self.lib.wo_gsf_stats(self.h, _p(out, C.c_int64))
        keys = ["deliveries", "tasks", "cond_runs", "draws", "eval_entries", "eval_bytes", "updates", "cycles", "sends",
                "multi_sends", "send_bytes", "max_queue"]
        return dict(zip(keys, out.tolist()))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.wo_gsf_destroy(self.h)
            self.h = None


class OracleSanFermin(_NetCtl):
    _ctl = "wo_sf_net_ctl"

    """protocols/SanFerminSignature.java through the oracle (nodes are built by the constructor)."""

    def __init__(self, node_count, threshold, pairing_time, signature_size, reply_timeout, candidate_count, node_builder, latency):
        self.lib = load()
        self.n = node_count
        self.h = C.c_void_p(self.lib.wo_sf_create(node_count, threshold, pairing_time, signature_size, reply_timeout, candidate_count,
                                                  _b(node_builder), _b(latency)))
        if not self.h:
            raise ValueError(self.lib.wo_last_error().decode())

    def set_seed(self, s):
        self.lib.wo_sf_set_seed(self.h, C.c_int64(s))

    def init(self):
        if self.lib.wo_sf_init(self.h) != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())

    def run_ms(self, ms):
        r = self.lib.wo_sf_run_ms(self.h, ms)
        if r < 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        return bool(r)

    @property
    def time(self):
        return self.lib.wo_sf_time(self.h)

    def msgs_live(self):
        return self.lib.wo_sf_msgs_live(self.h)

    def rng_state(self):
        return int(self.lib.wo_sf_rng_state(self.h))

    def counters(self):
        out = np.zeros((5, self.n), np.int64)
        self.lib.wo_sf_node_counters(self.h, _p(out, C.c_int64))
        return out

    def attrs(self):
        x = np.zeros(self.n, np.int32); y = np.zeros(self.n, np.int32); e = np.zeros(self.n, np.int32)
        c = np.zeros(self.n, np.int32); s = np.zeros(self.n, np.float64); d = np.zeros(self.n, np.uint8)
        self.lib.wo_sf_node_attrs(self.h, _p(x, C.c_int32), _p(y, C.c_int32), _p(e, C.c_int32), _p(c, C.c_int32), _p(s, C.c_double), _p(d, C.c_uint8))
        return dict(x=x, y=y, extra=e, city=c, speed=s, down=d)

    def scalars(self):
        a = [np.zeros(self.n, np.int32) for _ in range(7)]
        t = np.zeros(self.n, np.int64)
        self.lib.wo_sf_node_scalars(self.h, *[_p(v, C.c_int32) for v in a], _p(t, C.c_int64))
        keys = ["agg", "cpl", "done", "threshold_done", "sent_requests", "received_requests", "swapping"]
        d = dict(zip(keys, a))
        d["threshold_at"] = t
        return d

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.wo_sf_destroy(self.h)
            self.h = None


class OracleHandel(_NetCtl):
    _ctl = "wo_handel_net_ctl"

    """protocols/Handel.java through the oracle."""

    def __init__(self, node_count, threshold, pairing_time, level_wait_time, extra_cycle, dissemination_period_ms, fast_path,
                 nodes_down, node_builder, latency, desynchronized_start=0, byzantine_suicide=False, seed=None, hidden_byzantine=False):
        self.lib = load()
        self.n = node_count
        arr = np.array([node_count, threshold, pairing_time, level_wait_time, extra_cycle, dissemination_period_ms, fast_path,
                        nodes_down, desynchronized_start, 1 if byzantine_suicide else 0, 1 if hidden_byzantine else 0], np.int32)
        self.h = C.c_void_p(self.lib.wo_handel_create(_p(arr, C.c_int), _b(node_builder), _b(latency)))
        if not self.h:
            raise ValueError(self.lib.wo_last_error().decode())
        if seed is not None:
            self.lib.wo_handel_set_seed(self.h, C.c_int64(seed))
        self.words = max(1, node_count // 64)

    def init(self):
        if self.lib.wo_handel_init(self.h) != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        self.L = self.lib.wo_handel_levels(self.h)

    def run_ms(self, ms):
        r = self.lib.wo_handel_run_ms(self.h, ms)
        if r < 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        return bool(r)

    def run_timed(self, ms, steps):
        t = self.lib.wo_handel_run_timed(self.h, ms, steps)
        if t < 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        return t

    @property
    def time(self):
        return self.lib.wo_handel_time(self.h)

    def msgs_live(self):
        return self.lib.wo_handel_msgs_live(self.h)

    def rng_state(self):
        return int(self.lib.wo_handel_rng_state(self.h))

    def continue_if(self):
        return bool(self.lib.wo_handel_continue_if(self.h))

    def counters(self):
        out = np.zeros((5, self.n), np.int64)
        self.lib.wo_handel_node_counters(self.h, _p(out, C.c_int64))
        return out

    def attrs(self):
        x = np.zeros(self.n, np.int32); y = np.zeros(self.n, np.int32); e = np.zeros(self.n, np.int32)
        c = np.zeros(self.n, np.int32); s = np.zeros(self.n, np.float64); d = np.zeros(self.n, np.uint8)
        self.lib.wo_handel_node_attrs(self.h, _p(x, C.c_int32), _p(y, C.c_int32), _p(e, C.c_int32), _p(c, C.c_int32), _p(s, C.c_double), _p(d, C.c_uint8))
        return dict(x=x, y=y, extra=e, city=c, speed=s, down=d)

    def scalars(self):
        out = np.zeros((9, self.n), np.int32)
        self.lib.wo_handel_node_scalars(self.h, _p(out, C.c_int32))
        keys = ["start_at", "pairing", "sigs_checked", "sig_queue_size", "msg_filtered", "window", "added_cycle", "total_sig_size", "queued"]
        return {k: out[i] for i, k in enumerate(keys)}

    def rows(self, which):
        out = np.zeros((self.n, self.words), np.uint64)
        self.lib.wo_handel_rows(self.h, which, _p(out, C.c_uint64), self.words)
        return out

    def level_scalars(self):
        a = [np.zeros((self.n, self.L), np.int32) for _ in range(4)]
        self.lib.wo_handel_level_scalars(self.h, self.L, *[_p(v, C.c_int32) for v in a])
        return dict(pos=a[0], outgoing_finished=a[1], suicide_biz_after=a[2], queue=a[3])

    def peers(self, node, level):
        out = np.zeros(max(1, self.n), np.int32)
        k = self.lib.wo_handel_peers(self.h, node, level, _p(out, C.c_int32), self.n)
        return out[:k].copy()

    def ranks(self, node):
        out = np.zeros(self.n, np.int32)
        self.lib.wo_handel_ranks(self.h, node, _p(out, C.c_int32))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.wo_handel_destroy(self.h)
            self.h = None


class OracleCasper(_NetCtl):
    _ctl = "wo_casper_net_ctl"

    """protocols/CasperIMD.java through the oracle; init(byz_delay) = init(new ByzBlockProducerWF(byz_delay, genesis))."""

    def __init__(self, cycle_length, random_on_ties, block_producers_count, attesters_per_round, block_construction_time,
                 attestation_construction_time, node_builder, latency):
        self.lib = load()
        arr = np.array([cycle_length, 1 if random_on_ties else 0, block_producers_count, attesters_per_round,
                        block_construction_time, attestation_construction_time], np.int32)
        self.n = 1 + block_producers_count + attesters_per_round * cycle_length
        self.h = C.c_void_p(self.lib.wo_casper_create(_p(arr, C.c_int), _b(node_builder), _b(latency)))
        if not self.h:
            raise ValueError(self.lib.wo_last_error().decode())

    def __del__(self):
        try:
            self.lib.wo_casper_destroy(self.h)
        except Exception:  # noqa: BLE001
            pass

    def set_seed(self, s):
        self.lib.wo_casper_set_seed(self.h, C.c_int64(s))

    def init(self, byz_delay=0, byz_kind="WF"):
        if self.lib.wo_casper_init_byz(self.h, {"plain": 3, "SF": 4, "NS": 5, "WF": 6}[byz_kind], int(byz_delay)) != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())

    def run_ms(self, ms):
        r = self.lib.wo_casper_run_ms(self.h, ms)
        if r < 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        return bool(r)

    def run_timed(self, ms, step):
        r = self.lib.wo_casper_run_timed(self.h, ms, step)
        if r < 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        return r

    @property
    def time(self):
        return self.lib.wo_casper_time(self.h)

    def msgs_live(self):
        return self.lib.wo_casper_msgs_live(self.h)

    def msgs_size_at(self, t):
        return self.lib.wo_casper_msgs_size_at(self.h, t)

    def rng_state(self):
        return int(self.lib.wo_casper_rng_state(self.h))

    def deliveries(self):
        return int(self.lib.wo_casper_deliveries(self.h))

    def counters(self):
        out = np.zeros((5, self.n), np.int64)
        self.lib.wo_casper_node_counters(self.h, _p(out, C.c_int64))
        return out

    def attrs(self):
        x = np.zeros(self.n, np.int32); y = np.zeros(self.n, np.int32); e = np.zeros(self.n, np.int32)
        c = np.zeros(self.n, np.int32); s = np.zeros(self.n, np.float64); d = np.zeros(self.n, np.uint8)
        self.lib.wo_casper_node_attrs(self.h, _p(x, C.c_int32), _p(y, C.c_int32), _p(e, C.c_int32), _p(c, C.c_int32), _p(s, C.c_double), _p(d, C.c_uint8))
        return dict(x=x, y=y, extra=e, city=c, speed=s, down=d)

    def blocks(self):
        nb = self.lib.wo_casper_block_count(self.h)
        v = [np.zeros(nb, np.int32) for _ in range(5)]
        self.lib.wo_casper_blocks(self.h, *[_p(x, C.c_int32) for x in v])
        return dict(zip(["height", "parent", "producer", "proposal_time", "included"], v))

    def block_attestations(self, block):
        cap = 1 << 16
        while True:
            att, h = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
            k = self.lib.wo_casper_block_attestations(self.h, int(block), _p(att, C.c_int32), _p(h, C.c_int32), cap)
            if k <= cap:
                return sorted(zip(att[:k].tolist(), h[:k].tolist()))
            cap = k

    def node_state(self):
        v = [np.zeros(self.n, np.int32) for _ in range(5)]
        hs = np.zeros(self.n, np.uint64)
        self.lib.wo_casper_node_state(self.h, *[_p(x, C.c_int32) for x in v], _p(hs, C.c_uint64))
        d = dict(zip(["head", "atts_received", "heads_with_atts", "blocks_received", "to_reevaluate"], v))
        d["att_hash"] = hs
        return d

    def byz(self):
        out = np.zeros(9, np.int32)
        self.lib.wo_casper_byz(self.h, _p(out, C.c_int32))
        return dict(zip(["to_send", "h", "late", "on_time", "delay", "on_direct_father", "on_older_ancestor", "inc_not_the_best_father",
                         "skipped"], out.tolist()))


class OracleCappos(_NetCtl):
    """protocols/SanFerminCappos.java through the oracle."""

    _ctl = "wo_cappos_net_ctl"

    def __init__(self, node_count, threshold, pairing_time, signature_size, timeout, candidate_count, node_builder, latency, seed=None):
        self.lib = load()
        self.n = node_count
        arr = np.array([node_count, threshold, pairing_time, signature_size, timeout, candidate_count], np.int32)
        self.h = C.c_void_p(self.lib.wo_cappos_create(_p(arr, C.c_int), _b(node_builder), _b(latency)))
        if not self.h:
            raise ValueError(self.lib.wo_last_error().decode())
        if seed is not None:
            self.lib.wo_cappos_set_seed(self.h, C.c_int64(seed))

    def __del__(self):
        try:
            self.lib.wo_cappos_destroy(self.h)
        except Exception:  # noqa: BLE001
            pass

    def init(self):
        if self.lib.wo_cappos_init(self.h) != 0:
            raise RuntimeError(self.lib.wo_last_error().decode())

    def run_ms(self, ms):
        r = self.lib.wo_cappos_run_ms(self.h, ms)
        if r < 0:
            raise RuntimeError(self.lib.wo_last_error().decode())
        return bool(r)

    @property
    def time(self):
        return self.lib.wo_cappos_time(self.h)

    def msgs_live(self):
        return self.lib.wo_cappos_msgs_live(self.h)

    def rng_state(self):
        return int(self.lib.wo_cappos_rng_state(self.h))

    def counters(self):
        out = np.zeros((5, self.n), np.int64)
        self.lib.wo_cappos_node_counters(self.h, _p(out, C.c_int64))
        return out

    def attrs(self):
        x = np.zeros(self.n, np.int32); y = np.zeros(self.n, np.int32); e = np.zeros(self.n, np.int32)
        c = np.zeros(self.n, np.int32); s = np.zeros(self.n, np.float64); d = np.zeros(self.n, np.uint8)
        self.lib.wo_cappos_node_attrs(self.h, _p(x, C.c_int32), _p(y, C.c_int32), _p(e, C.c_int32), _p(c, C.c_int32), _p(s, C.c_double), _p(d, C.c_uint8))
        return dict(x=x, y=y, extra=e, city=c, speed=s, down=d)

    def scalars(self):
        a = [np.zeros(self.n, np.int32) for _ in range(6)]
        t = np.zeros(self.n, np.int64)
        self.lib.wo_cappos_node_scalars(self.h, *[_p(v, C.c_int32) for v in a], _p(t, C.c_int64))
        d = dict(zip(["cpl", "sigs", "done", "threshold_done", "swapping", "cache_mask"], a))
        d["threshold_at"] = t
        return d
