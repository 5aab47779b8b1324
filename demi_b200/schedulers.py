"""Host-side mirror of the reference's scheduler plugin surface for the hot path.

Same names, argument meaning and error behaviour as the Scala classes; the work
is done by the CUDA engine behind the C ABI (include/demi_b200.h).  There is no
CPU path here: constructing an engine without a CUDA device raises.

  SchedulerConfig   <- SchedulerConfig.scala:9-37
  RandomScheduler   <- schedulers/RandomScheduler.scala:41 (explore :234, test :597,
                       setMaxMessages :55, setInvariant :521)
"""
import ctypes as C

import numpy as np

from . import _native as N
from .events import pack_externals


class DemiError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "demi_b200 error %d: %s" % (code, msg))
        self.code = code


class SchedulerConfig(object):
    """SchedulerConfig.scala:9-37, restricted to what has meaning without a JVM.
    `model`/`model_flags` stand in for the application under test; `invariant_check`
    is the violation code looked for by the model's built-in invariant (None = any)."""

    def __init__(self, model, model_flags=0, device=0, ignoreTimers=False, blocked_mask=0):
        self.model, self.model_flags, self.device = model, model_flags, device
        self.ignoreTimers, self.blocked_mask = ignoreTimers, blocked_mask


class Engine(object):
    """Owns one demi_handle."""

    def __init__(self, schedulerConfig):
        self.cfg = schedulerConfig
        self._h = C.c_void_p()
        c = N.Config(schedulerConfig.device, schedulerConfig.model, schedulerConfig.model_flags,
                     schedulerConfig.blocked_mask, 1 if schedulerConfig.ignoreTimers else 0)
        rc = N.lib().demi_create(C.byref(c), C.byref(self._h))
        if rc != N.OK:
            raise DemiError(rc, N.lib().demi_last_error(None).decode())
        self._ext = None

    def close(self):
        if self._h:
            N.lib().demi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != N.OK:
            raise DemiError(rc, N.lib().demi_last_error(self._h).decode())

    def set_externals(self, events):
        arr = events if isinstance(events, np.ndarray) else pack_externals(events)
        arr = np.ascontiguousarray(arr, dtype=N.EXT_DTYPE)
        self._ext = arr
        self._check(N.lib().demi_set_externals(self._h, arr.ctypes.data, len(arr)))

    def fuzz_batch(self, seed_base, n, max_messages, interval, looking_for=0, out=None):
        p = N.FuzzParams(seed_base, n, max_messages, interval, looking_for or 0, 0)
        if out is None:
            out = np.empty(n, dtype=N.RESULT_DTYPE)
        self._check(N.lib().demi_fuzz_batch(self._h, C.byref(p), out.ctypes.data))
        return out

    def fuzz_batch_dev(self, seed_base, n, max_messages, interval, out_ptr, stream=0, looking_for=0):
        p = N.FuzzParams(seed_base, n, max_messages, interval, looking_for or 0, 0)
        self._check(N.lib().demi_fuzz_batch_dev(self._h, C.byref(p), C.c_void_p(out_ptr), C.c_void_p(stream)))

    def fuzz_summary_dev(self, stream=0):
        nv, ss = C.c_uint64(), C.c_uint64()
        self._check(N.lib().demi_fuzz_summary_dev(self._h, None, 0, C.c_void_p(stream), C.byref(nv), C.byref(ss)))
        return nv.value, ss.value

    def fuzz_trace(self, seed, max_messages, interval, looking_for=0, cap_events=65536, cap_nodes=65536):
        p = N.FuzzParams(seed, 1, max_messages, interval, looking_for or 0, 0)
        ev = np.zeros(cap_events, dtype=N.EVENT_DTYPE)
        par = np.zeros(cap_nodes, dtype=np.uint16)
        ne, nn = C.c_uint32(), C.c_uint32()
        res = np.zeros(1, dtype=N.RESULT_DTYPE)
        self._check(N.lib().demi_fuzz_trace(self._h, C.byref(p), seed, ev.ctypes.data, cap_events, C.byref(ne),
                                            par.ctypes.data, cap_nodes, C.byref(nn), res.ctypes.data))
        return ev[:ne.value].copy(), par[:nn.value].copy(), res[0]

    def stats(self):
        s = N.Perf()
        self._check(N.lib().demi_stats(self._h, C.byref(s)))
        return s


class RandomScheduler(object):
    """RandomScheduler(schedulerConfig, max_executions, invariant_check_interval,
    randomizationStrategy = FullyRandom(seed)) — RandomScheduler.scala:41-44.

    The reference explores executions one after another; here execution i of
    `max_executions` is prefix `seed + i` of one batched kernel launch, and
    explore() returns the first violating one in that order."""

    def __init__(self, schedulerConfig, max_executions=1, invariant_check_interval=0, seed=0, engine=None):
        self.schedulerConfig = schedulerConfig
        self.max_executions = max_executions
        self.invariant_check_interval = invariant_check_interval
        self.seed = seed
        self.maxMessages = -1            # Int.MaxValue (RandomScheduler.scala:54)
        self.test_invariant = True       # models carry their invariant; None => "Must invoke setInvariant"
        self.engine = engine or Engine(schedulerConfig)
        self.stats = None
        self.last_results = None

    def setMaxMessages(self, n):
        self.maxMessages = n

    def setInvariant(self, invariant):
        self.test_invariant = invariant

    def getName(self):
        return "RandomScheduler"

    def explore(self, trace, lookingFor=None):
        """Returns (EventTrace records, violation code) of the first violating
        execution, else None (RandomScheduler.scala:234-272)."""
        if self.test_invariant is None:
            raise ValueError("Must invoke setInvariant before test()")   # :244-246
        self.engine.set_externals(trace)
        res = self.engine.fuzz_batch(self.seed, self.max_executions, self.maxMessages,
                                     self.invariant_check_interval, lookingFor or 0)
        self.last_results = res
        if self.stats is not None:
            self.stats.increment_replays(self.max_executions)
        bad = np.nonzero(res["status"])[0]
        if len(bad):
            raise DemiError(N.ERR_CAPACITY, "prefix %d overflowed (status %d)" % (bad[0], res["status"][bad[0]]))
        hits = np.nonzero(res["violation"])[0]
        if not len(hits):
            return None
        i = int(hits[0])
        ev, _, r = self.engine.fuzz_trace(self.seed + i, self.maxMessages, self.invariant_check_interval,
                                          lookingFor or 0)
        return ev, int(r["violation"])

    def test(self, events, violation_fingerprint, stats=None):
        """TestOracle.test: Some(trace) iff the violation was reproduced (RandomScheduler.scala:597-612)."""
        self.stats = stats
        r = self.explore(events, violation_fingerprint)
        return None if r is None else r[0]
