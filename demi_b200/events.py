"""External-event and trace vocabulary of the reference, as flat records.

Mirrors src/main/scala/verification/ExternalEvents.scala:62-91 (Start, Kill,
Send, WaitQuiescence, Partition, UnPartition) and the EventTrace element types
(EventTrace.scala:20; schedulers/AuxilaryTypes.scala:34-73).  Every external
event carries a stable id like UniqueExternalEvent._id (ExternalEvents.scala:14-31).
"""
import itertools

import numpy as np

from . import _native as N

_ids = itertools.count(1)   # IDGenerator (AuxilaryTypes.scala:83-93)


class ExternalEvent(object):
    kind = 0

    def __init__(self, a=0, b=0, type=0, p0=0, p1=0, _id=None):
        self.a, self.b, self.type, self.p0, self.p1 = a, b, type, p0, p1
        self._id = next(_ids) if _id is None else _id

    def label(self):
        return "e%d" % self._id

    def __eq__(self, other):            # UniqueExternalEvent.equals: by id
        return isinstance(other, ExternalEvent) and other._id == self._id

    def __hash__(self):
        return self._id

    def __repr__(self):
        return "%s(a=%d,b=%d,type=%d,p0=%d,p1=%d)#%d" % (
            type(self).__name__, self.a, self.b, self.type, self.p0, self.p1, self._id)


class Start(ExternalEvent):
    kind = N.EXT_START

    def __init__(self, name, **kw):
        ExternalEvent.__init__(self, a=name, **kw)


class Kill(ExternalEvent):
    kind = N.EXT_KILL

    def __init__(self, name, **kw):
        ExternalEvent.__init__(self, a=name, **kw)


class HardKill(ExternalEvent):
    """HardKill(name) (ExternalEvents.scala:69-71): the actor is stopped, not just isolated."""
    kind = N.EXT_HARD_KILL

    def __init__(self, name, **kw):
        ExternalEvent.__init__(self, a=name, **kw)


class Send(ExternalEvent):
    kind = N.EXT_SEND

    def __init__(self, name, type, p0=0, p1=0, **kw):
        ExternalEvent.__init__(self, a=name, type=type, p0=p0, p1=p1, **kw)


class WaitQuiescence(ExternalEvent):
    kind = N.EXT_WAIT_QUIESCENCE

    def __init__(self, **kw):
        ExternalEvent.__init__(self, **kw)


class Partition(ExternalEvent):
    kind = N.EXT_PARTITION

    def __init__(self, a, b, **kw):
        ExternalEvent.__init__(self, a=a, b=b, **kw)


class UnPartition(ExternalEvent):
    kind = N.EXT_UNPARTITION

    def __init__(self, a, b, **kw):
        ExternalEvent.__init__(self, a=a, b=b, **kw)


def pack_externals(events):
    """Seq[ExternalEvent] -> flat demi_ext_event records."""
    arr = np.zeros(len(events), dtype=N.EXT_DTYPE)
    for i, e in enumerate(events):
        arr[i] = (e.kind, e.a, e.b, e.type, e.p0, e.p1, e._id)
    return arr


def unpack_externals(arr):
    cls = {N.EXT_HARD_KILL: HardKill, N.EXT_START: Start, N.EXT_KILL: Kill, N.EXT_SEND: Send, N.EXT_WAIT_QUIESCENCE: WaitQuiescence,
           N.EXT_PARTITION: Partition, N.EXT_UNPARTITION: UnPartition}
    out = []
    for r in arr:
        e = ExternalEvent.__new__(cls[int(r["kind"])])
        ExternalEvent.__init__(e, a=int(r["a"]), b=int(r["b"]), type=int(r["type"]),
                               p0=int(r["p0"]), p1=int(r["p1"]), _id=int(r["id"]))
        out.append(e)
    return out


# ---- canonical external programs of BASELINE.json's configs
def raft5_program(client_cmds=0):
    """Start x5, bootstrap Send to each, optional client commands, WaitQuiescence."""
    ev = [Start(a) for a in range(5)]
    ev += [Send(a, 1, 0x1F) for a in range(5)]          # Raft5::BOOT, membership bitmask
    ev += [Send(i % 5, 2, 1 + i) for i in range(client_cmds)]   # Raft5::CLIENT_CMD
    ev.append(WaitQuiescence())
    return ev


def pingpong3_program(n_sends=100):
    """Start x3, then n_sends Send(X, Ping(k)) with X cycling A,B,C, WaitQuiescence."""
    ev = [Start(a) for a in range(3)]
    ev += [Send(k % 3, 1, k) for k in range(n_sends)]   # PingPong3::PING
    ev.append(WaitQuiescence())
    return ev


def bcast32_program(ttl=3):
    ev = [Start(a) for a in range(32)]
    ev.append(Send(0, 2, ttl))                          # Bcast32::INJECT
    ev.append(WaitQuiescence())
    return ev
