"""Flat on-disk experiment format.

The reference persists experiments as Java-serialized object streams
(Serialization.scala:57-74, :176-254: event_trace.bin, original_externals.bin, violation.bin, mcs.bin,
minimization_stats.json, ...), which only a JVM with the application's classes can read.  Here every
file is a raw little-endian array of the C-ABI records (include/demi_b200.h), so a JVM host can map
them with a ByteBuffer and this package with numpy:

    <dir>/externals.bin          demi_ext_event[ ]   (16 B each)   original_externals
    <dir>/event_trace.bin        demi_event[ ]       (16 B each)   the recorded EventTrace
    <dir>/dep_parent.bin         uint16[ ]                          DepTracker tree (optional)
    <dir>/mcs.bin                uint64[ ]                          MCS as a mask over externals (optional)
    <dir>/meta.json              model, flags, seed, bounds, violation code, MinimizationStats
"""
import json
import os

import numpy as np

from . import _native as N


def save_experiment(path, model, model_flags, externals, events, violation, dep_parent=None, mcs=None, **meta):
    os.makedirs(path, exist_ok=True)
    np.ascontiguousarray(externals, dtype=N.EXT_DTYPE).tofile(os.path.join(path, "externals.bin"))
    np.ascontiguousarray(events, dtype=N.EVENT_DTYPE).tofile(os.path.join(path, "event_trace.bin"))
    if dep_parent is not None:
        np.ascontiguousarray(dep_parent, dtype="<u2").tofile(os.path.join(path, "dep_parent.bin"))
    if mcs is not None:
        np.ascontiguousarray(mcs, dtype="<u8").tofile(os.path.join(path, "mcs.bin"))
    m = dict(meta, model=int(model), model_flags=int(model_flags), violation=int(violation), format="demi_b200/1")
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump(m, f, indent=1, sort_keys=True)


def load_experiment(path):
    out = {"externals": np.fromfile(os.path.join(path, "externals.bin"), dtype=N.EXT_DTYPE),
           "events": np.fromfile(os.path.join(path, "event_trace.bin"), dtype=N.EVENT_DTYPE)}
    for name, dt in (("dep_parent", "<u2"), ("mcs", "<u8")):
        p = os.path.join(path, name + ".bin")
        out[name] = np.fromfile(p, dtype=dt) if os.path.exists(p) else None
    with open(os.path.join(path, "meta.json")) as f:
        out["meta"] = json.load(f)
    return out
