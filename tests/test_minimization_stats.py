"""MinimizationStats (minification/Minimizer.scala:30-237) keeps the reference's shape and minimization_stats.json keys;
conjoinAtoms (minification/Util.scala:167-178) makes two externals one atomic event for DDMin."""
import json

import numpy as np

import demi_b200 as D

REFERENCE_KEYS = {   # InnerStats.toJson (:207-217) + the stats map (:129-156)
    "name", "iteration_size", "internal_iteration_size", "total_replays", "maxDistance",
    "prune_duration_seconds", "prune_start_epoch", "prune_end_epoch", "replay_duration_seconds", "replay_end_epoch",
    "replay_start_epoch", "original_duration_seconds", "total_inputs", "total_events",
    "initial_verification_runs_needed", "minimized_deliveries", "minimized_externals", "minimized_timers"}


def test_json_shape_and_round_trip():
    s = D.MinimizationStats()
    s.updateStrategy("DDMin", "STSSchedNoPeek")
    s.record_prune_start()
    for size in (8, 4, 4, 2):
        s.increment_replays()
        s.record_iteration_size(size)
    s.record_iteration_size(2)                       # the fencepost record of DDMin.minimize lands on the last replay number
    s.record_prune_end()
    s.updateStrategy("LeftToRightOneAtATime", "STSSched")
    s.increment_replays()
    s.record_internal_size(17)
    s.record_distance_increase(2)
    s.recordDeliveryStats(40, 5, 3)
    arr = json.loads(s.toJson())
    assert [a["name"] for a in arr] == ["(DDMin,STSSchedNoPeek)", "(LeftToRightOneAtATime,STSSched)"]   # Tuple2.toString
    for a in arr:
        assert set(a) == REFERENCE_KEYS
    assert arr[0]["iteration_size"] == {"1": 8, "2": 4, "3": 4, "4": 2} and arr[0]["total_replays"] == 4
    assert arr[0]["prune_duration_seconds"] >= 0 and arr[0]["replay_duration_seconds"] == -1.0
    assert arr[1]["internal_iteration_size"] == {"1": 17} and arr[1]["maxDistance"] == {"2": 1}
    assert arr[1]["minimized_deliveries"] == 40.0 and arr[1]["minimized_timers"] == 3.0
    back = D.MinimizationStats.fromJson(s.toJson())
    assert json.loads(back.toJson()) == arr
    s.reset()                                        # as written: the maps and stats are cleared, total_replays is not
    assert s.inner().iterationSize == {} and s.inner().total_replays == 1


def test_record_series_is_the_sequential_call_sequence():
    s = D.MinimizationStats()
    s.updateStrategy("DDMin", "STSSchedNoPeek")
    s.record_series([8, 4, 4, 2, 2])                 # four tests + the fencepost, as demi_ddmin returns them
    s.total_replays = 4
    assert s.inner().iterationSize == {1: 8, 2: 4, 3: 4, 4: 2} and s.iteration_size == [8, 4, 4, 2, 2]


def test_conjoined_atoms_in_the_oracle_split(oracle):
    """Externals 2 and 7 conjoined: they form one atom whose first element is the lower index, so split_list keeps
    them on the same side and a superset oracle that needs only one of them still gets both back."""
    prog = [D.Start(a) for a in range(3)] + [D.Send(k % 3, 1, k) for k in range(8)]
    ext = D.pack_externals(prog)
    K = np.zeros(1, dtype=np.uint64)
    K[0] = np.uint64(1 << 5)                          # the violation needs external 5 only
    rc, mcs, total, iters, log = oracle.ddmin_superset(ext, K)
    assert rc == 0 and int(mcs[0]) == 1 << 5
    oracle.set_conjoined([(5, 9)], len(ext))
    try:
        rc, mcs2, total2, iters2, log2 = oracle.ddmin_superset(ext, K)
    finally:
        oracle.set_conjoined([], len(ext))
    assert rc == 0 and int(mcs2[0]) == (1 << 5) | (1 << 9)            # kept or removed together
    for m in log2:
        assert ((int(m[0]) >> 5) & 1) == ((int(m[0]) >> 9) & 1)       # no test ever separates them
