"""MinimizationStats (minification/Minimizer.scala:30-237): the statistics the minimizers drive, with the
reference's shape and the JSON keys of minimization_stats.json (Serialization.scala:112-119;
src/main/python/minimization_stats/generate_graph.py reads them)."""
import json
import time

_STAT_KEYS = [
    "prune_duration_seconds", "prune_start_epoch", "prune_end_epoch",
    "replay_duration_seconds", "replay_end_epoch", "replay_start_epoch",
    "original_duration_seconds", "total_inputs", "total_events",
    "initial_verification_runs_needed", "minimized_deliveries", "minimized_externals", "minimized_timers",
]


class InnerStats(object):
    """MinimizationStats.InnerStats (Minimizer.scala:102-217): one <strategy, oracle> pair."""

    def __init__(self, name):
        self.name = name
        self.total_replays = 0            # schedules attempted, not counting the initial verification replay (:117-119)
        self.iterationSize = {}           # i-th replay attempt -> external events left (:104-105)
        self.internalIterationSize = {}   # i-th replay attempt -> internal events left (:107-108)
        self.maxDistance = {}             # new maxDistance -> iteration at which it was raised (:110-113)
        self.stats = {}
        # the sizes in recording order (the map above keeps only the last record per replay number; the engine
        # returns the whole series, which the parity tests compare)
        self.iteration_size_log = []
        self.internal_size_log = []
        self.reset()

    def reset(self):                      # :125-157 — as written, total_replays is NOT reset
        self.iterationSize.clear()
        self.internalIterationSize.clear()
        self.maxDistance.clear()
        self.iteration_size_log = []
        self.internal_size_log = []
        self.stats = {k: -1.0 for k in _STAT_KEYS[:7]}
        self.stats.update({k: 0.0 for k in _STAT_KEYS[7:]})

    def increment_replays(self, n=1):
        self.total_replays += n

    def record_replay_start(self):
        self.stats["replay_start_epoch"] = float(int(time.time() * 1000))

    def record_replay_end(self):
        self.stats["replay_end_epoch"] = float(int(time.time() * 1000))
        self.stats["replay_duration_seconds"] = (self.stats["replay_end_epoch"] - self.stats["replay_start_epoch"]) / 1000

    def record_prune_start(self):
        self.stats["prune_start_epoch"] = float(int(time.time() * 1000))

    def record_prune_end(self):
        self.stats["prune_end_epoch"] = float(int(time.time() * 1000))
        self.stats["prune_duration_seconds"] = (self.stats["prune_end_epoch"] - self.stats["prune_start_epoch"]) / 1000

    def record_iteration_size(self, iteration_size):          # :186-188
        self.iterationSize[self.total_replays] = int(iteration_size)
        self.iteration_size_log.append(int(iteration_size))

    def record_internal_size(self, iteration_size):           # :192-194
        self.internalIterationSize[self.total_replays] = int(iteration_size)
        self.internal_size_log.append(int(iteration_size))

    def recordDeliveryStats(self, minimized_deliveries, minimized_externals, minimized_timers):
        self.stats["minimized_deliveries"] = minimized_deliveries / 1.0
        self.stats["minimized_externals"] = minimized_externals / 1.0
        self.stats["minimized_timers"] = minimized_timers / 1.0

    def record_distance_increase(self, newDistance):          # :203-205
        self.maxDistance[int(newDistance)] = self.total_replays

    def toJson(self):                                          # :207-217
        d = {"name": self.name,
             "iteration_size": {str(k): v for k, v in self.iterationSize.items()},
             "internal_iteration_size": {str(k): v for k, v in self.internalIterationSize.items()},
             "total_replays": self.total_replays,
             "maxDistance": {str(k): v for k, v in self.maxDistance.items()}}
        d.update(self.stats)
        return d


class MinimizationStats(object):
    """class MinimizationStats (Minimizer.scala:30-100)."""

    def __init__(self):
        self.minimization_strategy = ""
        self.test_oracle = ""
        self.stats = []

    def updateStrategy(self, _minimization_strategy, _test_oracle):     # :41-47
        self.minimization_strategy = _minimization_strategy
        self.test_oracle = _test_oracle
        self.stats.append(InnerStats("(%s,%s)" % (_minimization_strategy, _test_oracle)))   # Tuple2.toString

    def inner(self):
        if not self.stats:                                     # the drivers always call updateStrategy first (RunnerUtils)
            self.updateStrategy(self.minimization_strategy, self.test_oracle)
        return self.stats[-1]

    def reset(self):
        self.inner().reset()

    def increment_replays(self, n=1):
        self.inner().increment_replays(n)

    def record_replay_start(self):
        self.inner().record_replay_start()

    def record_replay_end(self):
        self.inner().record_replay_end()

    def record_prune_start(self):
        self.inner().record_prune_start()

    def record_prune_end(self):
        self.inner().record_prune_end()

    def record_iteration_size(self, iteration_size):
        self.inner().record_iteration_size(iteration_size)

    def record_internal_size(self, iteration_size):
        self.inner().record_internal_size(iteration_size)

    def record_distance_increase(self, newDistance):
        self.inner().record_distance_increase(newDistance)

    def recordDeliveryStats(self, minimized_deliveries, minimized_externals, minimized_timers):
        self.inner().recordDeliveryStats(minimized_deliveries, minimized_externals, minimized_timers)

    def toJson(self):                                          # :98-100: a JSON array of the inner objects
        return json.dumps([s.toJson() for s in self.stats])

    @staticmethod
    def fromJson(text):                                        # :219-236
        outer = MinimizationStats()
        for inner in json.loads(text):
            obj = InnerStats(inner["name"])
            obj.total_replays = int(inner["total_replays"])
            for k in list(obj.stats.keys()):
                obj.stats[k] = float(inner[k])
            obj.maxDistance = {int(k): int(v) for k, v in inner["maxDistance"].items()}
            obj.iterationSize = {int(k): int(v) for k, v in inner["iteration_size"].items()}
            obj.internalIterationSize = {int(k): int(v) for k, v in inner["internal_iteration_size"].items()}
            outer.stats.append(obj)
        return outer

    # ---- what the engine-driven minimizers fill in: the counters of the sequential algorithm
    @property
    def total_replays(self):
        return self.inner().total_replays

    @total_replays.setter
    def total_replays(self, v):
        self.inner().total_replays = int(v)

    @property
    def iteration_size(self):
        return self.inner().iteration_size_log

    @property
    def internal_sizes(self):
        return self.inner().internal_size_log

    def record_series(self, sizes, internal=False):
        """The engine returns the whole record_*_size series of a minimization in one array; replay it as the calls
        the sequential algorithm makes: one record per test, keyed by the replay number, plus the fencepost record
        (DeltaDebugging.scala:60; OneAtATimeRemoval) which lands on the last replay number again."""
        inner = self.inner()
        base = inner.total_replays
        for i, s in enumerate(sizes):
            key = base + min(i + 1, len(sizes) - 1 if len(sizes) > 1 else 1)
            (inner.internalIterationSize if internal else inner.iterationSize)[key] = int(s)
            (inner.internal_size_log if internal else inner.iteration_size_log).append(int(s))
