"""Known-answer checks for the oracle's ProvenanceTracker restatement (schedulers/Util.scala:267-376).

The expected masks are derived by hand from the reference's definition: happens-before = reflexive-transitive
closure of {earlier delivery on the same receiver, delivery -> message it created}; position t stays iff it
happens before the last delivery on some affected node and that delivery does not happen before it."""
import numpy as np

from oracle import binding as O

EV_MSG_EVENT = 2


def deliveries(seq):
    ev = np.zeros(len(seq), dtype=O.EVENT_DTYPE)
    for i, (node, rcv) in enumerate(seq):
        ev[i]["kind"] = EV_MSG_EVENT
        ev[i]["dst"] = rcv
        ev[i]["node"] = node
    return ev


# root 0; n1 -> actor0, n2 -> actor1 (externals, parent root); n3 created by n1 -> actor1; n4 created by n2 -> actor0;
# n5 created by n3 -> actor0
PARENT = [0, 0, 0, 1, 2, 3]
TRACE = deliveries([(1, 0), (2, 1), (3, 1), (4, 0), (5, 0)])


def test_provenance_hand_derived():
    keep, out = O.provenance(TRACE, PARENT, 0b01, 1)      # last delivery on actor 0 = n5; everything else precedes it
    assert out["status"] == 0 and out["n_trace"] == 6
    assert int(keep[0]) == 0b011111 and out["n_kept"] == 5
    keep, out = O.provenance(TRACE, PARENT, 0b10, 1)      # last on actor 1 = n3: past = root, n1 (parent), n2 (same receiver)
    assert int(keep[0]) == 0b000111 and out["n_kept"] == 3
    keep, out = O.provenance(TRACE, PARENT, 0b11, 1)      # n3 precedes n5, so n3 stays; n5 precedes nothing
    assert int(keep[0]) == 0b011111
    keep, out = O.provenance(TRACE, PARENT, 0, 1)         # forall over no last events is true: everything is pruned
    assert int(keep[0]) == 0 and out["n_kept"] == 0
    keep, out = O.provenance(TRACE, PARENT, 0b100, 1)     # affected node never received anything
    assert int(keep[0]) == 0


def test_provenance_last_event_itself_is_pruned():
    # a single delivery: (u,u) is in happensBefore (Util.scala:295-297), so u is "after" itself and is dropped,
    # while the root stays
    keep, out = O.provenance(deliveries([(1, 0)]), [0, 0], 0b1, 1)
    assert int(keep[0]) == 0b01 and out["n_trace"] == 2


def test_provenance_repeated_unique():
    # the same Unique delivered twice in a row on its receiver: no cycle
    keep, out = O.provenance(deliveries([(1, 0), (1, 0), (2, 0)]), [0, 0, 0], 0b1, 1)
    assert out["status"] == 0
    assert int(keep[0]) == 0b0111       # root and both deliveries of n1 precede n2
    # ... with another delivery on that receiver in between: n1 -> n2 -> n1 is a cycle (Util.scala:506 sys.error)
    keep, out = O.provenance(deliveries([(1, 0), (2, 0), (1, 0)]), [0, 0, 0], 0b1, 1)
    assert out["status"] == 1 and int(keep[0]) == 0


def test_provenance_overflow():
    keep, out = O.provenance(deliveries([(1, 0)] * 70), [0, 0], 0b1, 1)
    assert out["status"] == 2


def test_fuzz_provenance_affected_nodes():
    from demi_b200 import events as E
    from demi_b200 import _native as N
    ext = E.pack_externals(E.raft5_program())
    res = O.fuzz_batch(N.MODEL_RAFT5, ext, 1, 4000, 50, 5, model_flags=1)
    viol = np.nonzero(res["violation"])[0]
    assert len(viol) > 0
    for i in viol[:20]:
        keep, out = O.fuzz_provenance(N.MODEL_RAFT5, ext, 1 + int(i), 50, 5, 1, model_flags=1)
        assert out["status"] == 0 and out["violation"] == res["violation"][i]
        assert bin(int(out["affected_mask"])).count("1") == 2          # both invariants are about a pair of nodes
        assert out["n_trace"] == res["steps"][i] + 1
        assert 0 < out["n_kept"] < out["n_trace"]
        assert int(keep[0]) & 1                                          # the root precedes everything
