// ORACLE — test infrastructure.  Akka actors that implement the models of DESIGN.md §3 exactly as oracle/models.c
// specifies them, so that the unmodified reference can be run over them.  Messages are (type, p0, p1) triples whose
// toString is stable: BasicFingerprint (MessageFingerprints.scala:42-51) fingerprints by toString.
package demi_oracle

import akka.actor.{ Actor, ActorRef, Cancellable, Props }
import scala.collection.mutable
import scala.concurrent.duration._

/** One message of a model: type + two payload words (demi_msg without the addressing). */
final case class M(t: Int, p0: Long, p1: Long) { override def toString = "M(" + t + "," + p0 + "," + p1 + ")" }

/** Actor states, published after every receive() so that the invariant can read them without checkpoint messages
  * (the engine's invariant reads model state directly; DESIGN.md §8). */
object StateRegistry {
  val words = new mutable.HashMap[Int, Array[Long]]
  def reset() = words.synchronized { words.clear() }
  def publish(actor: Int, w: Array[Long]) = words.synchronized { words(actor) = w.clone() }
}

abstract class ModelActor(val self_idx: Int, n: Int) extends Actor {
  def peer(i: Int): ActorRef = context.actorFor("../" + i)
  def send(dst: Int, t: Int, p0: Long = 0, p1: Long = 0) = peer(dst) ! M(t, p0, p1)
  def sender_idx: Int = { val nm = sender().path.name; if (nm.forall(_.isDigit)) nm.toInt else 0xFF }   // deadLetters / Timer
}

// ---- pingpong3: Ping(k) to X => X counts it and sends Pong(k) to (X+1)%3; Pong => X counts it
class PingPong(idx: Int) extends ModelActor(idx, 3) {
  var pings = 0L; var pongs = 0L
  def receive = {
    case M(1, k, _) => pings += 1; send((idx + 1) % 3, 2, k); StateRegistry.publish(idx, Array(pings, pongs))
    case M(2, _, _) => pongs += 1; StateRegistry.publish(idx, Array(pings, pongs))
    case _ =>
  }
}

// ---- bcast32: Flood(ttl) / Inject(ttl) => count++, maxTtlSeen, rebroadcast Flood(ttl-1) to the 31 peers while ttl > 0
class Bcast(idx: Int) extends ModelActor(idx, 32) {
  var count = 0L; var maxTtl = 0L
  def receive = {
    case M(t, ttl, _) if t == 1 || t == 2 =>
      count += 1; if (ttl + 1 > maxTtl) maxTtl = ttl + 1
      if (ttl > 0) for (j <- 0 until 32 if j != idx) send(j, 1, ttl - 1)
      StateRegistry.publish(idx, Array(count, maxTtl))
    case _ =>
  }
}

// ---- raft5: Raft Fig. 2, 5 nodes, tick-driven timers, log capacity 8, one entry per AppendEntries (oracle/models.c)
object Raft {
  val BOOT = 1; val CLIENT_CMD = 2; val ELECTION_TICK = 3; val REQUEST_VOTE = 4; val VOTE_REPLY = 5
  val HEARTBEAT_TICK = 6; val APPEND_ENTRIES = 7; val APPEND_REPLY = 8
  val INIT = 0; val FOLLOWER = 1; val CANDIDATE = 2; val LEADER = 3
  val LOG_CAP = 8; val NONE = 0xFF
}
class Raft(idx: Int, flags: Int) extends ModelActor(idx, 5) {
  import Raft._
  import context.dispatcher
  var role = INIT; var term = 0; var voted = NONE; var votes = 0; var heard = 0; var commit = 0
  val logTerm = new Array[Int](LOG_CAP); val logVal = new Array[Int](LOG_CAP); var logLen = 0
  val next = new Array[Int](5); val matchIdx = new Array[Int](5)
  var heartbeat: Cancellable = null
  // tick timers: WeaveActor.aj:264-279 turns `schedule` into a repeating timer the scheduler owns
  def scheduleRepeating(t: Int) = context.system.scheduler.schedule(1.second, 1.second, self, M(t, 0, 0))
  def stepDown(t: Int) = {
    if (role == LEADER && heartbeat != null) { heartbeat.cancel(); heartbeat = null }
    if (t > term) { term = t; voted = NONE }
    role = FOLLOWER; votes = 0
  }
  def sendAppend(j: Int) = {
    val prev = next(j); val pt = if (prev > 0) logTerm(prev - 1) else 0
    val has = if (prev < logLen) 1 else 0
    val et = if (has == 1) logTerm(prev) else 0; val ev = if (has == 1) logVal(prev) else 0
    send(j, APPEND_ENTRIES, term | (prev << 8) | (pt << 16) | (commit.toLong << 24), has | (et << 8) | (ev << 16))
  }
  def publish() = StateRegistry.publish(idx, Array[Long](role, term, voted, votes, logLen, commit, heard) ++
    logTerm.map(_.toLong) ++ logVal.map(_.toLong) ++ next.map(_.toLong) ++ matchIdx.map(_.toLong))
  def receive = { case M(t, p0, p1) => handle(t, p0.toInt, p1.toInt, sender_idx); publish() case _ => }
  def handle(ty: Int, p0: Int, p1: Int, src: Int): Unit = {
    val lastIdx = logLen; val lastTerm = if (lastIdx > 0) logTerm(lastIdx - 1) else 0
    val t = p0 & 0xFF
    if (ty != BOOT && ty != CLIENT_CMD && role == INIT) return
    ty match {
      case BOOT => if (role == INIT) { role = FOLLOWER; scheduleRepeating(ELECTION_TICK) }
      case CLIENT_CMD => if (role == LEADER && logLen < LOG_CAP) { logTerm(logLen) = term; logVal(logLen) = p0 & 0x7F; logLen += 1 }
      case ELECTION_TICK =>
        if (role == LEADER) return
        if (heard != 0) { heard = 0; return }
        if (term == 255) return
        term += 1; role = CANDIDATE; voted = idx; votes = 1 << idx
        for (j <- 0 until 5 if j != idx) send(j, REQUEST_VOTE, term | (lastIdx << 8) | (lastTerm << 16))
      case REQUEST_VOTE =>
        val li = (p0 >> 8) & 0xFF; val lt = (p0 >> 16) & 0xFF
        if (t > term) stepDown(t)
        val upToDate = lt > lastTerm || (lt == lastTerm && li >= lastIdx)
        val canVote = voted == NONE || voted == src || (flags & 1) != 0                 // bit0: the seeded double-vote bug
        val grant = if (t == term && canVote && upToDate) 1 else 0
        if (grant == 1) { voted = src; heard = 1 }
        send(src, VOTE_REPLY, term | (grant << 8))
      case VOTE_REPLY =>
        val g = (p0 >> 8) & 1
        if (t > term) { stepDown(t); return }
        if (role == CANDIDATE && t == term && g == 1) {
          votes |= 1 << src
          if (Integer.bitCount(votes) >= 3) {
            role = LEADER
            for (j <- 0 until 5) { next(j) = logLen; matchIdx(j) = 0 }
            if (logLen < LOG_CAP) { logTerm(logLen) = term; logVal(logLen) = 0x80 | idx; logLen += 1 }   // leader no-op
            for (j <- 0 until 5 if j != idx) sendAppend(j)
            heartbeat = scheduleRepeating(HEARTBEAT_TICK)
          }
        }
      case HEARTBEAT_TICK => if (role == LEADER) for (j <- 0 until 5 if j != idx) sendAppend(j)
      case APPEND_ENTRIES =>
        val prev = (p0 >> 8) & 0xFF; val pt = (p0 >> 16) & 0xFF; val lc = (p0 >> 24) & 0xFF
        val has = p1 & 1; val et = (p1 >> 8) & 0xFF; val ev = (p1 >> 16) & 0xFF
        if (t < term) { send(src, APPEND_REPLY, term); return }
        if (t > term || role != FOLLOWER) stepDown(t)
        heard = 1
        val ok = prev <= logLen && (prev == 0 || logTerm(prev - 1) == pt)
        if (!ok) { send(src, APPEND_REPLY, term); return }
        var mi = prev
        if (has == 1) {
          if (logLen > prev && logTerm(prev) != et) { for (k <- prev until LOG_CAP) { logTerm(k) = 0; logVal(k) = 0 }; logLen = prev }
          if (logLen == prev && prev < LOG_CAP) { logTerm(prev) = et; logVal(prev) = ev; logLen = prev + 1 }
          if (logLen > prev) mi = prev + 1
        }
        val nc = math.min(lc, mi); if (nc > commit) commit = nc
        send(src, APPEND_REPLY, term | (1 << 8) | (mi << 16))
      case APPEND_REPLY =>
        val ok = (p0 >> 8) & 1; val mi = (p0 >> 16) & 0xFF
        if (t > term) { stepDown(t); return }
        if (role != LEADER || t != term) return
        if (ok == 1) {
          if (mi > matchIdx(src)) matchIdx(src) = mi
          if (mi > next(src)) next(src) = mi
          var i = logLen; var done = false
          while (i > commit && !done) {
            if (logTerm(i - 1) == term || (flags & 2) != 0) {                           // bit1: the seeded stale-term commit bug
              val cnt = 1 + (0 until 5).count(k => k != idx && matchIdx(k) >= i)
              if (cnt >= 3) { commit = i; done = true }
            }
            i -= 1
          }
        } else if (next(src) > 0) next(src) -= 1
      case _ =>
    }
  }
}

object ModelProps {
  def of(model: String, idx: Int, flags: Int): Props = model match {
    case "pingpong3" => Props(classOf[PingPong], idx)
    case "raft5" => Props(classOf[Raft], idx, flags)
    case "bcast32" => Props(classOf[Bcast], idx)
  }
  def actors(model: String) = model match { case "pingpong3" => 3 case "raft5" => 5 case "bcast32" => 32 }
}
