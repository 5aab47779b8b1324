// ORACLE — test infrastructure.  Drives the UNMODIFIED reference over the model actors and dumps what it records in
// the flat experiment layout of include/demi_b200.h (externals.bin, event_trace.bin, dep_parent.bin, mcs.bin,
// meta.json), with ids normalised per execution (see README.md).
package demi_oracle

import akka.dispatch.verification._
import java.io.{ File, FileOutputStream }
import java.nio.{ ByteBuffer, ByteOrder }
import scala.collection.mutable

object Runner {
  // the canonical external programs of demi_b200/events.py
  def program(model: String, flags: Int, clientCmds: Int): Seq[ExternalEvent] = {
    val n = ModelProps.actors(model)
    val starts = (0 until n).map(i => Start(() => ModelProps.of(model, i, flags), i.toString))
    val sends: Seq[ExternalEvent] = model match {
      case "pingpong3" => (0 until 100).map(k => Send((k % 3).toString, BasicMessageConstructor(M(1, k, 0))))
      case "raft5" => (0 until 5).map(i => Send(i.toString, BasicMessageConstructor(M(1, 0x1F, 0)))) ++
        (0 until clientCmds).map(i => Send((i % 5).toString, BasicMessageConstructor(M(2, 1 + i, 0))))
      case "bcast32" => Seq(Send("0", BasicMessageConstructor(M(2, 3, 0))))
    }
    starts ++ sends :+ WaitQuiescence()
  }

  // model invariants over the published states, as oracle/models.c computes them
  def invariant(model: String, flags: Int): TestOracle.Invariant = (_, _) => {
    val st = StateRegistry.words.synchronized { StateRegistry.words.toMap }
    val code = model match {
      case "raft5" =>
        val ids = st.keys.toSeq.sorted
        var c = 0
        for (i <- ids; j <- ids if i < j) {
          val a = st(i); val b = st(j)
          if (a(0) == 3 && b(0) == 3 && a(1) == b(1)) c = if (c == 0 || c == 2) 1 else c
          if (c == 0) { val m = math.min(a(5), b(5)).toInt; if ((0 until m).exists(k => a(7 + k) != b(7 + k) || a(15 + k) != b(15 + k))) c = 2 }
        }
        c
      case "pingpong3" => if ((flags & 1) != 0 && st.get(0).exists(_(1) >= (flags >> 8))) 7 else 0
      case "bcast32" => if (flags != 0 && st.values.exists(_(0) >= flags)) 3 else 0
    }
    if (code == 0) None else Some(CodeFingerprint(code))
  }
  case class CodeFingerprint(code: Int) extends ViolationFingerprint {
    def matches(other: ViolationFingerprint) = other == this
    def affectedNodes() = Seq.empty
  }

  def idx(name: String): Int = if (name.nonEmpty && name.forall(_.isDigit)) name.toInt else if (name == "Timer") 0xFE else 0xFF
  def payload(msg: Any): (Int, Long, Long) = msg match { case M(t, a, b) => (t, a, b) case _ => (0, 0, 0) }

  /** EventTrace -> demi_event records (16 bytes), ids renumbered per execution. */
  def dump(dir: String, model: String, flags: Int, externals: Seq[ExternalEvent], trace: EventTrace, code: Int,
           depParent: Map[Int, Int], mcs: Option[Seq[ExternalEvent]]) {
    new File(dir).mkdirs()
    val uniq = new mutable.HashMap[Int, Int]; val node = new mutable.LinkedHashMap[Int, Int]
    def u(id: Int) = uniq.getOrElseUpdate(id, uniq.size + 1)
    def rec(kind: Int, src: Int, dst: Int, t: Int, p0: Long, p1: Long, un: Int, nd: Int) = {
      val b = ByteBuffer.allocate(16).order(ByteOrder.LITTLE_ENDIAN)
      b.put(kind.toByte).put(src.toByte).put(dst.toByte).put(t.toByte).putInt(p0.toInt).putInt(p1.toInt).putShort(un.toShort).putShort(nd.toShort)
      b.array()
    }
    val out = new FileOutputStream(dir + "/event_trace.bin")
    for (e <- trace.events) e match {
      case UniqueMsgSend(MsgSend(s, r, m), id) => val (t, a, b) = payload(m); out.write(rec(1, idx(s), idx(r), t, a, b, u(id), 0))
      case UniqueMsgEvent(MsgEvent(s, r, m), id) => val (t, a, b) = payload(m); out.write(rec(2, idx(s), idx(r), t, a, b, u(id), 0))
      case SpawnEvent(_, _, name, _) => out.write(rec(3, 0xFF, idx(name), 0, 0, 0, 0, 0))
      case KillEvent(name) => out.write(rec(4, 0xFF, idx(name), 0, 0, 0, 0, 0))
      case PartitionEvent((a, b)) => out.write(rec(5, idx(a), idx(b), 0, 0, 0, 0, 0))
      case UnPartitionEvent((a, b)) => out.write(rec(6, idx(a), idx(b), 0, 0, 0, 0, 0))
      case BeginWaitQuiescence => out.write(rec(7, 0xFF, 0xFF, 0, 0, 0, 0, 0))
      case Quiescence => out.write(rec(8, 0xFF, 0xFF, 0, 0, 0, 0, 0))
      case _ =>
    }
    out.close()
    val xo = new FileOutputStream(dir + "/externals.bin")
    for ((e, i) <- externals.zipWithIndex) {
      val b = ByteBuffer.allocate(16).order(ByteOrder.LITTLE_ENDIAN)
      e match {
        case Start(_, name) => b.put(1.toByte).put(idx(name).toByte).put(0.toByte).put(0.toByte).putInt(0).putInt(0)
        case Kill(name) => b.put(2.toByte).put(idx(name).toByte).put(0.toByte).put(0.toByte).putInt(0).putInt(0)
        case Send(name, ctor) => val (t, a, c) = payload(ctor()); b.put(3.toByte).put(idx(name).toByte).put(0.toByte).put(t.toByte).putInt(a.toInt).putInt(c.toInt)
        case WaitQuiescence() => b.put(4.toByte).put(0.toByte).put(0.toByte).put(0.toByte).putInt(0).putInt(0)
        case Partition(x, y) => b.put(5.toByte).put(idx(x).toByte).put(idx(y).toByte).put(0.toByte).putInt(0).putInt(0)
        case UnPartition(x, y) => b.put(6.toByte).put(idx(x).toByte).put(idx(y).toByte).put(0.toByte).putInt(0).putInt(0)
        case _ => b.put(0.toByte).put(0.toByte).put(0.toByte).put(0.toByte).putInt(0).putInt(0)
      }
      b.putInt(i + 1); xo.write(b.array())
    }
    xo.close()
    mcs.foreach { m =>
      val words = (externals.size + 63) / 64; val bits = new Array[Long](words)
      for ((e, i) <- externals.zipWithIndex if m.contains(e)) bits(i / 64) |= 1L << (i % 64)
      val b = ByteBuffer.allocate(8 * words).order(ByteOrder.LITTLE_ENDIAN); bits.foreach(b.putLong)
      val mo = new FileOutputStream(dir + "/mcs.bin"); mo.write(b.array()); mo.close()
    }
    val meta = new FileOutputStream(dir + "/meta.json")
    meta.write(("{\n \"format\": \"demi_b200/1\",\n \"model\": " + Map("pingpong3" -> 1, "raft5" -> 2, "bcast32" -> 3)(model) +
      ",\n \"model_flags\": " + flags + ",\n \"violation\": " + code + ",\n \"source\": \"NetSys/demi on the JVM\"\n}\n").getBytes)
    meta.close()
  }

  def main(args: Array[String]) {
    EventTypes.setExternalMessageFilter { case M(t, _, _) => t == 1 || t == 2 case _ => false }   // Instrumenter.scala:1115-1117
    args(0) match {
      case "fuzz" =>
        val model = args(1); val seed = args(2).toLong; val maxMessages = args(3).toInt; val interval = args(4).toInt; val dir = args(5)
        val flags = if (args.length > 6) args(6).toInt else 1
        val prog = program(model, flags, 0)
        StateRegistry.reset()
        val sched = new RandomScheduler(SchedulerConfig(messageFingerprinter = new FingerprintFactory), 1, interval,
          randomizationStrategy = new FullyRandom(seed = seed))
        sched.setInvariant(invariant(model, flags))
        if (maxMessages >= 0) sched.setMaxMessages(maxMessages)
        Instrumenter().scheduler = sched
        val found = sched.explore(prog)
        val code = found.map(_._2.asInstanceOf[CodeFingerprint].code).getOrElse(0)
        val trace = found.map(_._1).getOrElse(sched.event_orchestrator.events)
        dump(dir, model, flags, prog, trace, code, Map.empty, None)
        sched.shutdown()
      case "ddmin" =>
        sys.error("see README.md: replay the dumped externals through RunnerUtils.stsSchedDDMin and dump mcs.bin")
    }
    System.exit(0)
  }
}
