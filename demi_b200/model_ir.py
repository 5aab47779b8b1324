"""Assembler for the model IR (include/demi_model_ir.h): how a host describes ITS application to demi_load_model —
receive() and invariant as register programs, initial states, the external-message filter and the name table."""
import struct

import numpy as np

OPS = {"HALT": 0, "LDI": 1, "MOV": 2, "ADD": 3, "SUB": 4, "MUL": 5, "AND": 6, "OR": 7, "XOR": 8, "SHL": 9, "SHR": 10, "MOD": 11,
       "LDW": 12, "STW": 13, "LDA": 14, "JMP": 15, "JEQ": 16, "JNE": 17, "JLT": 18, "JGE": 19, "SEND": 20,
       "SCHED_ONCE": 21, "SCHED_REPEAT": 22, "CANCEL": 23, "RET": 24}
MAGIC, VERSION, HEADER_WORDS = 0x52494D44, 1, 16
MODEL_IR = 100


def _reg(x):
    assert isinstance(x, str) and x[0] == "r" and 0 <= int(x[1:]) < 16, x
    return int(x[1:])


def assemble(lines):
    """[("LDI", "r6", 1), ("label", "again"), ("JNE", "r2", "r6", "again"), ...] -> list of u32 words.
    Immediates of jumps are labels (resolved to word offsets) or integers."""
    words, labels, fixups = [], {}, []
    for ln in lines:
        op = ln[0]
        if op == "label":
            labels[ln[1]] = len(words)
            continue
        code = OPS[op]
        if op == "LDI":
            words += [code | (_reg(ln[1]) << 8), int(ln[2]) & 0xFFFFFFFF]
        elif op == "JMP":
            fixups.append((len(words) + 1, ln[1])); words += [code, 0]
        elif op in ("JEQ", "JNE", "JLT", "JGE"):
            fixups.append((len(words) + 1, ln[3])); words += [code | (_reg(ln[1]) << 8) | (_reg(ln[2]) << 16), 0]
        elif op in ("HALT",):
            words.append(code)
        elif op in ("MOV", "LDW", "STW"):
            words.append(code | (_reg(ln[1]) << 8) | (_reg(ln[2]) << 16))
        elif op == "RET":
            words.append(code | (_reg(ln[1]) << 8))
        else:                                              # three-register forms
            words.append(code | (_reg(ln[1]) << 8) | (_reg(ln[2]) << 16) | (_reg(ln[3]) << 24))
    for pos, target in fixups:
        words[pos] = labels[target] if isinstance(target, str) else int(target)
    return words


def build_blob(actor_names, type_names, state_words, receive, invariant, init=None, external_types=(), fanout=1):
    """type_names[t] names message type t (index 0 is unused by convention).  `init`: n_actors x state_words."""
    na, nt = len(actor_names), len(type_names)
    recv, inv = assemble(receive), assemble(invariant)
    init = np.zeros((na, state_words), dtype=np.uint32) if init is None else np.asarray(init, dtype=np.uint32).reshape(na, state_words)
    names = b"".join(n.encode() + b"\0" for n in list(actor_names) + list(type_names))
    names += b"\0" * (-len(names) % 4)
    ext_mask = 0
    for t in external_types:
        ext_mask |= 1 << t
    hdr = [MAGIC, VERSION, na, state_words, nt, len(recv), len(inv), ext_mask, fanout, len(names)] + [0] * 6
    return struct.pack("<%dI" % (HEADER_WORDS + len(recv) + len(inv)), *(hdr + recv + inv)) + init.astype("<u4").tobytes() + names


def pingpong3_blob():
    """The pingpong3 model of DESIGN.md §3 in the IR: Ping(k) to X => X counts it and sends Pong(k) to (X+1)%3; Pong => X
    counts it.  State: w0 = pings, w1 = pongs.  Invariant (test hook, as in the compiled model): flags bit0 set and actor 0
    has received >= flags>>8 pongs -> code 7, affected = {actor 0}."""
    PING, PONG = 1, 2
    receive = [
        ("LDI", "r6", PING), ("LDI", "r7", PONG), ("LDI", "r8", 0), ("LDI", "r9", 1), ("LDI", "r10", 3),
        ("JNE", "r2", "r6", "not_ping"),
        ("LDW", "r11", "r8"), ("ADD", "r11", "r11", "r9"), ("STW", "r8", "r11"),          # pings++
        ("ADD", "r12", "r0", "r9"), ("MOD", "r12", "r12", "r10"),                          # (self + 1) % 3
        ("MOV", "r13", "r3"), ("LDI", "r14", 0),                                           # p0 = k, p1 = 0
        ("SEND", "r12", "r7", "r13"),
        ("HALT",),
        ("label", "not_ping"),
        ("JNE", "r2", "r7", "done"),
        ("LDW", "r11", "r9"), ("ADD", "r11", "r11", "r9"), ("STW", "r9", "r11"),          # pongs++
        ("label", "done"),
        ("HALT",),
    ]
    invariant = [
        ("LDI", "r6", 1), ("LDI", "r7", 8), ("LDI", "r8", 0),
        ("AND", "r9", "r5", "r6"), ("JEQ", "r9", "r8", "ok"),                              # flags & 1
        ("SHR", "r10", "r5", "r7"),                                                         # threshold = flags >> 8
        ("LDA", "r11", "r8", "r6"),                                                         # actor 0, word 1 = pongs
        ("JLT", "r11", "r10", "ok"),
        ("LDI", "r0", 7), ("LDI", "r1", 1), ("RET", "r0"),
        ("label", "ok"),
        ("LDI", "r0", 0), ("LDI", "r1", 0), ("RET", "r0"),
    ]
    return build_blob(["A", "B", "C"], ["-", "Ping", "Pong"], 2, receive, invariant, external_types=(PING,), fanout=1)
