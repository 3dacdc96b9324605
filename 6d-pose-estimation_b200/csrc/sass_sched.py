#!/usr/bin/env python3
"""sass_sched.py -- post-link scheduling pass for the ADD-S scan loop (build step of libp6d.so).

Why: the scan's tile is 3 FADD2 + FMUL2 + 2 FFMA2 (packed FP32, two issue cycles each) plus one
FMNMX3 per two point pairs.  Where the FMNMX3 sits between the packed ops decides whether it
overlaps with them; ptxas parks the software-pipelined minima at the top of the loop body and
offers no control over placement (inline-asm order is not preserved).  Measured on B200
(NOTES.md, tools/exp_block.cu): same source, minima behind every third FADD2, packed ops encoded
with stall 1, yield hint every 8th instruction: +5 % over the best ptxas schedule.

What: permutes the 128-bit instruction words of ONE loop body in place and rewrites only the
stall / yield / reuse bits of their scheduling control fields.  No instruction is added,
removed or re-encoded; branches and every instruction the parser does not fully understand stay
where they are.

Safety rules (checked; the patch is refused when one fails):
  * only FMNMX / FMNMX3 without scoreboard traffic (no wait mask, no barrier) are moved;
  * every pair of instructions with a register conflict (R, UR, P, UP; at least one write) keeps its
    original relative order, so data flow -- including loop-carried values -- is unchanged;
  * fixed-latency results (read after write): the consumer is issued at least LAT cycles after the producer, counted
    from encoded stall counts plus the two cycles a packed op occupies the FMA pipe (in-order
    issue makes real gaps >= modelled gaps).  LAT = 4 packed->packed and 5 packed->minimum (the
    smallest gaps ptxas itself uses in this loop), 6 minimum->minimum; every other dependent pair
    keeps at least the gap it had in the ptxas schedule (up to 16 cycles, which covers the longest
    fixed latency, predicate -> branch);
  * variable-latency producers (LDS) are covered by scoreboard fields, which travel with their
    instructions (those instructions are never moved);
  * `.reuse` flags survive only where the following instruction is the original successor;
  * the patched image is written to a temporary file, disassembled again and compared with the plan;
    only then does it replace the output (a failed run leaves the library as it was);
  * a loop that no longer carries ptxas' schedule (already re-laid) is refused.
Encoding packed ops with stall 1 relies on the FMA pipe's own interlock (a second packed op waits
for the pipe: ncu reports it as the math-pipe-throttle stall); the model above therefore counts two
pipe cycles per packed op, never the encoded stall alone.  Evidence beyond the model: ~4,200
re-laid candidates (tools/sched_search.py) each reproduced the output hash of the ptxas-scheduled
kernel over 65,536-524,288 poses, and the GPU parity tests (bit-exact against the oracle) run on
the patched library.

Usage: sass_sched.py FILE KERNEL-SUBSTRING POLICY [OUT | --out=OUT] [--loop=uniform|0xADDR]
                     [--packed-stall=1] [--yield=periodP,PHASE|mask0110..|0|1] [--order=i,j,..] [--mark=CLASS] [--show]
       sass_sched.py FILE --plan=PLAN.json [--out=OUT] [--loop=uniform] [--mark=CLASS]     (order + yield mask from
                                                                                   tools/sched_search.py)
POLICY: identity | cluster_end | spaced=FADD2:2[,FMUL2:1][/next=FADD2+FMUL2]
"""
import os
import re
import struct
import subprocess
import sys

PACKED = ("FADD2", "FMUL2", "FFMA2")
MINS = ("FMNMX", "FMNMX3")
LAT_FMA_FMA, LAT_FMA_MIN, LAT_MIN_MIN = 4, 5, 6
BRANCH_RE = r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)"


class Ins:
    __slots__ = ("addr", "text", "lo", "hi", "op", "dst", "src", "pinned", "var_lat")

    def field(self):
        c = self.hi >> 41
        return dict(stall=c & 15, y=(c >> 4) & 1, wb=(c >> 5) & 7, rb=(c >> 8) & 7, wait=(c >> 11) & 63,
                    reuse=(c >> 17) & 15)

    def with_ctrl(self, stall, y, reuse):
        c = self.hi >> 41
        c = (c & ~15) | (stall & 15)
        c = (c & ~(1 << 4)) | ((y & 1) << 4)
        c = (c & ~(15 << 17)) | ((reuse & 15) << 17)
        return (self.hi & ((1 << 41) - 1)) | (c << 41)

    def payload(self):
        """the encoding without stall / yield / reuse bits"""
        mask = ~(((15) | (1 << 4) | (15 << 17)) << 41) & ((1 << 64) - 1)
        return self.lo, self.hi & mask


def regs_of(tok):
    out = set()
    wide = 2 if (".F32x2" in tok or ".64" in tok) else 1
    for kind, n in re.findall(r"(?<![A-Za-z0-9_])(UR|R|UP|P)(\d+)", tok):
        n = int(n)
        if kind in ("R", "UR"):
            out.update(f"{kind}{n + k}" for k in range(wide))
        else:
            out.add(f"{kind}{n}")
    return out


def decode(ins):
    t = ins.text
    ins.pinned, ins.var_lat = False, False
    src, dst = set(), set()
    pred = re.match(r"^@(!?)(U?P\d)\s+", t)
    if pred:
        src.add(pred.group(2))
        t = t[pred.end():]
    parts = t.split(None, 1)
    op = ins.op = parts[0].split(".")[0]
    ops = [o.strip() for o in parts[1].split(",")] if len(parts) > 1 else []
    if op in PACKED:
        n = int(re.search(r"R(\d+)", ops[0]).group(1))
        dst |= {f"R{n}", f"R{n + 1}"}
        for o in ops[1:]:
            src |= regs_of(o)
    elif op in MINS:
        dst |= regs_of(ops[0])
        for o in ops[1:]:
            src |= regs_of(o)
    elif op in ("IADD3", "UIADD3"):
        dst |= regs_of(ops[0])
        rest = ops[1:]
        while rest and re.fullmatch(r"U?PT|U?P\d", rest[0]):   # carry-outs
            dst |= regs_of(rest[0])
            rest = rest[1:]
        for o in rest:
            src |= regs_of(o)
    elif op in ("ISETP", "UISETP"):
        dst |= regs_of(ops[0]) | regs_of(ops[1])
        for o in ops[2:]:
            src |= regs_of(o)
    elif op == "LDS":
        n = int(re.search(r"R(\d+)", ops[0]).group(1))
        width = 4 if ".128" in parts[0] else (2 if ".64" in parts[0] else 1)
        dst |= {f"R{n + k}" for k in range(width)}
        src |= regs_of(ops[1])
        ins.var_lat = True
    else:
        ins.pinned = True          # unknown: nothing may cross it
        for o in ops:
            src |= regs_of(o)
    ins.dst, ins.src = dst - {"RZ", "URZ"}, src - {"RZ", "URZ"}


def load(path, kernel):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, out, last, names = None, [], None, set()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or kernel not in cur:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;\s+/\* 0x([0-9a-f]{16}) \*/", line)
        if m:
            last = Ins()
            last.addr, last.text, last.lo = int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)
            continue
        m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/\s*$", line)
        if m and last is not None:
            last.hi = int(m.group(1), 16)
            decode(last)
            out.append(last)
            names.add(cur)
            last = None
    if not out:
        raise SystemExit(f"sass_sched: no kernel matching {kernel!r} in {path}")
    if len(names) != 1:
        raise SystemExit(f"sass_sched: {kernel!r} matches several kernels: {sorted(names)}")
    return out


def loops(instrs):
    idx = {i.addr: k for k, i in enumerate(instrs)}
    for k, i in enumerate(instrs):
        m = re.search(BRANCH_RE, i.text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt in idx and idx[tgt] <= k:
            body = instrs[idx[tgt]:k + 1]
            if not any("BRA" in b.text or "EXIT" in b.text or "BAR" in b.text for b in body[:-1]):
                yield tgt, body


def pick_loop(instrs, how):
    cands = list(loops(instrs))
    if not cands:
        raise SystemExit("sass_sched: no branch-free loop found")
    if how and how.startswith("0x"):
        for tgt, body in cands:
            if tgt == int(how, 16):
                return body
        raise SystemExit(f"sass_sched: no loop at {how}")
    top = max(sum(b.op in PACKED for b in body) for _, body in cands)
    best = [body for _, body in cands if sum(b.op in PACKED for b in body) == top]
    if how == "uniform":   # the unit-stride scan addresses shared memory through a uniform register
        best = [b for b in best if all("[UR" in i.text for i in b if i.op == "LDS")]
    if len(best) != 1:
        raise SystemExit(f"sass_sched: {len(best)} candidate loops; pass --loop=0xADDR")
    return best[0]


def fingerprint(body):
    """Identity of a loop as ptxas emitted it: the instruction texts with registers renamed in order of
    first appearance.  A plan (explicit order + yield mask) is only meaningful for the loop it was
    measured on; the same source compiled into another instruction order gets the generic recipe."""
    import hashlib
    names = {}

    def ren(m):
        return names.setdefault(m.group(0), f"{m.group(1)}#{len(names)}")
    txt = "\n".join(re.sub(r"(?<![A-Za-z0-9_])(UR|R|UP|P)(\d+)", ren, i.text) for i in body)
    return hashlib.md5(txt.encode()).hexdigest()


def movable(i):
    f = i.field()
    return i.op in MINS and f["wait"] == 0 and f["wb"] == 7 and f["rb"] == 7


def conflicts(a, b):
    return bool((a.dst & (b.dst | b.src)) or (a.src & b.dst))


def check_order(body, order):
    pos = {k: p for p, k in enumerate(order)}
    n = len(body)
    if sorted(order) != list(range(n)) or order[-1] != n - 1:
        raise SystemExit("sass_sched: not a permutation that keeps the branch last")
    for i in range(n):
        if not movable(body[i]):
            # everything that is not a movable minimum keeps its order among its kind
            for j in range(i + 1, n):
                if not movable(body[j]) and pos[i] > pos[j]:
                    raise SystemExit("sass_sched: a pinned instruction was re-ordered")
        for j in range(i + 1, n):
            if (conflicts(body[i], body[j]) or body[i].pinned or body[j].pinned) and pos[i] > pos[j]:
                raise SystemExit(f"sass_sched: dependency order broken:\n  {body[i].text}\n  {body[j].text}")


def model_times(body, order, enc):
    """lower bounds of the issue times: encoded stalls + two FMA-pipe cycles per packed op"""
    t, fma_free = [], 0
    for p, k in enumerate(order):
        e = 0 if p == 0 else t[p - 1] + enc[p - 1]
        if body[k].op in PACKED:
            e = max(e, fma_free)
            fma_free = e + 2
        t.append(e)
    return t


def pipe(i):
    return "fma" if i.op in PACKED else ("alu" if i.op in MINS else i.op)


def latency(prod, cons):
    if prod.op in PACKED and cons.op in PACKED:
        return LAT_FMA_FMA
    if prod.op in PACKED and cons.op in MINS:
        return LAT_FMA_MIN
    if prod.op in MINS and cons.op in MINS:
        return LAT_MIN_MIN
    return None     # keep the original gap


def assign_stalls(body, order, packed_stall):
    n = len(order)
    ident = list(range(n))
    t_orig = model_times(body, ident, [i.field()["stall"] for i in body])
    enc = []
    for p, k in enumerate(order):
        i = body[k]
        if k == n - 1:
            enc.append(i.field()["stall"])      # the branch keeps its own
        elif i.op in PACKED:
            enc.append(packed_stall)
        elif i.op in MINS:
            enc.append(1)
        else:
            enc.append(max(1, i.field()["stall"]) if i.pinned else 1)
    # raise stalls until every fixed-latency RAW gap holds
    for _ in range(4 * n):
        t = model_times(body, order, enc)
        last_write, fixed = {}, True
        for p, k in enumerate(order):
            i = body[k]
            need = 0
            for r in i.src | i.dst:
                q = last_write.get(r)
                if q is None:
                    continue
                prod = body[order[q]]
                if prod.var_lat:
                    continue
                if r not in i.src:
                    # write after write: results of one pipe retire in order, and a variable-latency
                    # write (LDS) lands long after any fixed-latency one
                    if i.var_lat or pipe(prod) == pipe(i):
                        continue
                lat = latency(prod, i) if r in i.src else None
                if lat is None:
                    lat = max(1, min(16, t_orig[k] - t_orig[order[q]]))  # original gap of this very pair (16 covers predicate -> branch)
                need = max(need, t[q] + lat)
            if need > t[p]:
                enc[p - 1] += need - t[p]
                if enc[p - 1] > 15:
                    raise SystemExit(f"sass_sched: stall > 15 needed before {i.text}")
                fixed = False
                break
            for r in i.dst:
                last_write[r] = p
        if fixed:
            break
    else:
        raise SystemExit("sass_sched: stall assignment did not converge")
    # loop-carried values: last writer in the body -> first reader of the next trip
    t = model_times(body, order, enc)
    total = t[-1] + enc[-1]
    first_read, last_write = {}, {}
    for p, k in enumerate(order):
        for r in body[k].src:
            first_read.setdefault(r, p)
        for r in body[k].dst:
            last_write[r] = p
    for r, q in last_write.items():
        c = first_read.get(r)
        if c is None or c > q or body[order[q]].var_lat:
            continue
        lat = latency(body[order[q]], body[order[c]]) or 16
        if total - t[q] + t[c] < lat:
            raise SystemExit(f"sass_sched: loop-carried latency on {r} would be violated")
    return enc, total


def sink(body, place):
    """walk the ptxas order; movable minima are held back and released by `place`"""
    n = len(body)
    order, held = [], []
    for k in range(n - 1):
        i = body[k]
        if movable(i):
            held.append(k)
            continue
        while any(conflicts(body[h], i) for h in held) or (i.pinned and held):
            order.append(held.pop(0))
        order.append(k)
        nxt = next((body[j] for j in range(k + 1, n) if not movable(body[j])), None)
        for h in place(body, held, order, nxt):
            held.remove(h)
            order.append(h)
    order.extend(held)
    order.append(n - 1)
    return order


def spaced(rules, extra):
    since = {op: 10 ** 6 for op in rules}

    def place(body, held, order, nxt):
        last = body[order[-1]]
        if last.op in PACKED:
            for op in since:
                since[op] += 1
        ok_next = "next" not in extra or (nxt is not None and nxt.op in extra["next"])
        if held and last.op in rules and since[last.op] > rules[last.op] and ok_next:
            since[last.op] = 0
            return held[:1]
        return []
    return place


def make_order(body, policy):
    if policy == "identity":
        return list(range(len(body)))
    if policy == "cluster_end":
        return sink(body, lambda b, held, order, nxt: [])
    if policy.startswith("spaced="):
        spec, extra = policy[7:], {}
        if "/" in spec:
            spec, tail = spec.split("/", 1)
            for kv in tail.split("/"):
                key, val = kv.split("=")
                extra[key] = tuple(val.split("+"))
        rules = {kv.split(":")[0]: int(kv.split(":")[1]) for kv in spec.split(",")}
        return sink(body, spaced(rules, extra))
    raise SystemExit(f"sass_sched: unknown policy {policy}")


def emit(body, order, stalls, yields):
    words = []
    for p, k in enumerate(order):
        i = body[k]
        keep_reuse = p + 1 < len(order) and order[p + 1] == k + 1
        f = i.field()
        words.append(struct.pack("<QQ", i.lo, i.with_ctrl(stalls[p], f["y"] if yields is None else yields[p],
                                                          f["reuse"] if keep_reuse else 0)))
    return b"".join(words)


STATE_TAG = b"P6D-SCHED-STATE:"      # followed by one letter per mesh-size class: p = ptxas, t = tuned


def patch_file(path, tmp, kernel_instrs, body, blob, mark=None):
    """Write the patched image to `tmp` (never over the input: the caller verifies `tmp` and only then
    moves it into place).  mark = index of the class letter to flip from 'p' to 't'."""
    data = bytearray(open(path, "rb").read())
    if mark is not None:
        # tell the library that this class's scan loop was re-laid (p6d_sched_state in p6d_add.cu);
        # a letter that is already 't' means the loop is no longer ptxas' -- never patch twice
        if data.count(STATE_TAG) != 1:
            raise SystemExit("sass_sched: state marker not found exactly once")
        at = data.find(STATE_TAG) + len(STATE_TAG) + mark
        if data[at:at + 1] != b"p":
            raise SystemExit(f"sass_sched: class {mark} is already marked {bytes(data[at:at + 1])!r}; refusing to patch twice")
        data[at:at + 1] = b"t"
    whole = b"".join(struct.pack("<QQ", i.lo, i.hi) for i in kernel_instrs)
    if data.count(whole) != 1:
        raise SystemExit(f"sass_sched: kernel image occurs {data.count(whole)} times in {path}")
    at = data.find(whole) + (body[0].addr - kernel_instrs[0].addr)
    old = b"".join(struct.pack("<QQ", i.lo, i.hi) for i in body)
    assert bytes(data[at:at + len(old)]) == old
    data[at:at + len(old)] = blob
    with open(tmp, "wb") as fh:
        fh.write(data)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = dict(a[2:].split("=", 1) if "=" in a else (a[2:], "1") for a in sys.argv[1:] if a.startswith("--"))
    if "plan" in opts:      # a schedule found by tools/sched_search.py: explicit order + yield mask
        import json
        plan = json.load(open(opts["plan"]))
        opts.setdefault("order", ",".join(str(k) for k in plan["order"]))
        opts.setdefault("yield", "mask" + plan["yield_mask"])
        opts.setdefault("packed-stall", str(plan.get("packed_stall", 1)))
        if len(args) == 1:
            args += [plan["kernel"], "plan"]
    path, kernel, policy = args[:3]
    out = opts.get("out") or (args[3] if len(args) > 3 else None)
    kernel_instrs = load(path, kernel)
    body = pick_loop(kernel_instrs, opts.get("loop"))
    if "fingerprint" in opts:
        print(fingerprint(body))
        return
    if "plan" in opts and plan.get("loop_fingerprint") != fingerprint(body):
        raise SystemExit(f"sass_sched: {opts['plan']} was measured on another instruction order of this loop "
                         f"(fingerprint {plan.get('loop_fingerprint')} != {fingerprint(body)}); not applied")
    order = [int(x) for x in opts["order"].split(",")] if "order" in opts else make_order(body, policy)
    check_order(body, order)
    n = len(order)
    if policy == "identity" and "packed-stall" not in opts:
        stalls = [i.field()["stall"] for i in body]
        total = model_times(body, order, stalls)[-1] + stalls[-1]
    else:
        stalls, total = assign_stalls(body, order, int(opts.get("packed-stall", 2)))
    yields = None
    y = opts.get("yield")
    if y in ("0", "1"):
        yields = [int(y)] * n
    elif y and y.startswith("mask"):
        yields = [int(c) for c in y[4:]]
        if len(yields) != n:
            raise SystemExit("sass_sched: yield mask length does not match the loop")
    elif y and y.startswith("period"):
        per, ph = (int(x) for x in y[6:].split(","))
        yields = [0 if (q % per) == ph else 1 for q in range(n)]
    npacked = sum(i.op in PACKED for i in body)
    moved = sum(1 for p, k in enumerate(order) if p != k)
    print(f"sass_sched: {kernel}: loop at 0x{body[0].addr:x}, {n} instructions ({npacked} packed, "
          f"{sum(movable(i) for i in body)} movable minima), {moved} positions changed, "
          f"one-warp length {total} cycles (packed ops alone {2 * npacked})")
    if "show" in opts:
        for p, k in enumerate(order):
            print(f"   st={stalls[p]:2d} y={(body[k].field()['y'] if yields is None else yields[p])}  {body[k].text}")
    if out:
        # a loop this pass (or anything else) has already re-laid is not ptxas' any more: the "keep the
        # original gap" rule would then trust gaps that are not the compiler's.  ptxas encodes the packed
        # ops of this loop with stall 2 (a handful with 1 next to a minimum); this pass with stall 1.
        short = sum(1 for i in body if i.op in PACKED and i.field()["stall"] < 2)
        if 4 * short > sum(1 for i in body if i.op in PACKED):
            raise SystemExit("sass_sched: the loop does not carry ptxas' schedule (already re-laid?); refusing")
        mark = int(opts["mark"]) if "mark" in opts else None
        tmp = out + ".sched-tmp"
        try:
            patch_file(path, tmp, kernel_instrs, body, emit(body, order, stalls, yields), mark)
            # read back: same instructions (payloads) in the planned order, same stalls
            again = pick_loop(load(tmp, kernel), "0x%x" % body[0].addr)
            want = [body[k].payload() for k in order]
            if [i.payload() for i in again] != want or [i.field()["stall"] for i in again] != stalls:
                raise SystemExit("sass_sched: read-back of the patched loop does not match the plan")
            os.chmod(tmp, os.stat(path).st_mode & 0o777)
            os.replace(tmp, out)              # only a verified image ever reaches `out`
        finally:
            if os.path.exists(tmp):
                os.remove(tmp)
        print(f"sass_sched: patched {out}")


if __name__ == "__main__":
    main()
