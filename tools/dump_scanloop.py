#!/usr/bin/env python3
"""Dump the ADD-S scan loop of one kernel of a built libp6d.so, one instruction per line with its
scheduling control field (stall, yield, scoreboards, reuse), so that the loop ptxas emitted and the
loop csrc/sass_sched.py re-laid can be reviewed side by side (profiles/scanloop_*.sass).

    python tools/dump_scanloop.py LIB KERNEL-SUBSTRING [> profiles/scanloop_n2048_relaid.sass]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "6d-pose-estimation_b200", "csrc"))
import sass_sched as S  # noqa: E402


def main():
    lib, kernel = sys.argv[1], sys.argv[2]
    body = S.pick_loop(S.load(lib, kernel), "uniform")
    t = S.model_times(body, list(range(len(body))), [i.field()["stall"] for i in body])
    print(f"# {kernel}: loop at 0x{body[0].addr:x}, {len(body)} instructions, "
          f"{sum(i.op in S.PACKED for i in body)} packed, {sum(i.op in S.MINS for i in body)} minima; "
          f"modelled one-warp length {t[-1] + body[-1].field()['stall']} cycles")
    print("# addr   stall y wb rb wait reuse  instruction")
    for i in body:
        f = i.field()
        print(f"{i.addr:06x}  {f['stall']:2d}    {f['y']} {f['wb']}  {f['rb']}  {f['wait']:02x}   {f['reuse']:x}      {i.text}")


if __name__ == "__main__":
    main()
