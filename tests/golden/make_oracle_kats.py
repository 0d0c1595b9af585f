#!/usr/bin/env python
"""Generate tests/golden/oracle_kats.json: intermediate known answers from the CPU oracle.

Runs oracle/_build/nbody_oracle (full 200000-step, three-query solve) on every golden input and
records the values the reference's goldens do not hold (argmin step, per-device missile reach
step, per-device query-3 outcome) next to the three output lines.  The three output lines are
checked against testcases/bN.out on the spot: a mismatch aborts.

usage: python tests/golden/make_oracle_kats.py [--cases b20,b30,...] [--mode 0|1] [--threads N]
Strict mode (0, pow) is used up to b200 by default, sqrt3 mode (1) for b512/b1024 (hours otherwise).
"""
import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CASES = ["b20", "b30", "b40", "b50", "b60", "b70", "b80", "b90", "b100", "b200", "b512", "b1024"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--mode", type=int, default=None)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--out", default=os.path.join(HERE, "oracle_kats.json"))
    args = ap.parse_args()
    exe = os.path.join(ROOT, "oracle", "_build", "nbody_oracle")
    kats = {}
    if os.path.exists(args.out):
        kats = json.load(open(args.out))
    for case in args.cases.split(","):
        inp = os.path.join(HERE, "testcases", case + ".in")
        gold = open(os.path.join(HERE, "testcases", case + ".out")).read()
        n = int(open(inp).readline().split()[0])
        mode = args.mode if args.mode is not None else (0 if n <= 200 else 1)
        t0 = time.time()
        subprocess.check_call([exe, inp, "/tmp/kat_%s.out" % case, "200000", str(mode), str(args.threads),
                               "/tmp/kat_%s.json" % case])
        dt = time.time() - t0
        out = open("/tmp/kat_%s.out" % case).read()
        if out != gold:
            sys.exit("oracle output for %s differs from the golden:\n%s\nvs\n%s" % (case, out, gold))
        k = json.load(open("/tmp/kat_%s.json" % case))
        k["oracle_seconds"] = round(dt, 1)
        k["byte_identical_to_golden"] = True
        kats[case] = k
        json.dump(kats, open(args.out, "w"), indent=1, sort_keys=True)
        print(case, "ok mode", mode, "%.1fs" % dt, flush=True)


if __name__ == "__main__":
    main()
