"""Count, per DFMA/DMUL/DADD in a kernel's SASS, the distinct source register pairs and .reuse flags."""
import re, subprocess, sys, collections
obj, fun = sys.argv[1], sys.argv[2]
out = subprocess.check_output(["cuobjdump", "-sass", obj]).decode()
# split by function
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks:
    name = b.split("\n", 1)[0]
    if fun not in name: continue
    stat = collections.Counter()
    prev = None
    for line in b.split("\n"):
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?(DFMA|DMUL|DADD)\s+(.*?);", line)
        if not m: prev = None if re.search(r"\*/\s+\S", line) and not m else prev; continue
        op, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
        srcs = args[1:]
        regs = [re.sub(r"[-|]|\.reuse", "", a) for a in srcs if re.match(r"-?\|?R\d", a)]
        distinct = len(set(regs))
        reuse = sum(".reuse" in a for a in srcs)
        # a slot whose register equals the previous FP64 instruction's same slot with .reuse set there is free
        free = 0
        if prev is not None:
            for k, a in enumerate(srcs):
                if k < len(prev) and ".reuse" in prev[k] and re.sub(r"\.reuse", "", prev[k]) == re.sub(r"\.reuse", "", a):
                    free += 1
        stat[(op, distinct, free)] += 1
        prev = srcs
    print(name[:100])
    tot = 0; cyc = 0
    for (op, d, f), c in sorted(stat.items()):
        eff = max(2, d - f)
        print("  %-5s distinct=%d reused_from_prev=%d : %4d  -> %d cyc" % (op, d, f, c, eff))
        tot += c; cyc += c * eff
    print("  total %d FP64 instr, est. %d pipe cycles, %.3f cyc/instr, max inst-rate %.1f%%" % (tot, cyc, cyc / tot, 200.0 * tot / cyc))
