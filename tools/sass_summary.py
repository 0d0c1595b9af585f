"""Per-kernel SASS instruction counts of the built library (profiles/r02_sass_summary.md): FP64 instructions, MUFU seeds,
shuffles, TMA bulk copies (UBLKCP, .MULTICAST), cluster barriers (UCGABAR), mbarrier ops (SYNCS), 256-bit sector
accesses (ENL2.256), fences, reductions."""
import collections, re, subprocess, sys
so = sys.argv[1]
out = subprocess.check_output(["cuobjdump", "-sass", so]).decode()
keys = ["DFMA", "DFMA.reuse", "DMUL", "DADD", "MUFU.RSQ64H", "MUFU.RCP64H", "SHFL", "LDS", "UBLKCP", "MULTICAST", "UCGABAR", "SYNCS",
        "ENL2.256", "MEMBAR", "RED"]
print("| kernel | " + " | ".join(keys) + " |")
print("|---|" + "---:|" * len(keys))
for b in re.split(r"\n\s*Function : ", out)[1:]:
    name = b.split("\n", 1)[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"nb::\(anonymous namespace\)::|nb::sym::\(anonymous namespace\)::|\(anonymous namespace\)::", "", dem)
    dem = re.sub(r"\(.*", "", dem)
    if any(x in dem for x in ("dfma", "pack", "unpack", "wait_kernel")):
        continue
    c = collections.Counter()
    for line in b.split("\n"):
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for k in ("UBLKCP", "UCGABAR", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU.RSQ64H", "MUFU.RCP64H", "SHFL", "LDS", "MEMBAR", "RED"):
            if op.startswith(k):
                c[k] += 1
        c["MULTICAST"] += ".MULTICAST" in line
        c["ENL2.256"] += "ENL2.256" in line
        c["DFMA.reuse"] += (".reuse" in line and op.startswith("DFMA"))
    print("| `%s` | " % dem[:60] + " | ".join(str(c.get(k, 0)) for k in keys) + " |")
