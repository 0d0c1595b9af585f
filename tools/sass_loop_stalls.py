"""Sum of ptxas' own stall counts (scheduling control field, bits 41..44 of the upper instruction word) over every loop of a
kernel that contains at least `min_fp64` FP64 instructions: what ONE warp needs per trip when nothing but the fixed-latency
dependencies holds it back (variable latencies - MUFU, LDS - come on top).  Used for the pair loops of nb_grid.cu:
363 cycles per 8 pairs at 166 registers, 319 after the register re-allocation, 668 when the role branches were not dominated
by their setmaxnreg (profiles/r02_grid_exchange.md, sections 5.6 and 5.7).

    python tools/sass_loop_stalls.py <cubin|.o|.so> <kernel name substring> [min_fp64]
"""
import re
import subprocess
import sys


def main():
    obj, fun = sys.argv[1], sys.argv[2]
    min_fp64 = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = subprocess.check_output(["cuobjdump", "-sass", obj]).decode()
    for b in re.split(r"\n\s*Function : ", out)[1:]:
        name = b.split("\n", 1)[0]
        if fun not in name:
            continue
        lines = b.split("\n")
        ins = []
        i = 0
        while i < len(lines):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", lines[i])
            if m and i + 1 < len(lines):
                m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
                if m2:
                    w = int(m2.group(1), 16)
                    ins.append((int(m.group(1), 16), m.group(2).strip(), (w >> 41) & 0xf))
                    i += 2
                    continue
            i += 1
        idx = {a: k for k, (a, _, _) in enumerate(ins)}
        print(name[:110])
        for k, (a, t, st) in enumerate(ins):
            m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
            if not m:
                continue
            tg = int(m.group(1), 16)
            if tg <= a and tg in idx:
                seg = ins[idx[tg]:k + 1]
                nfp = sum(1 for _, t2, _ in seg if re.match(r"(@!?U?P\d+\s+)?(DFMA|DMUL|DADD)", t2))
                if nfp >= min_fp64:
                    stalls = sum(s for _, _, s in seg)
                    print("  loop 0x%x..0x%x: %d instr, %d FP64, sum of stall counts %d -> one warp alone keeps the FP64 pipe "
                          "%.0f %% busy (2.13 cycles per FP64 instruction)" % (tg, a, len(seg), nfp, stalls, 100 * 2.13 * nfp / stalls))


if __name__ == "__main__":
    main()
