"""Decode the scheduling control fields (stall count, yield, barriers, reuse) of a kernel's SASS (Volta+ encoding:
upper 64-bit word, stall = bits 41..44, yield = 45, write barrier = 46..48, read barrier = 49..51, wait mask = 52..57)."""
import re, subprocess, sys
obj, fun = sys.argv[1], sys.argv[2]
lo, hi = (int(sys.argv[3], 0), int(sys.argv[4], 0)) if len(sys.argv) > 4 else (0, 10**9)
out = subprocess.check_output(["cuobjdump", "-sass", obj]).decode()
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks:
    if fun not in b.split("\n", 1)[0]:
        continue
    lines = b.split("\n")
    k = 0
    i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                w = int(m2.group(1), 16)
                stall, yld, wb, rb, wm = (w >> 41) & 0xf, (w >> 45) & 1, (w >> 46) & 7, (w >> 49) & 7, (w >> 52) & 0x3f
                addr = int(m.group(1), 16)
                if lo <= addr <= hi:
                    print("%05x  st=%2d y=%d wb=%d rb=%d wait=%02x  %s" % (addr, stall, yld, wb, rb, wm, m.group(2).strip()))
                i += 2
                continue
        i += 1
