"""Static schedule of a SASS range: python tools/sass_sched.py file.sass(from `nvdisasm -hex`) [start_addr end_addr]
Prints every instruction with the stall count, yield flag, scoreboard set (W/R) and wait mask decoded from the control bits
(bits 105..125 of the 128-bit encoding), and the sum of stall counts = issue cycles of one warp running alone."""
import re, sys
lines = open(sys.argv[1]).read().splitlines()
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
ins = []
i = 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        if m2:
            ins.append((int(m.group(1), 16), m.group(2).strip(), int(m2.group(1), 16)))
            i += 2
            continue
    i += 1
tot = n = 0
for a, t, h in ins:
    if not (lo <= a <= hi):
        continue
    c = h >> 41
    stall, yld, wb, rb, wm = c & 15, (c >> 4) & 1, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 63
    tot += stall
    n += 1
    print(f"{a:05x} s{stall:2d} {'Y' if yld else ' '} W{wb if wb != 7 else '-'} R{rb if rb != 7 else '-'} wait{wm:02x}  {t}")
print(f"# {n} instructions, stall sum {tot} cycles, {tot / max(n, 1):.2f} per instruction")
