"""Opcode mix of the unrolled decode body: python tools/sass_stats.py [C] [B] [MODE] [obj] [S] [PAIR]  (reads sea_codec_b200/build/decode_fast.o,
decode_fast_mono.o for C = 1; MODE 2 = pair-replicated table, 0 = plain; S = scale_factor_bits instance 3 / 4 / 5, 0 = run-time)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C, B, R = (sys.argv[1:4] + ["2", "3", "2"][len(sys.argv) - 1:])[:3]
obj = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] else os.path.join(ROOT, "sea_codec_b200", "build", "decode_fast.o" if C == "2" else "decode_fast_mono.o")
S = sys.argv[5] if len(sys.argv) > 5 else "4"
P = sys.argv[6] if len(sys.argv) > 6 else "0"
fun = f"_ZN3sea22decode_unrolled_kernelILi{C}ELi{B}ELi{R}ELi{S}ELb{P}EEEvPKhPsPKNS_9DecStreamENS_13DecFastParamsEPKiPi"
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
back = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        back.append((a, int(m.group(1), 16)))
a1, a0 = sorted(back, key=lambda x: x[0] - x[1])[-2]  # the half loop is the second largest loop (the round loop encloses it)
c = collections.Counter()
for a, t in ins:
    if a0 <= a <= a1:
        toks = t.split()
        c[toks[1] if toks[0].startswith("@") else toks[0]] += 1
tot = sum(c.values())
print(f"kernel <{C},{B},{R}>: {len(ins)} instructions, half body {tot} = {tot / 80:.3f} per sample")
for k, v in c.most_common(int(os.environ.get("TOP", "18"))):
    print(f"  {k:24s} {v:5d} {v / 80:.3f}")
