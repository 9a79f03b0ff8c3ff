"""Address-ordered SASS of the instructions outside a source-line window, with the stall samples converted to cycles per execution
(single resident warp: samples are proportional to time): python tools/ncu_timeline.py rep file lo hi cycles_per_iter exec_count"""
import csv, subprocess, sys
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = cur = curline = None
sass = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]
    elif r and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if r[2] == "-": curline = (cur, int(r[0]))
        else:
            try: sass.append((int(r[2], 16), r[3].strip(), int(d["# Samples"]), int(d["Instructions Executed"]), curline, d))
            except ValueError: pass
sass.sort()
tot = sum(x[2] for x in sass); base = sass[0][0]
mx = max(x[3] for x in sass)
from collections import Counter
common = Counter(x[3] for x in sass if x[3] > 0).most_common(1)[0][0]
inwin = lambda ln: ln[0] == fname and lo <= ln[1] <= hi
hot = [x for x in sass if x[3] >= common // 2]
in_s = sum(x[2] for x in hot if inwin(x[4])); out_s = sum(x[2] for x in hot if not inwin(x[4]))
print(f"# total samples {tot}; per-iteration exec count {common}; window share {in_s/tot:.3f}, outside {out_s/tot:.3f}; "
      f"instr in window {sum(1 for x in hot if inwin(x[4]))}, outside {sum(1 for x in hot if not inwin(x[4]))}")
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 0.001
for a, ins, smp, ex, ln, d in hot:
    if inwin(ln) or smp / tot < thr: continue
    st = sorted(((int(d[c] or 0), c[6:]) for c in hdr if c.startswith('stall_') and 'Not' not in c), reverse=True)[0]
    print(f"{a-base:6x} {100*smp/tot:5.2f}% {ln[0][:16]}:{ln[1]:<4d} {st[1]:14s} {ins[:84]}")
