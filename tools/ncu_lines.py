"""Per-CUDA-line stall samples and instruction counts of an .ncu-rep captured with --import-source on:
python tools/ncu_lines.py report.ncu-rep [top_n] [file-substring]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":  # a CUDA line row (SASS rows carry an address)
        d = dict(zip(hdr, r))
        try:
            lines.append((int(d["# Samples"]), int(d["Instructions Executed"]), cur.split("/")[-1], int(r[0]), r[1].strip(), d))
        except ValueError:
            pass
tot_s, tot_i = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"# {rep}: {tot_s} samples, {tot_i} warp instructions")
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
for s, i, f, n, src, d in sorted(lines, key=lambda l: -l[0])[:top]:
    st = sorted(((int(d[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/max(tot_i,1):5.1f}% ins  {f}:{n:<5d} {' '.join(f'{k}={v}' for v,k in st if v):40s} | {src[:110]}")
