"""Second plain-copy probe: which downloads does an upload slow down?  Downloads in 27 pieces on one stream, per-piece GB/s printed;
uploads (a) one long copy, a little late, (b) 27 pieces back to back, (c) one piece per download piece, released by an event
recorded behind a small head of that download piece plus a delay.  python tools/copy_probe2.py"""
import sys
import time

import torch

up_bytes, dn_bytes, pieces = 1093342208, 5419008000, 27
h_up = torch.empty(up_bytes, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(dn_bytes, dtype=torch.uint8).pin_memory()
d_up = torch.empty(up_bytes, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(dn_bytes, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
CYC_PER_MS = 1.9e6


def cut(n, k):
    step = (n + k - 1) // k
    return [(i, min(n, i + step)) for i in range(0, n, step)]


def run(mode, delay_ms=0.0, head=1 << 20):
    dn_cut, up_cut = cut(dn_bytes, pieces), cut(up_bytes, pieces)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(pieces + 1)]
    heads = [torch.cuda.Event() for _ in range(pieces)]
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s_dn):
        ev[0].record()
        for i, (a, b) in enumerate(dn_cut):
            if mode == "gated":
                h_dn[a:a + head].copy_(d_dn[a:a + head], non_blocking=True)
                heads[i].record()
                h_dn[a + head:b].copy_(d_dn[a + head:b], non_blocking=True)
            else:
                h_dn[a:b].copy_(d_dn[a:b], non_blocking=True)
            ev[i + 1].record()
    with torch.cuda.stream(s_up):
        if mode == "one_late":
            torch.cuda._sleep(int(delay_ms * CYC_PER_MS))
            u0.record()
            d_up.copy_(h_up, non_blocking=True)
        elif mode == "pieces":
            u0.record()
            for a, b in up_cut:
                d_up[a:b].copy_(h_up[a:b], non_blocking=True)
        elif mode == "gated":
            u0.record()
            for i, (a, b) in enumerate(up_cut):
                s_up.wait_event(heads[i])
                if delay_ms:
                    torch.cuda._sleep(int(delay_ms * CYC_PER_MS))
                d_up[a:b].copy_(h_up[a:b], non_blocking=True)
        else:
            u0.record()
        u1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    rates = [(b - a) / max(ev[i].elapsed_time(ev[i + 1]), 1e-6) / 1e6 for i, (a, b) in enumerate(dn_cut)]
    return wall, u0.elapsed_time(u1), ev[0].elapsed_time(ev[pieces]), rates


for name, kw in [("download alone", dict(mode="none")),
                 ("one upload, 0.3 ms late", dict(mode="one_late", delay_ms=0.3)),
                 ("one upload, 2 ms late", dict(mode="one_late", delay_ms=2.0)),
                 ("27 upload pieces back to back", dict(mode="pieces")),
                 ("upload piece i released by the head of download piece i", dict(mode="gated")),
                 ("... plus 0.2 ms", dict(mode="gated", delay_ms=0.2)),
                 ("... plus 1 ms", dict(mode="gated", delay_ms=1.0))]:
    run(**kw)
    best = min((run(**kw) for _ in range(3)), key=lambda r: r[0])
    print(f"{name:58s}: wall {best[0]:6.1f} ms, uploads span {best[1]:6.1f} ms, downloads {best[2]:6.1f} ms; per piece GB/s: "
          + " ".join(f"{r:.0f}" for r in best[3]), flush=True)
