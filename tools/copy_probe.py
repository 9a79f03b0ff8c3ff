"""Plain-copy controls for the host-buffer decode (no codec involved): how much does an upload crossing the link slow the download,
as one long copy each or cut into pieces?  python tools/copy_probe.py"""
import time

import torch

up_bytes, dn_bytes, pieces = 1093342208, 5419008000, 27
h_up = torch.empty(up_bytes, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(dn_bytes, dtype=torch.uint8).pin_memory()
d_up = torch.empty(up_bytes, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(dn_bytes, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def cut(n, k):
    step = (n + k - 1) // k
    return [(i, min(n, i + step)) for i in range(0, n, step)]


def run(up_pieces, dn_pieces, do_up=True, do_dn=True, delay_up_ms=0.0):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if do_dn:
        with torch.cuda.stream(s_dn):
            ev[2].record()
            for a, b in cut(dn_bytes, dn_pieces):
                h_dn[a:b].copy_(d_dn[a:b], non_blocking=True)
            ev[3].record()
    if do_up:
        with torch.cuda.stream(s_up):
            if delay_up_ms:
                torch.cuda._sleep(int(delay_up_ms * 1.9e6))
            ev[0].record()
            for a, b in cut(up_bytes, up_pieces):
                d_up[a:b].copy_(h_up[a:b], non_blocking=True)
            ev[1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    up = ev[0].elapsed_time(ev[1]) if do_up else 0.0
    dn = ev[2].elapsed_time(ev[3]) if do_dn else 0.0
    return wall, up, dn


for name, kw in [("upload alone, 1 piece", dict(up_pieces=1, dn_pieces=1, do_dn=False)),
                 ("download alone, 1 piece", dict(up_pieces=1, dn_pieces=1, do_up=False)),
                 ("download alone, 27 pieces", dict(up_pieces=1, dn_pieces=pieces, do_up=False)),
                 ("both, 1 piece each", dict(up_pieces=1, dn_pieces=1)),
                 ("both, upload 27 pieces, download 1", dict(up_pieces=pieces, dn_pieces=1)),
                 ("both, upload 1, download 27 pieces", dict(up_pieces=1, dn_pieces=pieces)),
                 ("both, 27 pieces each", dict(up_pieces=pieces, dn_pieces=pieces)),
                 ("both, 270 upload pieces, 27 download pieces", dict(up_pieces=270, dn_pieces=pieces)),
                 ("both, 1 piece each, upload 30 ms late", dict(up_pieces=1, dn_pieces=1, delay_up_ms=30.0))]:
    run(**kw)
    best = min((run(**kw) for _ in range(3)), key=lambda r: r[0])
    print(f"{name:46s}: wall {best[0]:7.1f} ms, upload {best[1]:6.1f} ms ({up_bytes / max(best[1], 1e-9) / 1e6:5.1f} GB/s), "
          f"download {best[2]:6.1f} ms ({dn_bytes / max(best[2], 1e-9) / 1e6:5.1f} GB/s)", flush=True)
