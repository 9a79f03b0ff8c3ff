"""Host-buffer decode (sea_b200_decode_batch) against a plain-copy control, with the pipeline's switches:
python tools/e2e_probe.py [streams] [seconds]   -- prints ms per call for SEA_B200_DEC_DEFER = 1 / 0 and several group sizes,
then one traced call (SEA_B200_TRACE: per-group download start / duration on stderr)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sea_codec_b200 as S  # noqa: E402
from bench import Batch, encode_device, RATE, CHANNELS  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 60
own_stream = len(sys.argv) > 3 and sys.argv[3] == "own"
dev = torch.device("cuda", 0)
ctx = S.Context(0)
if not own_stream:
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
frames = seconds * RATE
b = Batch(torch, dev, n, frames, CHANNELS)
ctx.synth_pcm_device(b.pcm.data_ptr(), b.spp, np.arange(n, dtype=np.uint32), frames, CHANNELS, RATE)
sea, stride, lens, _ = encode_device(ctx, torch, dev, b, n, S.EncoderSettings(), RATE)
spp = b.spp
pcm_out = torch.empty(n * spp, dtype=torch.int16, device=dev)
headers = sea.view(n, stride)[:, :22].cpu().numpy()
off = np.arange(n, dtype=np.uint64)
ctx.decode_batch_device(sea.data_ptr(), off * stride, lens, headers, pcm_out.data_ptr(), off * spp)
ref = pcm_out.cpu().numpy()
h_sea = torch.empty(n * stride, dtype=torch.uint8).pin_memory()
h_pcm = torch.empty(n * spp, dtype=torch.int16).pin_memory()
h_sea.copy_(sea)
torch.cuda.synchronize()
bytes_total = n * int(lens[0]) + n * spp * 2


def call():
    ctx.decode_batch_host(h_sea.data_ptr(), off * stride, lens, h_pcm.data_ptr(), off * spp)


def timed(reps=4):
    call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def control(reps=4, duplex=True):
    def step():
        if duplex:
            with torch.cuda.stream(s_up):
                sea.copy_(h_sea, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_pcm.copy_(pcm_out, non_blocking=True)
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


c_ms, d_ms = control(), control(duplex=False)
print(f"control: H2D + D2H at once {c_ms:.1f} ms ({bytes_total / c_ms / 1e6:.1f} GB/s), D2H alone {d_ms:.1f} ms ({n * spp * 2 / d_ms / 1e6:.1f} GB/s)")
for ahead, defer in (("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")):
    for grp in (48, 96, 192):
        os.environ["SEA_B200_DEC_UPLOAD_AHEAD"] = ahead
        os.environ["SEA_B200_DEC_DEFER"] = defer
        os.environ["SEA_B200_DEC_GROUP_SAMPLES"] = str(grp << 20)
        ms = timed()
        h_pcm.zero_()
        call()
        ok = np.array_equal(h_pcm.numpy(), ref)
        print(f"upload_ahead={ahead} defer={defer} group={grp} Msamples: {ms:.1f} ms per call = {n * spp / ms / 1e3:.0f} Msamples/s, {c_ms / ms:.3f} of the control, bit-exact={ok}", flush=True)
c2 = control()
print(f"control again: {c2:.1f} ms")
for ahead in ("1", "0"):
    os.environ["SEA_B200_DEC_UPLOAD_AHEAD"] = ahead
    os.environ["SEA_B200_DEC_DEFER"] = "1"
    os.environ["SEA_B200_DEC_GROUP_SAMPLES"] = str(96 << 20)
    os.environ["SEA_B200_TRACE"] = "1"
    sys.stderr.write(f"--- trace, upload_ahead={ahead}\n")
    call()
    del os.environ["SEA_B200_TRACE"]
