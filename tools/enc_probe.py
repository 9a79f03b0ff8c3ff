"""Small encode-only workload for profiling: python tools/enc_probe.py [streams] [seconds] [bits] [vbr]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sea_codec_b200 as S
from sea_codec_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
secs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
bits = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
vbr = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
ch = int(sys.argv[5]) if len(sys.argv) > 5 else 2
dev = torch.device("cuda:0")
ctx = S.Context(0)
frames = secs * 44100
st = S.EncoderSettings(residual_bits=bits, vbr=vbr)
u = next(k for k in range(min(n, 16), 0, -1) if n % k == 0)
pcm = synth.gen_batch_torch(u, frames, ch, 44100, dev).repeat(n // u, 1).contiguous()
bound = ctx.encode_bound(frames, ch, st)
stride = (bound + 15) // 16 * 16
out = torch.zeros(n * stride, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for i in range(3):
    lens = ctx.encode_batch_device(pcm.data_ptr(), np.arange(n) * frames * ch, np.full(n, frames), 44100, ch, st, out.data_ptr(), np.arange(n) * stride)
    print(f"iter {i}: {ctx.last_kernel_ms:.2f} ms  {n*frames*ch/ctx.last_kernel_ms/1e3:.1f} Msamples/s ties={ctx.last_vbr_ties}")
