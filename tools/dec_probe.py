"""Small decode-only workload for profiling / tuning: python tools/dec_probe.py [streams] [seconds] [bits] [channels] [iters] [frames]
Honours SEA_B200_LIB (tuning builds from tools/build_variant.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sea_codec_b200 as S
from sea_codec_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
secs = int(sys.argv[2]) if len(sys.argv) > 2 else 60
bits = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
ch = int(sys.argv[4]) if len(sys.argv) > 4 else 2
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = torch.device("cuda:0")
ctx = S.Context(0)
if os.environ.get("PROBE_STREAM") == "default":
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)   # the legacy default stream
elif os.environ.get("PROBE_STREAM") == "torch":
    _ts = torch.cuda.Stream()
    ctx.set_stream(_ts.cuda_stream)
frames = int(sys.argv[6]) if len(sys.argv) > 6 else secs * 44100  # argv[6]: exact frame count (e.g. a multiple of 5120: no partial chunk)
vbr = bool(int(os.environ.get("PROBE_VBR", "0")))
sfb = int(os.environ.get("PROBE_SFB", "4"))
st = S.EncoderSettings(residual_bits=bits, vbr=vbr, scale_factor_bits=sfb)
u = next(k for k in range(min(n, 16), 0, -1) if n % k == 0)
pcm = synth.gen_batch_torch(u, frames, ch, 44100, dev)
bound = ctx.encode_bound(frames, ch, st)
stride = (bound + 15) // 16 * 16
sea_u = torch.zeros(u * stride, dtype=torch.uint8, device=dev)
ctx.encode_batch_device(pcm.data_ptr(), np.arange(u) * frames * ch, np.full(u, frames), 44100, ch, st, sea_u.data_ptr(), np.arange(u) * stride)
sea = sea_u.view(u, stride).repeat(n // u, 1).contiguous().view(-1)
headers = np.tile(sea_u.view(u, stride)[:, :22].cpu().numpy(), (n // u, 1))
spp = frames * ch
out = torch.empty(n * spp, dtype=torch.int16, device=dev)
ms = []
for i in range(iters):
    ctx.decode_batch_device(sea.data_ptr(), np.arange(n, dtype=np.uint64) * stride, np.full(n, bound, dtype=np.uint64), headers, out.data_ptr(),
                            np.arange(n, dtype=np.uint64) * spp)
    ms.append(ctx.last_kernel_ms)
ok = bool(torch.equal(out.view(n, spp)[:u].view(u, frames, ch), pcm.view(u, frames, ch)) is False)  # lossy codec: only replica equality below
rep = torch.equal(out.view(n, spp)[n - 1], out.view(n, spp)[(n - 1) % u])
best = min(ms[1:]) if len(ms) > 1 else ms[0]
if os.environ.get("PROBE_VERBOSE"):
    print("all iterations (ms):", " ".join(f"{m:.3f}" for m in ms))
sustained = float(np.mean(ms[len(ms) // 2:])) if len(ms) >= 20 else None
alg = n * bound + 2 * n * spp
print(f"lib={os.path.basename(os.environ.get('SEA_B200_LIB', 'default'))} n={n} ch={ch} bits={bits}{' vbr' if vbr else ''}{'' if sfb == 4 else f' sfb={sfb}'}: best {best:.3f} ms  "
      f"{n*spp/best/1e3:.0f} Msamples/s  {alg/best/1e6:.0f} GB/s ({alg/best/1e6/6550.4*100:.1f}% of HBM peak) replicas_equal={rep}"
      + (f"  sustained {sustained:.3f} ms ({alg/sustained/1e6/6550.4*100:.1f}%)" if sustained else ""))
