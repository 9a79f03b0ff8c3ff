"""Per-call latency of the one-chunk streaming seam (sea_b200_encoder_make_chunk / sea_b200_decoder_decode_chunk) next to the CPU
oracle's per-chunk time: python tools/latency_probe.py [channels] [reps]"""
import ctypes as C, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sea_codec_b200 as S
from sea_codec_b200 import api, synth


def measure(ctx, channels=2, reps=60, vbr=False, bits=3.0):
    L = api.lib()
    st = S.EncoderSettings(residual_bits=bits, vbr=vbr)
    stc = st._c()
    n_chunks = reps + 5
    pcm = synth.gen_stream(9, 5120 * n_chunks, channels, 44100)
    h = C.c_void_p()
    assert L.sea_b200_encoder_create(ctx._h, channels, 44100, C.byref(stc), C.byref(h)) == 0
    out = np.zeros(70000, dtype=np.uint8)
    n = C.c_uint64(0)
    chunks, enc_us = [], []
    for k in range(n_chunks):
        x = np.ascontiguousarray(pcm[k * 5120 * channels: (k + 1) * 5120 * channels])
        t0 = time.perf_counter()
        rc = L.sea_b200_encoder_make_chunk(h, x.ctypes.data, x.size, out.ctypes.data, out.size, C.byref(n))
        enc_us.append((time.perf_counter() - t0) * 1e6)
        assert rc == 0
        chunks.append(out[: n.value].copy())
    cs = L.sea_b200_encoder_chunk_size(h)
    L.sea_b200_encoder_destroy(h)
    hdr = np.frombuffer(api._serialize_header(channels, cs, 5120, 44100, 0), dtype=np.uint8)
    d = C.c_void_p()
    assert L.sea_b200_decoder_create(ctx._h, hdr.ctypes.data, 22, C.byref(d)) == 0
    dec = np.zeros(5120 * channels, dtype=np.int16)
    dec_us, k_ms = [], []
    for ck in chunks:
        t0 = time.perf_counter()
        rc = L.sea_b200_decoder_decode_chunk(d, ck.ctypes.data, ck.size, -1, dec.ctypes.data, dec.size, C.byref(n))
        dec_us.append((time.perf_counter() - t0) * 1e6)
        k_ms.append(ctx.last_kernel_ms)
        assert rc == 0 and n.value == 5120 * channels
    L.sea_b200_decoder_destroy(d)
    return {"channels": channels, "vbr": vbr, "residual_bits": bits, "calls": reps,
            "make_chunk_us_median": float(np.median(enc_us[5:])), "make_chunk_us_p90": float(np.percentile(enc_us[5:], 90)),
            "decode_chunk_us_median": float(np.median(dec_us[5:])), "decode_chunk_us_p90": float(np.percentile(dec_us[5:], 90)),
            "decode_kernel_us_median": float(np.median(k_ms[5:])) * 1e3,
            "chunk_bytes": int(cs), "frames_per_chunk": 5120}


def cpu_reference(channels=2, vbr=False, bits=3.0, reps=20):
    """The oracle's per-chunk times (a restatement of the reference's scalar loop; single thread)."""
    from oracle import sea_oracle as O
    pcm = synth.gen_stream(9, 5120 * reps, channels, 44100)
    st = O.make_settings(bits, vbr=vbr)
    t0 = time.perf_counter()
    enc = O.sea_encode(pcm, 44100, channels, st)
    enc_us = (time.perf_counter() - t0) * 1e6 / reps
    t0 = time.perf_counter()
    O.sea_decode(enc)
    dec_us = (time.perf_counter() - t0) * 1e6 / reps
    return {"make_chunk_us": enc_us, "decode_chunk_us": dec_us, "kind": "port (oracle, 1 thread)"}


if __name__ == "__main__":
    ch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    ctx = S.Context(0)
    for vbr in (False, True):
        print(measure(ctx, ch, reps, vbr), cpu_reference(ch, vbr))
