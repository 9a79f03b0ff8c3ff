"""Deterministic synthetic 16-bit PCM (SURVEY.md 8d "Synthetic input"), integer-only so that numpy on the
host and torch on the GPU produce the same samples bit for bit.

x[n] = (A * sin_tab[phase(n) >> 20] >> 15) + noise(n)
  phase(n)  = (phase0_c + n * step_k) mod 2^32,  step_k = round(f_k / rate * 2^32), f_k = 220 * 2^((k mod 48)/12) Hz
  noise(n)  = splitmix64(seed_{k,c} + n) mapped to [-amp, +amp], amp = 0.05 * 32767
  seed_{k,c} = 0x5EA0000 + k * 256 + c   (stream k, channel c)
The 4096-entry sine table is generated once in float64 and rounded; both back-ends read the same table.
"""
from __future__ import annotations

import numpy as np

_TAB_BITS = 12
_A = 16383  # 0.5 * 32767
_NOISE_AMP = 1638  # 0.05 * 32767
_MASK64 = (1 << 64) - 1
SEED = 0x5EA0000


def sine_table() -> np.ndarray:
    n = 1 << _TAB_BITS
    return np.round(32767.0 * np.sin(2.0 * np.pi * (np.arange(n, dtype=np.float64) + 0.5) / n)).astype(np.int32)


def _step(k: int, rate: int) -> int:
    f = 220.0 * 2.0 ** ((k % 48) / 12.0)
    return int(round(f / rate * 4294967296.0)) & 0xFFFFFFFF


def _splitmix64_np(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def gen_stream(stream: int, n_frames: int, channels: int, rate: int, seed: int = 0x5EA0000) -> np.ndarray:
    """Interleaved int16 [n_frames * channels] for one stream (numpy)."""
    tab = sine_table()
    n = np.arange(n_frames, dtype=np.uint64)
    out = np.empty((n_frames, channels), dtype=np.int16)
    step = np.uint64(_step(stream, rate))
    with np.errstate(over="ignore"):
        for c in range(channels):
            phase0 = np.uint64((c * 0x1F3D5B79) & 0xFFFFFFFF)
            phase = (phase0 + n * step) & np.uint64(0xFFFFFFFF)
            tone = (np.int64(_A) * tab[(phase >> np.uint64(32 - _TAB_BITS)).astype(np.int64)].astype(np.int64)) >> 15
            s = np.uint64((seed + stream * 256 + c) & _MASK64)
            r = _splitmix64_np(s * np.uint64(0x2545F4914F6CDD1D) + n)
            noise = ((r >> np.uint64(40)) % np.uint64(2 * _NOISE_AMP + 1)).astype(np.int64) - _NOISE_AMP
            out[:, c] = np.clip(tone + noise, -32768, 32767).astype(np.int16)
    return out.reshape(-1)


def gen_batch(n_streams: int, n_frames: int, channels: int, rate: int, first_stream: int = 0) -> np.ndarray:
    return np.stack([gen_stream(first_stream + k, n_frames, channels, rate) for k in range(n_streams)])


def gen_batch_torch(n_streams: int, n_frames: int, channels: int, rate: int, device, first_stream: int = 0,
                    out=None, streams_per_pass: int = 16):
    """Same samples as gen_batch, computed on `device` with int64 tensor ops.  Returns int16 [n_streams, n_frames*channels]."""
    import torch

    tab = torch.from_numpy(sine_table().astype(np.int64)).to(device)
    if out is None:
        out = torch.empty((n_streams, n_frames * channels), dtype=torch.int16, device=device)
    n = torch.arange(n_frames, dtype=torch.int64, device=device)

    def to_i64(v: int) -> int:  # two's complement view of a u64 constant
        v &= _MASK64
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(x, k):  # logical shift right on int64
        return (x >> k) & ((1 << (64 - k)) - 1)

    c_gamma, c_m1, c_m2 = to_i64(0x9E3779B97F4A7C15), to_i64(0xBF58476D1CE4E5B9), to_i64(0x94D049BB133111EB)
    for k0 in range(0, n_streams, streams_per_pass):
        ks = range(k0, min(n_streams, k0 + streams_per_pass))
        for k in ks:
            stream = first_stream + k
            step = _step(stream, rate)
            view = out[k].view(n_frames, channels)
            for c in range(channels):
                phase0 = (c * 0x1F3D5B79) & 0xFFFFFFFF
                phase = (phase0 + n * step) & 0xFFFFFFFF
                tone = (_A * tab[phase >> (32 - _TAB_BITS)]) >> 15
                s = (0x5EA0000 + stream * 256 + c) & _MASK64
                x = n + to_i64((s * 0x2545F4914F6CDD1D) & _MASK64) + c_gamma
                z = (x ^ lsr(x, 30)) * c_m1
                z = (z ^ lsr(z, 27)) * c_m2
                r = z ^ lsr(z, 31)
                noise = torch.remainder(lsr(r, 40), 2 * _NOISE_AMP + 1) - _NOISE_AMP
                view[:, c] = torch.clamp(tone + noise, -32768, 32767).to(torch.int16)
    return out
