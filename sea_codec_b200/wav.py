"""LPCM .wav reading/writing with the conversions of the reference's helper (tests/wav.rs:1-75, built on hound 3.5.1):
8/16/24/32-bit integer and 32-bit float input are mapped to i16 exactly as read_wav does (f32 arithmetic, round half away
from zero, saturating cast); output is 16-bit PCM.  Host-side file plumbing only: no codec arithmetic here."""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np


@dataclass
class Wave:
    samples: np.ndarray  # interleaved int16
    channels: int
    sample_rate: int


class WavError(Exception):
    pass


def _round_half_away(x: np.ndarray) -> np.ndarray:
    """f32::round (half away from zero) on a float32 array, staying in float32."""
    return np.copysign(np.floor(np.abs(x) + np.float32(0.5)), x).astype(np.float32)


def _as_i16_saturating(x: np.ndarray) -> np.ndarray:
    """Rust `as i16` from f32: saturates, NaN -> 0."""
    x = np.nan_to_num(x, nan=0.0, posinf=32767.0, neginf=-32768.0)
    return np.clip(x, -32768.0, 32767.0).astype(np.int16)


def read_wav(path: str) -> Wave:
    data = open(path, "rb").read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise WavError("not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8: pos + 8 + size]
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            pcm = body
            break
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None or len(fmt) < 16:
        raise WavError("missing fmt or data chunk")
    tag, channels, rate, _, block_align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if tag == 0xFFFE and len(fmt) >= 26:  # WAVE_FORMAT_EXTENSIBLE: the sub-format GUID starts with the real tag
        tag = struct.unpack_from("<H", fmt, 24)[0]
    if channels > 2:
        raise WavError("More than 2 channels are not supported")  # tests/wav.rs:15-17
    if channels == 0 or rate == 0:
        raise WavError("bad fmt chunk")
    width = bits // 8
    n = len(pcm) // width if width else 0
    n -= n % channels
    if tag == 1 and bits == 8:  # hound yields i8 = u8 - 128; (s as i16) << 8
        s = (np.frombuffer(pcm, dtype=np.uint8, count=n).astype(np.int16) - 128) << 8
        samples = s.astype(np.int16)
    elif tag == 1 and bits == 16:
        samples = np.frombuffer(pcm, dtype="<i2", count=n).astype(np.int16)
    elif tag == 1 and bits == 24:
        b = np.frombuffer(pcm, dtype=np.uint8, count=n * 3).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        f = (v.astype(np.float32) / np.float32(1 << 23)) * np.float32(32767.0)
        samples = _as_i16_saturating(_round_half_away(f))
    elif tag == 1 and bits == 32:
        v = np.frombuffer(pcm, dtype="<i4", count=n)
        f = (v.astype(np.float32) / np.float32(2147483647)) * np.float32(32767.0)
        samples = _as_i16_saturating(_round_half_away(f))
    elif tag == 3 and bits == 32:
        v = np.frombuffer(pcm, dtype="<f4", count=n).astype(np.float32)
        samples = _as_i16_saturating(_round_half_away(v * np.float32(32767.0)))
    else:
        raise WavError(f"Unsupported format: tag {tag} with {bits} bits")  # tests/wav.rs:39-41
    return Wave(np.ascontiguousarray(samples), channels, rate)


def wav_header(channels: int, sample_rate: int, n_samples: int) -> bytes:
    data = n_samples * 2
    return (b"RIFF" + struct.pack("<I", 36 + data) + b"WAVEfmt " +
            struct.pack("<IHHIIHH", 16, 1, channels, sample_rate, sample_rate * channels * 2, channels * 2, 16) +
            b"data" + struct.pack("<I", data))


def write_wav(samples: np.ndarray, channels: int, sample_rate: int, path: str) -> None:
    """tests/wav.rs:52-75: 16-bit integer PCM."""
    s = np.ascontiguousarray(samples, dtype="<i2").reshape(-1)
    with open(path, "wb") as f:
        f.write(wav_header(channels, sample_rate, s.size))
        f.write(s.tobytes())
