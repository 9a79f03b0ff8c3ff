"""Shared helpers of the test-suite."""
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(b) -> str:
    return hashlib.sha256(np.ascontiguousarray(b).tobytes() if isinstance(b, np.ndarray) else b).hexdigest()


def gen_test_signal(channels: int, frames: int, rate: int = 44100, seed: int = 1) -> np.ndarray:
    """In the spirit of tests/helpers.rs:79-93 (square + sine mix, per-channel delay), but integer-exact: built from
    numpy float64 and rounded once, so every platform produces the same int16 samples."""
    n = np.arange(frames, dtype=np.float64)
    x = np.zeros(frames)

    def seg(a, b):
        return slice(int(frames * a), int(frames * b))

    def square(sl, gain, f):
        period = max(int(rate / f), 2)
        idx = np.arange(sl.stop - sl.start)
        x[sl] += gain * np.where((idx % period) < period // 2, 1.0, -1.0)

    def sine(sl, gain, f):
        idx = np.arange(sl.stop - sl.start)
        x[sl] += gain * np.sin(2 * np.pi * f / rate * idx)

    square(seg(0.0, 0.3), 0.5, 440.0)
    square(seg(0.1, 0.2), 0.3, 2150.1)
    sine(seg(0.1, 0.7), 0.5, 105.0)
    square(seg(0.6, 0.7), 0.5, 14000.0)
    sine(seg(0.5, 0.8), 0.8, 12000.0)
    sine(seg(0.8, 0.9), 1.0, 440.0)
    rng = np.random.default_rng(seed)
    x += rng.uniform(-0.01, 0.01, frames)
    delay = max(rate // 250, 1)
    out = np.zeros((frames, channels))
    for c in range(channels):
        d = min(delay * c, frames)
        out[d:, c] = x[: frames - d]
    return (np.clip(out, -1.0, 1.0) * 32767.0).astype(np.int16).reshape(-1)


def build_host_math():
    src = os.path.join(ROOT, "tests", "csrc", "host_math_check.cpp")
    out_dir = os.path.join(ROOT, "tests", "build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhost_math_check.so")
    dep = os.path.join(ROOT, "sea_codec_b200", "csrc", "sea_common.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-fwrapv", "-o", so, src], check=True)
    return so
