"""ctypes front-end of the CPU oracle (oracle/sea_oracle.c) and of the reference's C decoder (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (sea_codec_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsea_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_LIB_PATH = os.path.join(REF_DIR, "libsea_cref.so")
REF_BENCH_PATH = os.path.join(REF_DIR, "csea_bench")

ERR_PANIC = -100  # the reference would panic on this input
ERR_CAPACITY = -101


class OracleSettings(C.Structure):
    """Mirror of EncoderSettings (src/encoder.rs:16-35)."""

    _fields_ = [
        ("scale_factor_bits", C.c_uint8),
        ("scale_factor_frames", C.c_uint8),
        ("residual_bits", C.c_float),
        ("frames_per_chunk", C.c_uint16),
        ("vbr", C.c_uint8),
    ]


def make_settings(residual_bits=3.0, vbr=False, scale_factor_bits=4, scale_factor_frames=20, frames_per_chunk=5120):
    return OracleSettings(scale_factor_bits, scale_factor_frames, float(residual_bits), frames_per_chunk, 1 if vbr else 0)


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present) with oracle/Makefile."""
    src = os.path.join(HERE, "sea_oracle.c")
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    if force or stale or (os.path.exists("/root/reference/c/sea.h") and not os.path.exists(REF_LIB_PATH)):
        subprocess.run(["make", "-C", HERE, "-s"], check=True, stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_sea_encode.restype = C.c_int64
        L.oracle_sea_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.POINTER(OracleSettings),
                                        C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.oracle_sea_decode.restype = C.c_int
        L.oracle_sea_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_encoder_new.restype = C.c_void_p
        L.oracle_encoder_new.argtypes = [C.c_uint8, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(OracleSettings),
                                         C.c_void_p, C.POINTER(C.c_size_t)]
        L.oracle_encoder_free.argtypes = [C.c_void_p]
        L.oracle_encoder_encode_frame.restype = C.c_int
        L.oracle_encoder_encode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                                  C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.oracle_scale_factors.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.oracle_tables.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_quant_tab.argtypes = [C.c_int, C.c_void_p]
        L.oracle_vbr_params.argtypes = [C.POINTER(OracleSettings), C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_int),
                                        C.POINTER(C.c_size_t * 4)]
        L.oracle_sea_div.restype = C.c_int32
        L.oracle_sea_div.argtypes = [C.c_int32, C.c_int32]
        L.oracle_bench.restype = C.c_double
        L.oracle_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                   C.POINTER(OracleSettings), C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def have_ref() -> bool:
    try:
        build()
    except Exception:
        pass
    return os.path.exists(REF_LIB_PATH)


def ref():
    """The reference's own C decoder (c/sea.h) compiled into oracle/_ref (CBR only)."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref is not built (reference tree absent and no prebuilt library)")
        R = C.CDLL(REF_LIB_PATH)
        R.ref_csea_decode.restype = C.c_int
        R.ref_csea_decode.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p,
                                      C.POINTER(C.c_uint32)]
        R.ref_csea_decode_capture_lms.restype = C.c_int
        R.ref_csea_decode_capture_lms.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p,
                                                  C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        _ref = R
    return _ref


class OracleError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"oracle {what} failed with code {code}")
        self.code = code


def sea_encode(samples: np.ndarray, sample_rate: int, channels: int, settings: OracleSettings, return_ties=False):
    """lib.rs:13-36 restated.  samples: interleaved int16."""
    s = np.ascontiguousarray(samples, dtype=np.int16).reshape(-1)
    cap = s.size * 2 + 65536 + 64
    out = np.empty(cap, dtype=np.uint8)
    ties = C.c_uint64(0)
    n = lib().oracle_sea_encode(s.ctypes.data, s.size, sample_rate, channels, C.byref(settings), out.ctypes.data, cap,
                                C.byref(ties))
    if n < 0:
        raise OracleError(int(n), "sea_encode")
    enc = out[:n].tobytes()
    return (enc, int(ties.value)) if return_ties else enc


@dataclass
class DecodeInfo:
    samples: np.ndarray
    sample_rate: int
    channels: int


def sea_decode(encoded: bytes, max_samples: int | None = None) -> DecodeInfo:
    """lib.rs:44-63 restated."""
    buf = np.frombuffer(encoded, dtype=np.uint8)
    if max_samples is None:
        max_samples = max(len(encoded) * 16, 1 << 16)
        if len(encoded) >= 22:
            ch = encoded[5]
            fpc = encoded[8] | (encoded[9] << 8)
            cs = encoded[6] | (encoded[7] << 8)
            if cs:
                max_samples = (len(encoded) // cs + 2) * fpc * max(ch, 1)
    out = np.empty(max_samples + 64, dtype=np.int16)
    n = C.c_size_t(0)
    rate = C.c_uint32(0)
    ch = C.c_uint32(0)
    rc = lib().oracle_sea_decode(buf.ctypes.data, buf.size, out.ctypes.data, max_samples, C.byref(n), C.byref(rate), C.byref(ch))
    if rc != 0:
        raise OracleError(rc, "sea_decode")
    return DecodeInfo(out[: n.value].copy(), rate.value, ch.value)


def ref_c_decode(encoded: bytes) -> DecodeInfo:
    """Decode with the reference's c/sea.h (CBR only; frames must be a multiple of scale_factor_frames)."""
    R = ref()
    buf = np.frombuffer(encoded, dtype=np.uint8).copy()
    rate, ch, frames = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    rc = R.ref_csea_decode(buf.ctypes.data, buf.size, C.byref(rate), C.byref(ch), None, C.byref(frames))
    if rc:
        raise OracleError(rc, "ref c/sea.h header")
    fpc = encoded[8] | (encoded[9] << 8)
    out = np.zeros(frames.value * ch.value + fpc * ch.value + 4096, dtype=np.int16)  # slack: c/sea.h:168 over-run
    rc = R.ref_csea_decode(buf.ctypes.data, buf.size, C.byref(rate), C.byref(ch), out.ctypes.data, C.byref(frames))
    if rc:
        raise OracleError(rc, "ref c/sea.h decode")
    return DecodeInfo(out[: frames.value * ch.value].copy(), rate.value, ch.value)


def ref_c_decode_lms(encoded: bytes):
    """c/sea.h decode that also returns the LMS state the REFERENCE decoder held at the end of every chunk
    (oracle/ref_csea.c capture hook): (DecodeInfo, lms[chunk][channel][8] int32 = history[4] then weights[4])."""
    R = ref()
    buf = np.frombuffer(encoded, dtype=np.uint8).copy()
    rate, ch, frames = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    rc = R.ref_csea_decode(buf.ctypes.data, buf.size, C.byref(rate), C.byref(ch), None, C.byref(frames))
    if rc:
        raise OracleError(rc, "ref c/sea.h header")
    fpc = encoded[8] | (encoded[9] << 8)
    n_chunks = (frames.value + fpc - 1) // fpc
    out = np.zeros(frames.value * ch.value + fpc * ch.value + 4096, dtype=np.int16)
    lms = np.zeros((max(n_chunks, 1), ch.value, 8), dtype=np.int32)
    got = C.c_uint32(0)
    rc = R.ref_csea_decode_capture_lms(buf.ctypes.data, buf.size, C.byref(rate), C.byref(ch), out.ctypes.data, C.byref(frames),
                                       lms.ctypes.data, n_chunks, C.byref(got))
    if rc:
        raise OracleError(rc, "ref c/sea.h decode")
    assert got.value == n_chunks, (got.value, n_chunks)
    return DecodeInfo(out[: frames.value * ch.value].copy(), rate.value, ch.value), lms[:n_chunks]


def chunk_header_lms(encoded: bytes) -> np.ndarray:
    """The LMS block of every chunk header as written (lms.rs:64-78: 4 x i16 history, 4 x i16 weights per channel, LE):
    int16 array [chunk][channel][8].  Full chunks sit chunk_size apart from byte 22 (file.rs:185)."""
    ch = encoded[5]
    cs = encoded[6] | (encoded[7] << 8)
    n = (len(encoded) - 22 + cs - 1) // cs
    out = np.zeros((n, ch, 8), dtype=np.int16)
    for k in range(n):
        base = 22 + k * cs + 4
        out[k] = np.frombuffer(encoded, dtype="<i2", count=ch * 8, offset=base).reshape(ch, 8)
    return out


class StreamingEncoder:
    """encoder.rs:50-159 restated (SeaEncoder over an in-memory reader)."""

    def __init__(self, channels, sample_rate, total_frames, settings: OracleSettings):
        hdr = np.zeros(64, dtype=np.uint8)
        n = C.c_size_t(0)
        self._settings = settings
        self._h = lib().oracle_encoder_new(channels, sample_rate, 0 if total_frames is None else 1,
                                           0 if total_frames is None else total_frames, C.byref(settings),
                                           hdr.ctypes.data, C.byref(n))
        self.initial_bytes = hdr[: n.value].tobytes()
        self.channels = channels

    def encode_frame(self, samples: np.ndarray):
        """Feed the reader's remaining samples; returns (more, bytes_written, samples_consumed)."""
        s = np.ascontiguousarray(samples, dtype=np.int16).reshape(-1)
        cap = 70000 + 64
        out = np.empty(cap, dtype=np.uint8)
        n, used = C.c_size_t(0), C.c_size_t(0)
        rc = lib().oracle_encoder_encode_frame(self._h, s.ctypes.data, s.size, out.ctypes.data, cap, C.byref(n), C.byref(used))
        if rc < 0:
            raise OracleError(rc, "encode_frame")
        return bool(rc), out[: n.value].tobytes(), used.value

    def close(self):
        if self._h:
            lib().oracle_encoder_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def scale_factors(residual_bits: int, scale_factor_bits: int) -> np.ndarray:
    out = np.zeros(1 << scale_factor_bits, dtype=np.int32)
    lib().oracle_scale_factors(residual_bits, scale_factor_bits, out.ctypes.data)
    return out


def tables(residual_bits: int, scale_factor_bits: int):
    n = 1 << scale_factor_bits
    recip = np.zeros(n, dtype=np.int32)
    dqt = np.zeros((n, 1 << residual_bits), dtype=np.int32)
    lib().oracle_tables(residual_bits, scale_factor_bits, recip.ctypes.data, dqt.ctypes.data)
    return recip, dqt


def quant_tab(residual_bits: int) -> np.ndarray:
    out = np.zeros((1 << (residual_bits + 1)) + 1, dtype=np.uint8)
    lib().oracle_quant_tab(residual_bits, out.ctypes.data)
    return out


def vbr_params(settings: OracleSettings, items: int):
    target, base = C.c_float(0), C.c_int(0)
    counts = (C.c_size_t * 4)()
    lib().oracle_vbr_params(C.byref(settings), items, C.byref(target), C.byref(base), C.byref(counts))
    return target.value, base.value, list(counts)


def sea_div(v: int, recip: int) -> int:
    return lib().oracle_sea_div(v, recip)


def bench(mode: str, threads: int, reps: int, pcm: np.ndarray, sample_rate: int, channels: int, settings: OracleSettings,
          sea: bytes | None = None):
    """Times `reps` encodes/decodes of one stream on each of `threads` host threads; returns (seconds, samples)."""
    s = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
    done = C.c_uint64(0)
    if mode == "encode":
        secs = lib().oracle_bench(0, threads, reps, s.ctypes.data, s.size, sample_rate, channels, C.byref(settings), None, 0,
                                  C.byref(done))
    else:
        buf = np.frombuffer(sea, dtype=np.uint8)
        secs = lib().oracle_bench(1, threads, reps, None, s.size, sample_rate, channels, C.byref(settings), buf.ctypes.data,
                                  buf.size, C.byref(done))
    return secs, int(done.value)


def ref_c_bench(sea_path: str, procs: int, reps: int):
    """Times the reference's c/sea.h decoder on `procs` forked workers; returns (seconds, samples)."""
    out = subprocess.run([REF_BENCH_PATH, sea_path, str(procs), str(reps)], check=True, capture_output=True, text=True).stdout
    secs, samples = out.split()
    return float(secs), int(samples)
