"""seaconv -- converts between .wav and .sea (the reference's CLI, examples/seaconv.rs, over the GPU path).

    python -m sea_codec_b200.seaconv INPUT OUTPUT [-c CHUNK] [-b BITRATE] [-s SF_BITS] [-d SF_DISTANCE] [-v]

Same positional arguments, flags, defaults, validation ranges and messages as seaconv.rs:12-144; the conversion drives the same
SeaEncoder / SeaDecoder state machines (seaconv.rs:155-216), several chunks per launch.  Needs a CUDA device (no CPU path).
"""
from __future__ import annotations

import argparse
import io
import os
import sys

import numpy as np


def _die(msg: str) -> "NoReturn":  # seaconv.rs: eprintln! + exit(1)
    sys.stderr.write(f"Error: {msg}\n")
    raise SystemExit(1)


def get_encoder_settings(a):
    """seaconv.rs:12-91."""
    from . import api

    try:
        fpc = int(a.chunk_size)
        if not 0 <= fpc <= 0xFFFF:
            raise ValueError
    except ValueError:
        _die("Failed to parse chunk size")
    if fpc < 200 or fpc > 32000:
        _die("Chunk size must be between 200 and 32000")
    try:
        sfb = int(a.scalefactor_bits)
        if not 0 <= sfb <= 255:
            raise ValueError
    except ValueError:
        _die("Failed to parse scale factor bits")
    if sfb < 3 or sfb > 5:
        _die("Scale factor bits must be between 3 and 5")
    try:
        sff = int(a.scalefactor_distance)
        if not 0 <= sff <= 255:
            raise ValueError
    except ValueError:
        _die("Failed to parse scale factor frames")
    if sff < 1 or fpc % sff != 0:
        _die("Scale factor frames must be a divisor of chunk size")
    try:
        bits = float(np.float32(a.bitrate))
    except ValueError:
        _die("Failed to parse residual bits")
    if bits < 1.0 or bits > 8.0 or bits != bits:
        _die("Bitrate must be between 1.0 and 8.0")
    if a.vbr:
        if not (1.5 <= bits <= 8.0):
            _die("With VBR, bitrate must be between 1.5 and 8.0")
    elif bits != int(bits) or not (1 <= int(bits) <= 8):
        _die("Without VBR, bitrate must be an integer between 1 and 8")
    return api.EncoderSettings(scale_factor_bits=sfb, scale_factor_frames=sff, residual_bits=bits, frames_per_chunk=fpc, vbr=bool(a.vbr))


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="seaconv", description="Converts between .wav and .sea files")
    ap.add_argument("input", help="The input file in LPCM LE .wav or .sea format")
    ap.add_argument("output", help="The output file to save the conversion result (.sea or .wav)")
    ap.add_argument("-c", "--chunk-size", default="5120", help="Sets the number of frames within a chunk")
    ap.add_argument("-b", "--bitrate", default="3", help="Sets the bitrate for the conversion")
    ap.add_argument("-s", "--scalefactor-bits", default="4", help="Sets the bitrate for scale factors")
    ap.add_argument("-d", "--scalefactor-distance", default="20", help="Sets the distance between scale factors in frames")
    ap.add_argument("-v", "--vbr", action="store_true", help="Enables Variable Bit Rate (VBR)")
    ap.add_argument("--chunks-per-launch", type=int, default=256, help="(addition) chunks handed to the GPU per call")
    return ap


def main(argv=None) -> int:
    a = build_parser().parse_args(argv)
    settings = get_encoder_settings(a)
    from . import api, wav

    ext_in = os.path.splitext(a.input)[1].lstrip(".")
    ext_out = os.path.splitext(a.output)[1].lstrip(".")
    per = max(1, a.chunks_per_launch)
    if (ext_in, ext_out) == ("wav", "sea"):
        try:
            w = wav.read_wav(a.input)
        except Exception:
            _die("Failed to decode .wav file")
        try:
            out = open(a.output, "wb")
        except OSError:
            _die("Failed to create output file")
        with out:
            try:
                enc = api.SeaEncoder(w.channels, w.sample_rate, w.samples.size // w.channels, settings,
                                     io.BytesIO(w.samples.astype("<i2").tobytes()), out)
            except api.SeaError:
                _die("Failed to create encoder")
            try:
                while enc.encode_frames(per):
                    pass
            except api.SeaError:
                _die("Failed to encode frame")
            enc.finalize()
            enc.close()
    elif (ext_in, ext_out) == ("sea", "wav"):
        try:
            f = open(a.input, "rb")
        except OSError:
            _die("Failed to open input file")
        with f:
            pcm = io.BytesIO()
            try:
                dec = api.SeaDecoder(f, pcm)
                while dec.decode_frames(per):
                    pass
            except api.SeaError:
                _die("Failed to decode frame")
            dec.finalize()
            info = dec.get_header()
            dec.close()
        try:
            wav.write_wav(np.frombuffer(pcm.getvalue(), dtype="<i2"), info.channels, info.sample_rate, a.output)
        except OSError:
            _die("Failed to encode wav file")
    else:
        _die("Invalid file extensions. Supported conversions are .wav to .sea and .sea to .wav")
    return 0


if __name__ == "__main__":
    sys.exit(main())
