"""seaconv CLI + WAV plumbing: argument validation and .wav conversions mirror examples/seaconv.rs and
tests/wav.rs; runs on CPU (the conversions themselves are exercised on the GPU in test_gpu_next_rows.py)."""
import struct

import numpy as np
import pytest

from sea_codec_b200 import seaconv, wav


def _wav_bytes(tag, bits, channels, rate, payload, extensible=False):
    block = channels * bits // 8
    if extensible:
        fmt = struct.pack("<HHIIHHHHIH", 0xFFFE, channels, rate, rate * block, block, bits, 22, bits, 3, tag) + b"\x00" * 14
    else:
        fmt = struct.pack("<HHIIHH", tag, channels, rate, rate * block, block, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 3) + b"abc\x00" + \
        b"data" + struct.pack("<I", len(payload)) + payload
    return b"RIFF" + struct.pack("<I", len(body)) + body


def test_wav_round_trip_16(tmp_path):
    x = (np.arange(-500, 500) * 60).astype(np.int16)
    p = str(tmp_path / "a.wav")
    wav.write_wav(x, 2, 44100, p)
    w = wav.read_wav(p)
    assert w.channels == 2 and w.sample_rate == 44100 and np.array_equal(w.samples, x)
    raw = open(p, "rb").read()
    assert raw[:4] == b"RIFF" and struct.unpack_from("<I", raw, 40)[0] == x.size * 2 and len(raw) == 44 + x.size * 2


def test_wav_conversions_follow_the_reference_helper(tmp_path):
    # 8 bit: hound yields i8 = u8 - 128, then (s as i16) << 8          (tests/wav.rs:21-24)
    p = tmp_path / "u8.wav"
    p.write_bytes(_wav_bytes(1, 8, 1, 8000, bytes([0, 1, 127, 128, 129, 255])))
    assert wav.read_wav(str(p)).samples.tolist() == [-32768, -32512, -256, 0, 256, 32512]
    # 24 bit: ((s as f32 / 2^23) * 32767).round()                      (tests/wav.rs:26-29)
    vals = [0, 1, -1, 4194304, -4194304, 8388607, -8388608, 256, 128]
    payload = b"".join(struct.pack("<i", v)[:3] for v in vals)
    p = tmp_path / "s24.wav"
    p.write_bytes(_wav_bytes(1, 24, 1, 8000, payload, extensible=True))
    want = [int(np.copysign(np.floor(abs(np.float32(np.float32(v) / np.float32(8388608)) * np.float32(32767)) + np.float32(0.5)),
                            v)) for v in vals]
    assert wav.read_wav(str(p)).samples.tolist() == want
    assert want[5] == 32767 and want[6] == -32767 and want[3] == 16384 and want[7] == 1
    # 32 bit int: (s as f32 / i32::MAX as f32) * 32767                 (tests/wav.rs:30-33)
    vals = [0, 2147483647, -2147483648, 65536, -65536, 32768, 1 << 30]
    p = tmp_path / "s32.wav"
    p.write_bytes(_wav_bytes(1, 32, 2, 8000, struct.pack("<8i", *vals, 0)))
    got = wav.read_wav(str(p)).samples.tolist()
    assert got[:7] == [0, 32767, -32767, 1, -1, 0, 16384], got  # 32768/2^31*32767 = 0.49998 -> 0; 2^30 -> 16383.5 -> 16384
    # float: (s * 32767).round() as i16, saturating, NaN -> 0          (tests/wav.rs:34-37)
    vals = [0.0, 1.0, -1.0, 0.5, 2.0, -2.0, float("nan"), 1.5259e-5]
    p = tmp_path / "f32.wav"
    p.write_bytes(_wav_bytes(3, 32, 1, 8000, struct.pack("<8f", *vals)))
    assert wav.read_wav(str(p)).samples.tolist() == [0, 32767, -32767, 16384, 32767, -32768, 0, 0]


def test_wav_rejects_what_the_reference_rejects(tmp_path):
    p = tmp_path / "c3.wav"
    p.write_bytes(_wav_bytes(1, 16, 3, 8000, b"\x00" * 12))
    with pytest.raises(wav.WavError, match="More than 2 channels"):
        wav.read_wav(str(p))
    p = tmp_path / "f64.wav"
    p.write_bytes(_wav_bytes(3, 64, 1, 8000, b"\x00" * 16))
    with pytest.raises(wav.WavError, match="Unsupported format"):
        wav.read_wav(str(p))
    p = tmp_path / "junk.wav"
    p.write_bytes(b"not a wave file at all")
    with pytest.raises(wav.WavError):
        wav.read_wav(str(p))


@pytest.mark.parametrize("argv,msg", [
    (["a.wav", "b.sea", "-c", "199"], "Chunk size must be between 200 and 32000"),
    (["a.wav", "b.sea", "-c", "32001"], "Chunk size must be between 200 and 32000"),
    (["a.wav", "b.sea", "-c", "x"], "Failed to parse chunk size"),
    (["a.wav", "b.sea", "-s", "2"], "Scale factor bits must be between 3 and 5"),
    (["a.wav", "b.sea", "-s", "6"], "Scale factor bits must be between 3 and 5"),
    (["a.wav", "b.sea", "-d", "0"], "Scale factor frames must be a divisor of chunk size"),
    (["a.wav", "b.sea", "-d", "7"], "Scale factor frames must be a divisor of chunk size"),
    (["a.wav", "b.sea", "-b", "0.5"], "Bitrate must be between 1.0 and 8.0"),
    (["a.wav", "b.sea", "-b", "8.5"], "Bitrate must be between 1.0 and 8.0"),
    (["a.wav", "b.sea", "-b", "2.5"], "Without VBR, bitrate must be an integer between 1 and 8"),
    (["a.wav", "b.sea", "-b", "1.2", "-v"], "With VBR, bitrate must be between 1.5 and 8.0"),
    (["a.wav", "b.sea", "-b", "abc"], "Failed to parse residual bits"),
    (["a.wav", "b.mp3"], "Invalid file extensions"),
    (["a.sea", "b.sea"], "Invalid file extensions"),
])
def test_seaconv_validation_matches_seaconv_rs(argv, msg, capsys):
    with pytest.raises(SystemExit) as e:
        seaconv.main(argv)
    assert e.value.code == 1
    assert msg in capsys.readouterr().err


def test_seaconv_settings_defaults():
    a = seaconv.build_parser().parse_args(["in.wav", "out.sea"])
    s = seaconv.get_encoder_settings(a)
    assert (s.frames_per_chunk, s.scale_factor_bits, s.scale_factor_frames, s.residual_bits, s.vbr) == (5120, 4, 20, 3.0, False)
    a = seaconv.build_parser().parse_args(["in.wav", "out.sea", "-c", "1000", "-b", "4.5", "-s", "5", "-d", "10", "-v"])
    s = seaconv.get_encoder_settings(a)
    assert (s.frames_per_chunk, s.scale_factor_bits, s.scale_factor_frames, s.residual_bits, s.vbr) == (1000, 5, 10, 4.5, True)
