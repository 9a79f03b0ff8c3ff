"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol include/sea_b200.h declares,
host-only helpers agree with the oracle, the kernels' closed forms agree with the reference tables, and nothing works
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sea_codec_b200 as S
from sea_codec_b200 import api
from util import ROOT, build_host_math


def test_library_exports_every_declared_symbol():
    L = S.lib()
    hdr = open(os.path.join(ROOT, "include", "sea_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sea_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"libsea_b200.so does not export {name}"
    assert declared == set(api.SYMBOLS), declared ^ set(api.SYMBOLS)
    assert L.sea_b200_abi_version() == 1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(S.SeaError) as e:
        S.Context(0)
    assert e.value.code == api.ERR_CUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sea_codec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower().replace("no cpu", ""), f"{f} mentions the oracle"


def test_tables_match_oracle(oracle):
    L = S.lib()
    for s in range(1, 9):
        for b in range(1, 9):
            n = 1 << s
            r = np.zeros(n, np.int32)
            d = np.zeros((n, 1 << b), np.int32)
            assert L.sea_b200_tables(b, s, r.ctypes.data, d.ctypes.data) == 0
            r2, d2 = oracle.tables(b, s)
            assert np.array_equal(r, r2) and np.array_equal(d, d2)


def test_vbr_plan_and_chunk_bytes_match_oracle(oracle):
    L = S.lib()
    for bits in (1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0, 6.0, 7.0, 7.3):
        for items in (1, 3, 7, 255, 256, 512, 2048, 640):
            st = S.EncoderSettings(residual_bits=bits, vbr=True)._c()
            t, base = C.c_float(0), C.c_uint32(0)
            counts = (C.c_uint64 * 4)()
            assert L.sea_b200_vbr_plan(C.byref(st), items, C.byref(t), C.byref(base), C.byref(counts)) == 0
            t2, b2, c2 = oracle.vbr_params(oracle.make_settings(bits, True), items)
            assert (t.value, base.value, list(counts)) == (t2, b2, c2)
    b = C.c_uint32(0)
    for bits, size in zip(range(1, 9), (1572, 2852, 4132, 5412, 6692, 7972, 9252, 10532)):
        st = S.EncoderSettings(residual_bits=float(bits))._c()
        assert L.sea_b200_full_chunk_bytes(2, C.byref(st), C.byref(b)) == 0 and b.value == size
    st = S.EncoderSettings(residual_bits=4.0)._c()
    assert L.sea_b200_full_chunk_bytes(8, C.byref(st), C.byref(b)) == 0 and b.value == 21636
    st = S.EncoderSettings(residual_bits=3.0, vbr=True)._c()
    assert L.sea_b200_full_chunk_bytes(1, C.byref(st), C.byref(b)) == 0 and b.value == 1922


def test_settings_validation():
    L = S.lib()
    b = C.c_uint32(0)

    def rc(**kw):
        ch = kw.pop("channels", 2)
        st = S.EncoderSettings(**kw)._c()
        return L.sea_b200_full_chunk_bytes(ch, C.byref(st), C.byref(b))

    assert rc() == 0
    assert rc(residual_bits=8.0, vbr=True) == api.ERR_DOMAIN  # size 9: common.rs:34 panics (trap T20)
    assert rc(residual_bits=1.2, vbr=True) == api.ERR_DOMAIN  # base 0
    assert rc(residual_bits=0.5) == api.ERR_DOMAIN
    assert rc(residual_bits=9.0) == api.ERR_DOMAIN
    assert rc(scale_factor_frames=7) == api.ERR_DOMAIN  # chunk.rs:218 assert
    assert rc(scale_factor_bits=0) == api.ERR_INVALID_PARAMETERS
    assert rc(channels=0) == api.ERR_INVALID_PARAMETERS
    assert rc(channels=200, residual_bits=8.0) == api.ERR_DOMAIN  # chunk > 65535 bytes: header.chunk_size is u16


def test_header_parse():
    hdr = api._serialize_header(2, 4132, 5120, 44100, 441000)
    h = S.parse_header(hdr)
    assert (h.version, h.channels, h.chunk_size, h.frames_per_chunk, h.sample_rate, h.total_frames) == (1, 2, 4132, 5120, 44100, 441000)
    with pytest.raises(S.SeaError) as e:
        S.parse_header(b"saec" + hdr[4:])
    assert e.value.code == api.ERR_INVALID_FILE
    with pytest.raises(S.SeaError) as e:
        S.parse_header(api._serialize_header(2, 8, 5120, 44100, 1))  # chunk_size < 16 (file.rs:33-38)
    assert e.value.code == api.ERR_INVALID_FILE
    with pytest.raises(S.SeaError):
        S.parse_header(hdr[:10])


def test_quant_closed_form_matches_reference_table(oracle):
    """quant_code() == sea_div + clamp + SeaQuantTab lookup (encoder_base.rs:22-26, :66-72; qt.rs) for every residual
    size, every scale-factor reciprocal of sf_bits 3..5 and a dense sweep of residuals incl. rounding ties."""
    H = C.CDLL(build_host_math())
    H.hm_quant_code.restype = C.c_uint
    H.hm_quant_code.argtypes = [C.c_int, C.c_int, C.c_uint]
    rng = np.random.default_rng(0)
    for b in range(1, 9):
        qt = oracle.quant_tab(b)
        lim = 1 << b
        recips = sorted({int(r) for s in (3, 4, 5) for r in oracle.tables(b, s)[0]})
        for recip in recips:
            rs = set(range(-300, 301)) | {int(x) for x in rng.integers(-70000, 70000, 300)}
            # residuals that land exactly on .5 rounding boundaries
            for n in range(-lim - 2, lim + 3):
                v = (n * 65536 + 32768) // recip
                rs |= {v - 1, v, v + 1, -v}
            rs |= {-2**31, 2**31 - 1, 65535, -65535, 98302, -98303}
            for r in rs:
                scaled = oracle.sea_div(r, recip)
                clamped = max(-lim, min(lim, scaled))
                assert H.hm_quant_code(r, recip, b) == qt[clamped + lim], (b, recip, r)


def test_lms_arithmetic_matches_definition():
    H = C.CDLL(build_host_math())
    H.hm_penalty.restype = C.c_ulonglong
    rng = np.random.default_rng(1)
    for _ in range(2000):
        w = rng.integers(-2**31, 2**31, 4).astype(np.int32)
        h = rng.integers(-32768, 32768, 4).astype(np.int32)
        if rng.random() < 0.5:
            w = (w >> 14).astype(np.int32)
        acc = int(np.sum(w.astype(np.int64) * h.astype(np.int64))) & 0xFFFFFFFF
        acc = acc - (1 << 32) if acc >= (1 << 31) else acc
        assert H.hm_predict(w.ctypes.data, h.ctypes.data) == acc >> 13
        ssum = sum(int(x) * int(x) for x in w)
        ssum = ssum & 0xFFFFFFFFFFFFFFFF
        ssum = ssum - (1 << 64) if ssum >= (1 << 63) else ssum
        pen = max(0, (ssum >> 18) - 0x8FF)
        assert H.hm_penalty(w.ctypes.data) == (pen * pen) & 0xFFFFFFFFFFFFFFFF
        err = int(rng.integers(-65535, 65536))
        rank0 = int(rng.integers(0, 2**62))
        H.hm_rank_step.restype = C.c_ulonglong
        H.hm_rank_step.argtypes = [C.c_ulonglong, C.c_int, C.c_void_p]
        assert H.hm_rank_step(rank0, err, w.ctypes.data) == (rank0 + err * err + pen * pen) & 0xFFFFFFFFFFFFFFFF
        d = int(rng.integers(-30000, 30000))
        y = int(rng.integers(-32768, 32768))
        w2, h2 = w.copy(), h.copy()
        H.hm_update(w2.ctypes.data, h2.ctypes.data, y, d)
        delta = d >> 4
        exp = [(int(w[i]) + (-delta if h[i] < 0 else delta) + 2**31) % 2**32 - 2**31 for i in range(4)]
        assert w2.tolist() == exp and h2.tolist() == [h[1], h[2], h[3], y]


def test_c_headers_compile_as_c_and_cpp(tmp_path):
    """include/sea_b200.h is a plain C header (C99 and C++17), sea_compat.h keeps the reference's C-level names
    (wasm_api.rs:32-111, c/sea.h:189) and resolves them to exported symbols; sea_b200.hpp compiles on its own."""
    import subprocess

    inc = os.path.join(ROOT, "include")
    c_src = tmp_path / "t.c"
    c_src.write_text('#include "sea_compat.h"\n'
                     "int use(uint8_t *e, uint32_t n, int16_t *o) { uint32_t r, c, f; return sea_decode(e, n, &r, &c, o, &f); }\n"
                     "size_t enc(const int16_t *p, size_t n, uint8_t *o, size_t cap) { setup(); return wasm_sea_encode(p, n, 44100, 2, 3.0f, 0, o, cap); }\n")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + inc, "-c", str(c_src), "-o", str(tmp_path / "t.o")], check=True)
    syms = subprocess.run(["nm", "-u", str(tmp_path / "t.o")], capture_output=True, text=True, check=True).stdout
    L = S.lib()
    for name in ("sea_b200_csea_decode", "sea_b200_wasm_sea_encode", "sea_b200_wasm_setup"):
        assert name in syms and hasattr(L, name)
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "sea_b200.hpp"\nint main() { sea::EncoderSettings s; return s.frames_per_chunk == 5120 ? 0 : 1; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I" + inc, str(cpp)], check=True)


def test_direct_quantiser_thresholds_equal_the_closed_form():
    """encode_kernels.cu (kDirect, CBR sizes 1..3) replaces n = (r*recip + 2^15) >> 16, k = min(|n| >> 1, kmax) (size 2:
    |n| >= 3) by comparisons of A = 2|r| - (r < 0) against theta_j = 2*ceil(X/recip) - 1 + [recip divides X],
    X = 65536*m_j - 32768.  Exhaustive over every reciprocal of the sf-bits-4 tables, plus reciprocals that divide X exactly
    (the case where the positive and the negative threshold differ), and every residual in +-70000."""
    L = S.lib()
    recips = set()
    for b in (1, 2, 3):
        rc = np.zeros(16, dtype=np.int32)
        dq = np.zeros(16 << b, dtype=np.int32)
        assert L.sea_b200_tables(b, 4, rc.ctypes.data, dq.ctypes.data) == 0
        recips.update(int(x) for x in rc)
    recips.update([32768, 16384, 4096, 98304 // 3, 229376 // 7, 65536, 10922, 3, 1])  # several divide some X exactly
    r = np.arange(-70000, 70001, dtype=np.int64)
    A = (2 * np.abs(r) - (r < 0)).astype(np.int64)
    for b, ms in ((3, (2, 4, 6)), (2, (3,))):
        kmax = (1 << (b - 1)) - 1
        for rc in sorted(recips):
            n = (r * rc + 32768) >> 16
            an = np.abs(n)
            want = np.minimum(an >> 1, kmax) if b != 2 else (an >= 3).astype(np.int64)
            got = np.zeros_like(r)
            for m in ms:
                X = 65536 * m - 32768
                q, rem = divmod(X, rc)
                theta = 2 * (q + 1) - 1 if rem else 2 * q
                got += (A >= theta)
            assert np.array_equal(got, want), (b, rc)


def test_register_bitonic_network_sorts():
    """The compare-exchange schedule of warp_sort_512 (encode_kernels.cu: element e = r * 32 + lane, distance j < 32 by shuffle,
    j >= 32 between registers, direction from (e & k)) restated on arrays: it must sort any 512 keys ascending, ties included
    (the kernel's keys are unique: (rank << 9) | index)."""
    rng = np.random.default_rng(7)
    for trial in range(6):
        n = [512, 512, 300, 256, 2, 0][trial]
        keys = rng.integers(0, 1 << 20 if trial else 4, size=512).astype(np.uint64)
        v = np.where(np.arange(512) < n, (keys << np.uint64(9)) | np.arange(512, dtype=np.uint64), np.uint64(0xFFFFFFFFFFFFFFFF))
        e = np.arange(512)
        k = 2
        while k <= 512:
            j = k >> 1
            while j > 0:
                other = v[e ^ j]
                up = (e & k) == 0
                lower = (e & j) == 0
                take_min = lower == up
                v = np.where((other < v) == take_min, other, v)
                j >>= 1
            k <<= 1
        assert np.all(v[:-1] <= v[1:])
        if n:
            order = (v[:n] & np.uint64(511)).astype(np.int64)
            want = np.lexsort((np.arange(n), keys[:n]))  # by (key, index): the tie rule of SURVEY trap T13
            assert np.array_equal(order, want)


def test_sum32_penalty_proof_is_sufficient():
    """encode_kernels.cu, kRankSum32: whenever max|w(0)| + (sum|d| >> 4) + 20 <= 32767 over a 20-frame block, the weights
    penalty of lms.rs:53-62 computed with 32-bit wrapping sums and one 32-bit accumulator equals the exact value."""
    rng = np.random.default_rng(11)
    M32 = (1 << 32) - 1
    proved = 0
    for trial in range(3000):
        scale = int(rng.choice([200, 3000, 12000, 30000]))
        w = [int(x) for x in rng.integers(-scale, scale + 1, size=4)]
        dmax = int(rng.choice([50, 2000, 25245]))
        m0 = max(abs(x) for x in w)
        dsum, exact, pen32, ok32 = 0, 0, 0, True
        for f in range(20):
            s = sum(x * x for x in w)
            p = max(0, (s >> 18) - 0x8FF)
            exact += p * p
            s32 = sum((x * x) & M32 for x in w) & M32            # four 32-bit multiply-adds
            t = max(0, ((s32 >> 18) - 0x8FF))                    # VIADDMNMX on the (31-bit) shifted sum
            pen32 = (pen32 + t * t) & M32
            d = int(rng.integers(-dmax, dmax + 1))
            dsum += abs(d)
            delta = d >> 4
            w = [x + delta * int(rng.choice([-1, 1])) for x in w]
        if m0 + (dsum >> 4) + 20 <= 32767:
            proved += 1
            assert pen32 == exact, (trial, m0, dsum)
    assert proved > 300
