"""Regenerates tests/golden/*.sea and golden.json with the CPU oracle (and cross-checks CBR decodes with the reference's
own C decoder c/sea.h from oracle/_ref).  Run from the repo root in the build container: python tests/golden/make_golden.py

The reference holds no golden vectors of its own; these fixtures pin (a) the oracle against regressions and
(b) the CUDA path on the GPU box, where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import sea_oracle as O  # noqa: E402
from sea_codec_b200 import synth  # noqa: E402
from util import gen_test_signal, sha  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, generator, stream/seed, frames, channels, rate, settings kwargs
    ("stereo_cbr3", "synth", 0, 12000, 2, 44100, dict(residual_bits=3.0)),
    ("mono_vbr3", "synth", 1, 11111, 1, 48000, dict(residual_bits=3.0, vbr=True)),
    ("ch3_cbr5_sf5", "helpers", 3, 7000, 3, 44100, dict(residual_bits=5.0, scale_factor_bits=5)),
    ("ch8_cbr4", "synth", 2, 6000, 8, 48000, dict(residual_bits=4.0)),
    ("stereo_vbr45", "synth", 4, 10300, 2, 44100, dict(residual_bits=4.5, vbr=True)),
    ("mono_cbr1", "synth", 5, 3000, 1, 44100, dict(residual_bits=1.0)),
    ("mono_cbr8", "synth", 6, 3000, 1, 44100, dict(residual_bits=8.0)),
    ("stereo_cbr2_sf3_f10", "helpers", 7, 5000, 2, 44100, dict(residual_bits=2.0, scale_factor_bits=3, scale_factor_frames=10, frames_per_chunk=1000)),
]


def make_pcm(gen, seed, frames, channels, rate):
    if gen == "synth":
        return synth.gen_stream(seed, frames, channels, rate)
    return gen_test_signal(channels, frames, rate, seed)


def main():
    meta = {}
    for name, gen, seed, frames, ch, rate, kw in CASES:
        pcm = make_pcm(gen, seed, frames, ch, rate)
        st = O.make_settings(**kw)
        enc, ties = O.sea_encode(pcm, rate, ch, st, return_ties=True)
        dec = O.sea_decode(enc).samples
        entry = dict(gen=gen, seed=seed, frames=frames, channels=ch, rate=rate, settings=kw, pcm_sha=sha(pcm), sea_sha=sha(enc),
                     sea_len=len(enc), dec_sha=sha(dec), vbr_ties=ties, cref_checked=False)
        if not kw.get("vbr") and frames % kw.get("scale_factor_frames", 20) == 0 and O.have_ref():
            cref = O.ref_c_decode(enc).samples
            assert np.array_equal(cref, dec), name
            entry["cref_checked"] = True
        assert ties == 0, (name, ties)
        with open(os.path.join(HERE, name + ".sea"), "wb") as f:
            f.write(enc)
        meta[name] = entry
        print(name, len(enc), entry["cref_checked"])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
