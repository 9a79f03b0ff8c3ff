"""Tuning build of libsea_b200 with extra -D flags:  python tools/build_variant.py NAME -DSEA_DEC_WARPS=16 ...
Writes sea_codec_b200/variants/libsea_b200_NAME.so; select it with SEA_B200_LIB=<path> (api.lib())."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sea_codec_b200 import build as B

name, flags = sys.argv[1], sys.argv[2:]
vdir = os.path.join(B.HERE, "variants")
os.makedirs(vdir, exist_ok=True)
print(B.build(force=True, extra_flags=flags, lib_path=os.path.join(vdir, f"libsea_b200_{name}.so"), build_dir=os.path.join(vdir, "build_" + name)))
