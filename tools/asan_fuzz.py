"""Driver of tools/asan_fuzz.sh: slice-data / header corruptions of the generated streams through both parsers, and
container corruptions through the reader, against the sanitizer build named by HEIFCUDA_ASAN_LIB."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heif-decoder-lib_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import heif_b200._lib as _l  # noqa: E402

_l.lib_path = lambda host_only=False: os.environ["HEIFCUDA_ASAN_LIB"]
import numpy as np  # noqa: E402
import heif_b200 as hb  # noqa: E402
import test_fuzz as T  # noqa: E402


def corrupt_any(data, seed):
    """anywhere in the stream, parameter sets and slice headers included"""
    rng = np.random.default_rng(seed)
    d = bytearray(data)
    for _ in range(int(rng.integers(1, 8))):
        i = int(rng.integers(0, len(d)))
        for k in range(i, min(len(d), i + int(rng.integers(1, 8)))):
            d[k] = int(rng.integers(0, 256))
    if rng.integers(0, 4) == 0:
        d = d[:int(rng.integers(4, len(d)))]
    return bytes(d)


def main():
    first, last = int(sys.argv[1]), int(sys.argv[2])
    ok = rejected = 0
    for name in T._victims() + ["tiles_2x2", "pcm", "bypass", "c444_8", "mono_8"]:
        path = os.path.join(T.GEN_DIR, name + ".hevc")
        if not os.path.exists(path):
            continue
        data = open(path, "rb").read()
        for seed in range(first, last):
            for bad in (T.corrupt(data, seed * 7919 + len(name)), corrupt_any(data, seed * 31 + 7)):
                for fn in (hb.parse_picture, hb.parse_picture_k0):
                    try:
                        fn(bad, T._fmt(bad), host_only=True)
                        ok += 1
                    except hb.HeifCudaError:
                        rejected += 1
    print("streams: parsed", ok, "rejected", rejected)
    sys.argv = [sys.argv[0], os.path.join(ROOT, "heif-decoder-lib_b200"), os.path.join(ROOT, "tests"), ROOT]
    exec(T._CONTAINER_CHILD.replace("import heif_b200 as hb", "pass"), {"hb": hb, "__name__": "fuzz"})


if __name__ == "__main__":
    main()
