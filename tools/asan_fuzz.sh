#!/bin/sh
# Builds the host library (container reader, parameter sets, host CABAC parser, CPU form of the K0 parser, job planning)
# with AddressSanitizer + UBSan and runs the fuzz inputs of tests/test_fuzz.py against it. Found in round 2: the CPU form of
# K0's bitstream refill read up to 4 bytes behind a damaged substream (fixed in k0_core.cuh, Cabac::next32 / cab_next32).
#   tools/asan_fuzz.sh [first_seed last_seed]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/heifcuda_asan
mkdir -p "$OUT"
cd "$ROOT/heif-decoder-lib_b200/csrc"
g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -shared \
    -o "$OUT/libheifcuda_host.so" host/hevc_params.cc host/hevc_parse.cc host/k0_host.cc host/heif_reader.cc host/csc_select.cc \
    capi/capi_host.cc engine/heic_job.cc capi/no_engine.cc
cd "$ROOT"
LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
    HEIFCUDA_ASAN_LIB="$OUT/libheifcuda_host.so" python tools/asan_fuzz.py "${1:-0}" "${2:-40}"
