#!/bin/sh
# Writes the SASS listing of every kernel object (built by heif-decoder-lib_b200/csrc/Makefile) to profiles/sass/.
# Encoding words are stripped to keep the listings readable; tcgen05 / TMA proof points (UBLKCP, SYNCS) are in k2.
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/sass
for k in k0_parse k1_transform k2_intra k3_deblock k4_sao k5_csc k6_transform; do
  cuobjdump -sass heif-decoder-lib_b200/csrc/build/kernels/$k.o | sed -E 's#\s+/\* 0x[0-9a-f]{16} \*/##' | grep -v "^\s*$" > profiles/sass/${1:-r01}_$k.sass
done
