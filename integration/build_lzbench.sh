#!/bin/bash
# Build the reference harness twice from the reference's own sources (gcc/g++ on those files directly, not
# its Makefile), lz4 rows only:
#   oracle/_ref/lzbench_ref    unmodified harness  -> `lzbench -elz4 -b64 FILE` (the CPU row, BASELINE configs[0])
#   oracle/_ref/lzbench_b200   harness + the lz4_b200 row (integration/patch_lzbench.py applied to a staged copy
#                              under build/, never committed) linked against lzbench_b200/_lib/liblz4b200.so
# Outputs go to oracle/_ref/ (git-ignored, travels to the GPU box).  Needs /root/reference: build container only.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${REF:-/root/reference}"
[ -f "$REF/_lzbench/lzbench.cpp" ] || { echo "reference tree not present at $REF: keeping prebuilt binaries"; exit 0; }
OUT="$ROOT/oracle/_ref"; STG="$ROOT/build/lzbench_stage"
mkdir -p "$OUT" "$STG/_lzbench"
DEFS=""
for x in BLOSCLZ BRIEFLZ BROTLI BZIP2 CRUSH CSC DENSITY FASTLZ FASTLZMA2 GIPFELI GLZA LIBDEFLATE LIZARD LZ3 LZF LZFSE LZG \
         LZHAM LZJB LZLIB LZMA LZMAT LZO LZRW LZSSE LZVN PITHY QUICKLZ SHRINKER SLZ SNAPPY TORNADO UCL WFLZ XPACK XZ \
         YALZ77 YAPPY ZLIB ZLING ZSTD; do DEFS="$DEFS -DBENCH_REMOVE_$x"; done
CXXFLAGS="-std=c++11 -O3 -DNDEBUG -fomit-frame-pointer -fstrict-aliasing -ffast-math -w"
CFLAGS="-O3 -DNDEBUG -fomit-frame-pointer -fstrict-aliasing -ffast-math -w"
cd "$STG"
gcc $CFLAGS -c "$REF/lz4/lz4.c" -o lz4.o
gcc $CFLAGS -c "$REF/lz4/lz4hc.c" -o lz4hc.o
# 1. unmodified
g++ $CXXFLAGS $DEFS -I"$REF" -I"$REF/_lzbench" -c "$REF/_lzbench/lzbench.cpp" -o lzbench_ref.o
g++ $CXXFLAGS $DEFS -I"$REF" -I"$REF/_lzbench" -c "$REF/_lzbench/compressors.cpp" -o compressors_ref.o
g++ -o "$OUT/lzbench_ref" lzbench_ref.o compressors_ref.o lz4.o lz4hc.o -lpthread
# 2. with the lz4_b200 row
cp "$REF"/_lzbench/*.cpp "$REF"/_lzbench/*.h _lzbench/
python "$ROOT/integration/patch_lzbench.py" "$STG/_lzbench"
g++ $CXXFLAGS $DEFS -DBENCH_HAS_LZ4_B200 -I"$REF" -I"$STG/_lzbench" -c _lzbench/lzbench.cpp -o lzbench_b200.o
g++ $CXXFLAGS $DEFS -DBENCH_HAS_LZ4_B200 -I"$REF" -I"$STG/_lzbench" -c _lzbench/compressors.cpp -o compressors_b200.o
g++ -o "$OUT/lzbench_b200" lzbench_b200.o compressors_b200.o lz4.o lz4hc.o \
    -L"$ROOT/lzbench_b200/_lib" -llz4b200 -Wl,-rpath,'$ORIGIN/../../lzbench_b200/_lib' -lpthread
echo "built $OUT/lzbench_ref and $OUT/lzbench_b200"
