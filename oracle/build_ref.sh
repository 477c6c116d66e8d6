#!/bin/bash
# Pins the oracle against the reference itself, for anyone who HAS the sources the reference fetches at CMake time
# (this container has no network and none of them: SURVEY.md 8c, DESIGN.md section 5 -- so this recipe has never run
# here; it follows /root/reference/CMakeLists.txt:47-240 line by line and is kept short enough to fix by hand).
#
#   oracle/build_ref.sh SRC_DIR [REFERENCE_DIR]
#
# SRC_DIR must hold checkouts named exactly:
#   ropebwt3      lh3/ropebwt3       @ 0ea3919ed21f10d857d508b0ea728a27abbd8a35   (CMakeLists.txt:154-156)
#   abPOA         yangao07/abPOA     @ e6bb6fdfa40d573558e5e2b545bdf5769631eaaf   (:97-99)
#   ksw2          lh3/ksw2           @ HEAD (the reference pins nothing, :116-118)
#   htslib        samtools/htslib    @ the tag of CMakeLists.txt:66-68, with its htscodecs submodule
#   libdeflate    ebiggers/libdeflate @ 020133854ff73b8506fe59f92a9b5b622d360716   (:47-49)
#   rapidfuzz-cpp, interval-tree, spdlog  (header-only use; :131-141, :177-179)
# REFERENCE_DIR defaults to /root/reference.  Outputs go to oracle/_ref/ only (git-ignored): the SVDSS binary, then
# `oracle/ref_vectors.sh` runs it on the worlds of tests/sv_world.py and writes tests/golden/ref_*.{sfs,vcf,clusters}
# which tests/test_ref_golden.py compares the oracle and the GPU path with (it skips while they are absent).
set -euo pipefail
SRC=${1:?usage: oracle/build_ref.sh SRC_DIR [REFERENCE_DIR]}
REF=${2:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
mkdir -p "$OUT/obj"
CC=${CC:-gcc}; CXX=${CXX:-g++}
J=${J:-$(nproc)}

# libdeflate + htslib (CMakeLists.txt:43-83)
make -C "$SRC/libdeflate" -j"$J" libdeflate.a
( cd "$SRC/htslib" && autoheader && autoreconf -i && ./configure --disable-libcurl --disable-gcs --with-libdeflate \
    "CFLAGS=-O3 -I$SRC/libdeflate" "LDFLAGS=-L$SRC/libdeflate" && make -j"$J" libhts.a )
# abPOA: libabpoa.a (its own Makefile / CMake; -march=native unless the sed of CMakeLists.txt:89-96 is applied)
( cd "$SRC/abPOA" && make -j"$J" libabpoa ) || ( cd "$SRC/abPOA" && mkdir -p build && cd build && cmake .. && make -j"$J" abpoa )
ABPOA_LIB=$(find "$SRC/abPOA" -name libabpoa.a | head -1)
# ksw2: the two objects the reference links (CMakeLists.txt:224-225)
make -C "$SRC/ksw2" ksw2_extz2_sse.o ksw2_extd2_sse.o
# ropebwt3: the 15 C files of CMakeLists.txt:169, flags of :163
for f in build sais-ss libsais16x64 fm-index rld0 mrope rope io rle kthread kalloc misc ssa dawg libsais16; do
  $CC -g -Wall -O2 -fopenmp -c "$SRC/ropebwt3/$f.c" -o "$OUT/obj/rb3_$f.o"
done
# the reference's own 12 translation units (CMakeLists.txt:20), C++14, then the link line of :205-240
for f in assembler bam caller chromosomes clipper clusterer config ping_pong sfs smoother sv main; do
  $CXX -std=c++14 -O3 -fopenmp -Wall -Wextra -I"$SRC/htslib" -I"$SRC/abPOA/include" -I"$SRC/ksw2" -I"$SRC/rapidfuzz-cpp" \
       -I"$SRC/interval-tree/include" -I"$SRC/ropebwt3" -I"$SRC/spdlog/include" -I"$REF" -c "$REF/$f.cpp" -o "$OUT/obj/$f.o"
done
$CXX -fopenmp -o "$OUT/SVDSS" "$OUT"/obj/*.o "$ABPOA_LIB" "$SRC/ksw2/ksw2_extz2_sse.o" "$SRC/ksw2/ksw2_extd2_sse.o" \
     "$SRC/htslib/libhts.a" "$SRC/libdeflate/libdeflate.a" -lz -llzma -lbz2 -lpthread
echo "built $OUT/SVDSS"
