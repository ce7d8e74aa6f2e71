#!/usr/bin/env bash
# Compiles the UNMODIFIED reference sources (read where they lie under /root/reference/scr)
# against oracle/shim into oracle/_ref/: the full `dbslmm` CLI and a harness library.
# Outputs only under oracle/_ref/ (git-ignored).  Not run on the GPU box (no /root/reference there).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_SRC:-/root/reference/scr}"
OUT="$HERE/_ref"
CXX=/usr/bin/g++
FLAGS="-O3 -std=c++14 -fopenmp -mavx2 -mfma -fPIC -w -I $HERE/shim -I $REF"
[ -d "$REF" ] || { echo "no reference sources at $REF"; exit 0; }
mkdir -p "$OUT/obj"
SRCS="dtpr dbslmm dbslmmfit calc_asymptotic_variance subset_to_test_and_training helpers"
for s in $SRCS; do
  $CXX $FLAGS -c "$REF/$s.cpp" -o "$OUT/obj/$s.o"
done
$CXX $FLAGS -c "$REF/main_dbslmm.cpp" -o "$OUT/obj/main_dbslmm.o"
OBJS=""; for s in $SRCS; do OBJS="$OBJS $OUT/obj/$s.o"; done
$CXX -fopenmp -o "$OUT/dbslmm_ref" "$OUT/obj/main_dbslmm.o" $OBJS
$CXX $FLAGS -shared -o "$OUT/libref_harness.so" "$HERE/ref_harness.cpp" $OBJS
# the reference's second binary, `valid` (external validation): its Makefile never builds it, the sources compile as they are
$CXX $FLAGS -c "$REF/validate.cpp" -o "$OUT/obj/validate.o"
$CXX $FLAGS -c "$REF/main_valid.cpp" -o "$OUT/obj/main_valid.o"
$CXX -fopenmp -o "$OUT/valid_ref" "$OUT/obj/main_valid.o" "$OUT/obj/validate.o" "$OUT/obj/dtpr.o" "$OUT/obj/helpers.o"
echo "built $OUT/dbslmm_ref, $OUT/valid_ref and $OUT/libref_harness.so"
