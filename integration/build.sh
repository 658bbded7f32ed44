#!/usr/bin/env bash
# integration/build.sh — compiles the reference-side binding (amici_gpu.cpp) against the UNMODIFIED reference headers and links
# it, with one pre-generated reference model and the test driver, to oracle/_ref/libamici.a and amici_b200/lib/libamici_b200.so:
#   integration/_build/libbinding_<model>.so      (git-ignored; travels to the GPU box like oracle/_ref)
# Needs /root/reference (headers + the model's generated sources) and a built oracle/_ref; a no-op elsewhere.
set -euo pipefail
REF=${AMICI_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
OUT="$HERE/_build"
ORA="$ROOT/oracle/_ref"
if [ ! -d "$REF/src" ] || [ ! -f "$ORA/libamici.a" ]; then
  echo "integration/build.sh: reference or oracle/_ref not present — keeping prebuilt integration/_build" >&2
  exit 0
fi
mkdir -p "$OUT"
SS="$REF/ThirdParty/SuiteSparse"; SD="$REF/ThirdParty/sundials"
INC="-I$ORA/include -I$REF/include -I$REF/ThirdParty/gsl -I$SD/include -I$SD/src -I$SS/SuiteSparse_config -I$SS/KLU/Include -I$SS/AMD/Include -I$SS/BTF/Include -I$SS/COLAMD/Include -I$ROOT/include -I$HERE"
for m in ${MODELS:-jakstat_adjoint steadystate}; do
  MD="$REF/models/model_${m}_py"
  OBJ="$ORA/obj/model_$m"          # the model's objects, already compiled by oracle/build_ref.sh
  g++ -std=c++20 -O2 -fPIC -fopenmp -w $INC -I"$MD" -shared -o "$OUT/libbinding_$m.so" \
      "$HERE/amici_gpu.cpp" "$HERE/test_driver.cpp" $(ls "$OBJ"/*.o | grep -v ref_driver.o) \
      "$ORA/libamici.a" "$ORA/libsundials.a" "$ORA/libsuitesparse.a" \
      -L"$ROOT/amici_b200/lib" -lamici_b200 -Wl,-rpath,'$ORIGIN/../../amici_b200/lib' -ldl -lm
  echo "built $OUT/libbinding_$m.so"
done
