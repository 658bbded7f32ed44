#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE, not product code.
#
# Compiles the UNMODIFIED reference (AMICI + its vendored SUNDIALS 7.1.1 and
# SuiteSparse KLU) from the sources where they lie under /root/reference into
# oracle/_ref/ (git-ignored; travels to the GPU box with the gpurun snapshot).
# No reference source is copied into this repo and the reference's own build
# system (cmake) is not run: the few headers cmake would have generated
# (sundials_config.h, sundials_export.h, amici/version.h) and a 3-function
# cblas shim are written here.
#
# Outputs:  oracle/_ref/lib{suitesparse,sundials,amici}.a
#           oracle/_ref/libref_<model>.so   (one per pre-generated model; C entry
#                                            points of oracle/ref_driver.cpp)
set -euo pipefail
REF=${AMICI_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
JOBS=${JOBS:-$(nproc)}
MODELS=${MODELS:-"jakstat_adjoint steadystate events neuron dirac nested_events calvetti robertson robertson_ode massaction_small massaction"}
CFLAGS_COMMON="-O2 -fPIC -fopenmp -w"

if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: reference not present at $REF — keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT/include/sundials" "$OUT/include/amici" "$OUT/obj"

# ---- generated headers -------------------------------------------------------
cat > "$OUT/include/sundials/sundials_export.h" <<'EOF'
#ifndef SUNDIALS_EXPORT_H
#define SUNDIALS_EXPORT_H
#define SUNDIALS_EXPORT
#define SUNDIALS_NO_EXPORT
#define SUNDIALS_DEPRECATED __attribute__((__deprecated__))
#define SUNDIALS_DEPRECATED_EXPORT SUNDIALS_EXPORT SUNDIALS_DEPRECATED
#define SUNDIALS_DEPRECATED_NO_EXPORT SUNDIALS_NO_EXPORT SUNDIALS_DEPRECATED
#define SUNDIALS_STATIC_DEFINE
#endif
EOF
cat > "$OUT/include/sundials/sundials_config.h" <<'EOF'
#ifndef _SUNDIALS_CONFIG_H
#define _SUNDIALS_CONFIG_H
#include "sundials/sundials_export.h"
#if defined(__cplusplus)
#define SUNDIALS_NOEXCEPT noexcept
#else
#define SUNDIALS_NOEXCEPT
#endif
#ifndef SUNDIALS_DEPRECATED_MSG
#define SUNDIALS_DEPRECATED_MSG(msg) __attribute__((__deprecated__(msg)))
#endif
#ifndef SUNDIALS_DEPRECATED_EXPORT_MSG
#define SUNDIALS_DEPRECATED_EXPORT_MSG(msg) SUNDIALS_EXPORT SUNDIALS_DEPRECATED_MSG(msg)
#endif
#ifndef SUNDIALS_DEPRECATED_NO_EXPORT_MSG
#define SUNDIALS_DEPRECATED_NO_EXPORT_MSG(msg) SUNDIALS_NO_EXPORT SUNDIALS_DEPRECATED_MSG(msg)
#endif
#define SUNDIALS_VERSION "7.1.1"
#define SUNDIALS_VERSION_MAJOR 7
#define SUNDIALS_VERSION_MINOR 1
#define SUNDIALS_VERSION_PATCH 1
#define SUNDIALS_VERSION_LABEL ""
#define SUNDIALS_GIT_VERSION ""
#define SUNDIALS_C_COMPILER_HAS_BUILTIN_EXPECT
#define SUNDIALS_C_COMPILER_HAS_ATTRIBUTE_UNUSED
#define SUNDIALS_DOUBLE_PRECISION 1
#define SUNDIALS_INT64_T 1
#define SUNDIALS_INDEX_TYPE int64_t
#define SUNDIALS_HAVE_POSIX_TIMERS
#define SUNDIALS_LOGGING_LEVEL 0
#define SUN_C_COMPILER "GNU"
#define SUN_C_COMPILER_VERSION ""
#define SUN_C_COMPILER_FLAGS ""
#define SUN_CXX_COMPILER "GNU"
#define SUN_CXX_COMPILER_VERSION ""
#define SUN_CXX_COMPILER_FLAGS ""
#define SUN_FORTRAN_COMPILER ""
#define SUN_FORTRAN_COMPILER_VERSION ""
#define SUN_FORTRAN_COMPILER_FLAGS ""
#define SUN_BUILD_TYPE "Release"
#define SUN_JOB_ID ""
#define SUN_JOB_START_TIME ""
#define SUN_TPL_LIST "KLU"
#define SUN_TPL_LIST_SIZE "1"
#define SUNDIALS_SPACK_VERSION ""
#define SUNDIALS_KLU_ENABLED
#define SUN_KLU_VERSION ""
#define SUN_GINKGO_VERSION ""
#define SUN_HYPRE_VERSION ""
#define SUN_KOKKOS_VERSION ""
#define SUN_KOKKOS_KERNELS_VERSION ""
#define SUN_LAPACK_VERSION ""
#define SUN_MAGMA_VERSION ""
#define SUN_MPI_C_COMPILER ""
#define SUN_MPI_C_VERSION ""
#define SUN_MPI_CXX_COMPILER ""
#define SUN_MPI_CXX_VERSION ""
#define SUN_MPI_FORTRAN_COMPILER ""
#define SUN_MPI_FORTRAN_VERSION ""
#define SUN_ONEMKL_VERSION ""
#define SUN_OPENMP_VERSION ""
#define SUN_PETSC_VERSION ""
#define SUN_PTHREADS_VERSION ""
#define SUN_RAJA_VERSION ""
#define SUN_SUPERLUDIST_VERSION ""
#define SUN_SUPERLUMT_VERSION ""
#define SUN_TRILLINOS_VERSION ""
#define SUN_XBRAID_VERSION ""
#define SUN_CUDA_VERSION ""
#define SUN_CUDA_COMPILER ""
#define SUN_CUDA_ARCHITECTURES ""
#define SUN_HIP_VERSION ""
#define SUN_AMDGPU_TARGETS ""
#define SUN_SYCL_COMPILER ""
#define SUN_SYCL_COMPILER_VERSION ""
#define SUN_SYCL_COMPILER_FLAGS ""
#define SUNDIALS_NVECTOR_SERIAL 1
#define SUNDIALS_SUNMATRIX_BAND 1
#define SUNDIALS_SUNMATRIX_DENSE 1
#define SUNDIALS_SUNMATRIX_SPARSE 1
#define SUNDIALS_SUNLINSOL_BAND 1
#define SUNDIALS_SUNLINSOL_DENSE 1
#define SUNDIALS_SUNLINSOL_KLU 1
#define SUNDIALS_SUNLINSOL_PCG 1
#define SUNDIALS_SUNLINSOL_SPBCGS 1
#define SUNDIALS_SUNLINSOL_SPFGMR 1
#define SUNDIALS_SUNLINSOL_SPGMR 1
#define SUNDIALS_SUNLINSOL_SPTFQMR 1
#define SUNDIALS_SUNNONLINSOL_NEWTON 1
#define SUNDIALS_SUNNONLINSOL_FIXEDPOINT 1
#define SUNDIALS_CVODES 1
#define SUNDIALS_IDAS 1
#define SUNDIALS_MPI_ENABLED 0
#endif
EOF
printf '#ifndef AMICI_VERSION_H\n#define AMICI_VERSION_H\n#define AMICI_VERSION "%s"\n#endif\n' \
  "$(cat "$REF/version.txt")" > "$OUT/include/amici/version.h"
# cblas shim: AMICI calls only dgemm/dgemv/daxpy (reference src/cblas.cpp:33-62)
cat > "$OUT/include/cblas.h" <<'EOF'
#ifndef ORACLE_CBLAS_SHIM_H
#define ORACLE_CBLAS_SHIM_H
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, double,
                 const double*, int, const double*, int, double, double*, int);
void cblas_dgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int, int, double, const double*, int,
                 const double*, int, double, double*, int);
void cblas_daxpy(int, double, const double*, int, double*, int);
#endif
EOF

compile_list() { # $1=tag  $2=compiler+flags  rest=sources ; incremental, parallel
  local tag=$1 cc=$2; shift 2
  mkdir -p "$OUT/obj/$tag"
  local cmds=()
  for s in "$@"; do
    local o="$OUT/obj/$tag/$(basename "${s%.*}").o"
    if [ ! -f "$o" ] || [ "$s" -nt "$o" ]; then cmds+=("$cc -c '$s' -o '$o'"); fi
  done
  if [ ${#cmds[@]} -gt 0 ]; then printf '%s\n' "${cmds[@]}" | xargs -P "$JOBS" -I{} bash -c '{}'; fi
}

# ---- SuiteSparse: config + AMD + BTF + COLAMD + KLU --------------------------
SS="$REF/ThirdParty/SuiteSparse"
SSINC="-I$SS/SuiteSparse_config -I$SS/AMD/Include -I$SS/BTF/Include -I$SS/COLAMD/Include -I$SS/KLU/Include"
compile_list ss "gcc $CFLAGS_COMMON -DNTIMER $SSINC" \
  "$SS/SuiteSparse_config/SuiteSparse_config.c" "$SS"/AMD/Source/*.c "$SS"/BTF/Source/*.c \
  "$SS"/COLAMD/Source/*.c "$SS"/KLU/Source/*.c
rm -f "$OUT/libsuitesparse.a"; ar rcs "$OUT/libsuitesparse.a" "$OUT"/obj/ss/*.o

# ---- SUNDIALS: core + cvodes + idas + serial vector + matrices + solvers ------
SD="$REF/ThirdParty/sundials"
SDINC="-I$OUT/include -I$SD/include -I$SD/src -I$SD/src/sundials $SSINC"
SDSRC=("$SD"/src/sundials/*.c "$SD"/src/cvodes/*.c "$SD"/src/idas/*.c
       "$SD/src/nvector/serial/nvector_serial.c" "$SD/src/sunmemory/system/sundials_system_memory.c")
for d in sunmatrix/band sunmatrix/dense sunmatrix/sparse sunlinsol/band sunlinsol/dense sunlinsol/klu \
         sunlinsol/pcg sunlinsol/spbcgs sunlinsol/spfgmr sunlinsol/spgmr sunlinsol/sptfqmr \
         sunnonlinsol/newton sunnonlinsol/fixedpoint; do
  SDSRC+=("$SD"/src/$d/*.c)
done
compile_list sundials "gcc $CFLAGS_COMMON -DBUILD_SUNDIALS_LIBRARY $SDINC" "${SDSRC[@]}"
rm -f "$OUT/libsundials.a"; ar rcs "$OUT/libsundials.a" "$OUT"/obj/sundials/*.o

# ---- AMICI core (everything except the HDF5 I/O file) --------------------------
AMINC="-I$OUT/include -I$REF/include -I$REF/ThirdParty/gsl $SDINC"
AMSRC=()
for s in "$REF"/src/*.cpp; do case "$s" in */hdf5.cpp) ;; *) AMSRC+=("$s");; esac; done
compile_list amici "g++ -std=c++20 $CFLAGS_COMMON $AMINC" "${AMSRC[@]}"
gcc $CFLAGS_COMMON -I"$OUT/include" -c "$HERE/cblas_shim.c" -o "$OUT/obj/amici/cblas_shim.o"
rm -f "$OUT/libamici.a"; ar rcs "$OUT/libamici.a" "$OUT"/obj/amici/*.o

# ---- one shared library per pre-generated model --------------------------------
for m in $MODELS; do
  MD="$REF/models/model_${m}_py"
  if [ ! -d "$MD" ]; then
    # not a shipped model: written in the reference's generated-code format from the tables / ModelSpec the engine uses
    case "$m" in
      massaction*) MD="$(python "$HERE/gen_massaction_ref.py" "$m" "$OUT/gen" "$REF")" ;;
      *)           MD="$(python "$HERE/gen_ref_model.py" "$m" "$OUT/gen" "$REF")" ;;
    esac
  fi
  MSRC=()
  for s in "$MD"/*.cpp; do case "$s" in */main.cpp) ;; *) MSRC+=("$s");; esac; done
  compile_list "model_$m" "g++ -std=c++20 $CFLAGS_COMMON $AMINC -I$MD" "${MSRC[@]}"
  g++ -std=c++20 $CFLAGS_COMMON $AMINC -I"$MD" -c "$HERE/ref_driver.cpp" -o "$OUT/obj/model_$m/ref_driver.o"
  g++ -shared -fopenmp -o "$OUT/libref_$m.so" "$OUT/obj/model_$m"/*.o \
     "$OUT/libamici.a" "$OUT/libsundials.a" "$OUT/libsuitesparse.a" -ldl -lm
  echo "built $OUT/libref_$m.so"
done

# ---- the reference's HermiteSpline class behind a C entry point (oracle/ref_spline.cpp) ----------------------------
g++ -std=c++20 $CFLAGS_COMMON $AMINC -shared -o "$OUT/libref_spline.so" "$HERE/ref_spline.cpp" "$OUT/libamici.a" "$OUT/libsundials.a" \
   "$OUT/libsuitesparse.a" -ldl -lm
echo "built $OUT/libref_spline.so"
