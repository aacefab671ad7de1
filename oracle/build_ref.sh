#!/bin/bash
# Build the UNMODIFIED reference (nicholasmr/specfab, Fortran 90) into oracle/_ref/ on a machine that has gfortran,
# LAPACK/BLAS and numpy.f2py -- the toolchain this project's container and GPU image do NOT have (DESIGN.md section 6:
# parity is "unpinned against a compiled reference" until this recipe has been run once somewhere).
#
#   oracle/build_ref.sh [/path/to/reference]        default: /root/reference
#   python oracle/make_ref_fixtures.py               writes tests/golden/ref_compiled.npz
#   python -m pytest tests/test_oracle_golden.py     the oracle (and through it every GPU parity test) is then pinned
#
# Follows the reference's own recipe (src/Makefile:89-100, 125-136: objects in dependency order, libspecfab.a, f2py),
# but compiles the sources where they lie and writes every product (objects, .mod files, the extension module) under
# oracle/_ref/ -- nothing is written into the reference tree and no reference source is copied into this repository.
set -euo pipefail
REF=${1:-/root/reference}
SRC=$REF/src
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OBJ=$OUT/obj
command -v gfortran >/dev/null || { echo "build_ref.sh: gfortran not found (this image has no Fortran compiler)"; exit 3; }
python3 -c "import numpy.f2py" 2>/dev/null || { echo "build_ref.sh: numpy.f2py not importable"; exit 3; }
mkdir -p "$OBJ" "$OUT/specfabpy"
FC="gfortran -ffree-line-length-none -m64 -fPIC -Wno-integer-division -O2 -mcmodel=small -I$SRC -I$SRC/include -J$OBJ -I$OBJ"
# src/Makefile:40-44: DEPLESS, then WITHDEPS, then specfab.f90
MODS="header tensorproducts mandel reducedform gaunt golf rheologies elasticities moments lambdasolver dynamics damage homogenizations enhancementfactors rotation wavepropagation deformationmodes idealstate frames specfab_elmer specfab"
M77="amach derv1 dnqsol dnrm2 erfin ermsg ierm1 ierv1"
OBJS=""
for f in $M77; do gfortran -fPIC -O2 -c "$SRC/include/math77/$f.f" -o "$OBJ/$f.o"; OBJS="$OBJS $OBJ/$f.o"; done
for m in $MODS; do (cd "$SRC" && $FC -c "$m.f90" -o "$OBJ/$m.o"); OBJS="$OBJS $OBJ/$m.o"; done
rm -f "$OUT/libspecfab.a"; ar rcs "$OUT/libspecfab.a" $OBJS
# the f2py interface of src/specfabpy.f90 (src/Makefile:93-97); run from the object directory so that no file lands in $SRC
(cd "$OBJ" && python3 -m numpy.f2py --no-lower -m specfabpy -h specfabpy.pyf "$SRC/specfabpy.f90" --quiet --overwrite-signature \
  && python3 -m numpy.f2py -lm -llapack -lblas -L"$OUT" -lspecfab -I"$OBJ" -I"$SRC" -I"$SRC/include" $OBJS -c specfabpy.pyf "$SRC/specfabpy.f90" \
       --f90flags="-ffree-line-length-none -mcmodel=small -I$SRC -I$SRC/include" --quiet)
mv -f "$OBJ"/specfabpy.cpython* "$OUT/specfabpy/"
printf 'from .specfabpy import specfabpy as specfab\n' > "$OUT/specfabpy/__init__.py"
echo "built: $OUT/specfabpy ($(ls "$OUT/specfabpy" | tr '\n' ' '))"
