#!/bin/bash
# Builds the facade test binaries against the reference's own algorithm headers.  Only possible
# where /root/reference exists (this container); the binaries travel to the GPU box.
#
# A quoted #include inside a QCSim header is resolved in that header's own directory first, so the
# drop-in is done the way a maintainer would do it: QubitRegister.h is REPLACED in
# the QCSim source directory.  /root/reference is read-only, so the replacement happens in a staging
# directory of symlinks (tests/cpp/_stage, git-ignored): every QCSim header is linked as is
# (QubitRegisterDebug.h included: it compiles unchanged against the drop-in), QubitRegister.h points
# at qcsim_b200/cpp/.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${REFERENCE:-/root/reference}/QCSim"
[ -d "$REF" ] || { echo "no reference tree at $REF: keeping prebuilt binaries"; exit 0; }
stage() {  # $1 = dir, $2 = 1 to also replace QuantumFourierTransform.h by the one-call version
  rm -rf "$1"; mkdir -p "$1"
  for f in "$REF"/*.h; do ln -s "$f" "$1/$(basename "$f")"; done
  rm -f "$1/QubitRegisterCalculator.h"   # nothing may reach the CPU loops
  ln -sf "$ROOT/qcsim_b200/cpp/QubitRegister.h" "$1/QubitRegister.h"
  if [ "$2" = 1 ]; then ln -sf "$ROOT/qcsim_b200/cpp/fast/QuantumFourierTransform.h" "$1/QuantumFourierTransform.h"; fi
}
stage "$HERE/_stage/plain" 0
stage "$HERE/_stage/fast" 1
COMMON="-std=c++17 -O2 -I $ROOT/include -I $ROOT/oracle/eigen_shim"
LINK="-L $ROOT/qcsim_b200 -lqcsim_b200 -Wl,-rpath,\$ORIGIN/../../qcsim_b200"
g++ $COMMON -I "$HERE/_stage/plain" "$HERE/facade_test.cpp" -o "$HERE/facade_test.bin" $LINK
g++ $COMMON -DFACADE_FAST -I "$HERE/_stage/fast" "$HERE/facade_test.cpp" -o "$HERE/facade_test_fast.bin" $LINK
echo built
