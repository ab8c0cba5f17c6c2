#!/bin/bash
# oracle/build_ref.sh -- builds the REAL reference code for the hot path into oracle/_ref/.
#
#  * libv0_ref.so : the reference's V0 (namespace v0, /root/reference/core.cu:11-54) streamed
#    straight from where it lies into g++ (no copy of the source is written anywhere), with
#    oracle/ref_v0_shim_tail.inc appended to give it a C name.  Flags: -O3 -ffp-contract=off
#    (mandatory: keeps sub/mul/add separately rounded like the README's host build,
#    SURVEY.md section 8c) -fopenmp for the per-chunk "V0 OpenMP" wrapper.
#    -march=native is NOT used: the .so travels to a GPU box with a different CPU.
#  * ref_main (optional, needs nvcc): the whole unmodified reference benchmark
#    (main.cu + core.cu) for sm_100a, README.md:20's build line with sm_70 -> sm_100a and the
#    missing <thrust/extrema.h> force-included (SURVEY.md defect D0).  Only runs on a GPU box.
#
# Outputs only into oracle/_ref/ (git-ignored; NOT gpurun-ignored so it travels).
# The reference's own build system (a single README line) is not used.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -f "$REF/core.cu" ] || { echo "build_ref: $REF/core.cu not present; keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OUT"
# sanity: the line range must still be exactly `namespace v0 { ... }`
sed -n '11p' "$REF/core.cu" | grep -q 'namespace v0' || { echo "build_ref: core.cu:11 is not 'namespace v0'" >&2; exit 1; }
sed -n '55,56p' "$REF/core.cu" | grep -q 'v1' || { echo "build_ref: core.cu:55-56 is not the start of v1" >&2; exit 1; }
{ printf '#include <math.h>\n#include <stdlib.h>\n'; sed -n '11,54p' "$REF/core.cu"; cat "$HERE/ref_v0_shim_tail.inc"; } |
    /usr/bin/g++ -x c++ - -O3 -ffp-contract=off -fopenmp -fPIC -shared -o "$OUT/libv0_ref.so"
echo "built $OUT/libv0_ref.so"
#  * ref_gpu_probe (needs nvcc): oracle/ref_gpu_probe.cu (our driver) + the reference's core.cu included by
#    path, one variant on one shape with the reference's generator and timing; bench.py's `ref_gpu` record.
if command -v nvcc >/dev/null && [ ! -x "$OUT/ref_gpu_probe" -o "$HERE/ref_gpu_probe.cu" -nt "$OUT/ref_gpu_probe" ]; then
    nvcc -O2 -Xcompiler -fopenmp -arch=sm_100a -include thrust/extrema.h -I"$REF" "$HERE/ref_gpu_probe.cu" -o "$OUT/ref_gpu_probe" 2>/dev/null \
        && echo "built $OUT/ref_gpu_probe" || echo "build_ref: ref_gpu_probe did not build (optional)" >&2
fi
if [ "${BUILD_REF_MAIN:-1}" = "1" ] && command -v nvcc >/dev/null; then
    nvcc -Xcompiler -fopenmp -arch=sm_100a -include thrust/extrema.h -I"$REF" "$REF/main.cu" -o "$OUT/ref_main" 2>/dev/null \
        && echo "built $OUT/ref_main" || echo "build_ref: ref_main did not build (optional)" >&2
fi
