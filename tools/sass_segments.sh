#!/bin/bash
# Static SASS of one kernel (name regex) from the in-tree object: instruction counts per barrier-delimited segment.
# usage: tools/sass_segments.sh <object.o> <mangled-name-regex> [dump-file]
OBJ=${1:-/tmp/msda_b200.o}; PAT=${2:-'msda_bwd_mma_kernelI13__nv_bfloat16.*Lb0E'}; OUT=${3:-/tmp/kernel.sass}
cuobjdump -sass "$OBJ" | awk -v pat="$PAT" '/Function :/ {on = ($0 ~ pat)} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//' > "$OUT"
awk 'BEGIN{n=0;s=0} {n++; if ($0 ~ /BAR.SYNC/) {printf "seg %d: %d instrs (ends line %d)\n", ++s, n, NR; n=0}} END{printf "tail: %d instrs, total %d\n", n, NR}' "$OUT"
