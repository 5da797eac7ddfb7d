#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md):
#   UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAPF = TMA load / store / prefetch, HMMA = legacy mma.sync
# usage: tools/sass_opcodes.sh [lib.so] > profiles/rNN_sass_opcodes.txt
LIB="${1:-hirest_b200/libhirest_b200.so}"
echo "# cuobjdump -sass $LIB | per-kernel mnemonic counts (static instruction counts, loops counted once)"
echo "# linked libraries: $(ldd "$LIB" | awk '{print $1}' | tr '\n' ' ')"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ {f=$3}
{ for (i = 1; i <= NF; i++) if ($i ~ /^(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|HMMA|UBLKCP|MUFU)/) { split($i, b, "."); c[f" "b[1]]++ } }
END { for (k in c) print k, c[k] }' | c++filt | sed -E 's/\(anonymous namespace\)::|hb:://g; s/CUtensorMap_st/TMap/g' | sort
