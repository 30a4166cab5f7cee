#!/bin/bash
# SASS opcode evidence per kernel of the shipped library (runs without a GPU):
#   bash tools/sass_summary.sh > profiles/r02_sass_opcodes.md
# UTCHMMA = tcgen05.mma (bf16), UTMALDG = TMA tensor load, UTMASTG = TMA store, LDTM / STTM = tcgen05.ld / st (TMEM),
# UTCBAR = tcgen05.commit -> mbarrier, HMMA = legacy mma.sync (should be absent), SYNCS = mbarrier ops.
lib=${1:-brats21_b200/libb21.so}
echo "# SASS opcode counts per kernel — $(basename $lib), $(date -u +%Y-%m-%d), cuobjdump -sass (sm_100a)"
echo
echo "| kernel | UTCHMMA | UTMALDG | UTMASTG | LDTM | STTM | UTCBAR | SYNCS | HMMA | instructions |"
echo "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"
cuobjdump -sass "$lib" | awk '
  /Function :/ { if (name != "") flush(); name=$3; n=0; delete c }
  /^ +\/\*[0-9a-f]+\*\// { n++; for (k in keys) if (index($0, " " keys[k])) c[keys[k]]++ }
  BEGIN { split("UTCHMMA UTMALDG UTMASTG LDTM STTM UTCBAR SYNCS HMMA", keys, " ") }
  function flush() { printf "%s %d %d %d %d %d %d %d %d %d\n", name, c["UTCHMMA"], c["UTMALDG"], c["UTMASTG"], c["LDTM"], c["STTM"], c["UTCBAR"], c["SYNCS"], c["HMMA"], n }
  END { flush() }' | while read name a b c d e f g h n; do
    d2=$(echo "$name" | c++filt | sed 's/(.*//; s/void //; s/b21:://')
    echo "| \`$d2\` | $a | $b | $c | $d | $e | $f | $g | $h | $n |"
  done | sort -t'|' -k3,3nr -k11,11nr
