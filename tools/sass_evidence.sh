#!/bin/bash
# Static evidence from the built library (no GPU needed): per-kernel resource usage and the Blackwell SASS mnemonics
# (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, HMMA = legacy mma.sync) per object file.
#   bash tools/sass_evidence.sh > profiles/r1_sass_evidence.txt
cd "$(dirname "$0")/.." || exit 1
LIB=uncertainty-vit_b200/libb200vit.so
echo "== SASS mnemonic counts per translation unit (cuobjdump -sass) =="
for o in uncertainty-vit_b200/build/*.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  printf "%-24s UTC*MMA %4d  LDTM %4d  STTM %4d  UTMALDG %4d  UTMASTG %4d  HMMA %4d  SYNCS(mbarrier) %4d\n" "$(basename "$o")" \
    "$(grep -c 'UTC[A-Z]*MMA' <<<"$s")" "$(grep -c 'LDTM' <<<"$s")" "$(grep -c 'STTM' <<<"$s")" "$(grep -c 'UTMALDG' <<<"$s")" \
    "$(grep -c 'UTMASTG' <<<"$s")" "$(grep -c ' HMMA' <<<"$s")" "$(grep -c 'SYNCS' <<<"$s")"
done
echo
echo "== resource usage per kernel (cuobjdump -res-usage) =="
cuobjdump -res-usage "$LIB" 2>/dev/null | awk '/Function/ {name=$2} /REG:/ {print name, $0}' | sed 's/^_ZN[0-9]*_GLOBAL__N__[0-9a-f_]*cu_[0-9a-f]*//' | c++filt 2>/dev/null | sort | cut -c1-260
