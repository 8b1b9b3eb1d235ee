#!/bin/bash
# SASS opcode histogram of libmpqr.so per kernel family (proves tcgen05 / TMEM / TMA in the built library).
# usage: tools/sass_histogram.sh > profiles/r2_sass_histogram.txt
LIB=mixedprecisionblockqr_b200/libmpqr.so
echo "# cuobjdump -sass $LIB ($(date -u +%F)), sm_100a; counts of selected mnemonics over the whole library"
cuobjdump -sass $LIB > /tmp/mpqr.sass
for op in UTCHMMA UTCQMMA UTCBAR UTCCP LDTM STTM UTMALDG UTMASTG UTMAREDG UTMAPF SYNCS UCGABAR_ARV UCGABAR_WAIT ACQBULK ELECT HMMA FFMA DFMA MUFU.RSQ LDG STG LDS STS ATOMG RED LDGSTS BAR.SYNC; do
  printf "%-14s %8d\n" $op $(grep -c "[[:space:]]$op" /tmp/mpqr.sass)
done
echo
echo "# per kernel: total instructions, UTCHMMA, LDTM, UTMALDG, UTMASTG+UTMAREDG, FFMA"
awk '/Function :/ {name=$3} /^[[:space:]]+\/\*[0-9a-f]+\*\// {n[name]++; if ($0 ~ /UTCHMMA/) a[name]++; if ($0 ~ /LDTM/) b[name]++; if ($0 ~ /UTMALDG/) c[name]++; if ($0 ~ /UTMASTG|UTMAREDG/) d[name]++; if ($0 ~ /FFMA/) f[name]++} END {for (k in n) printf "%7d %5d %5d %5d %5d %7d  %s\n", n[k], a[k], b[k], c[k], d[k], f[k], k}' /tmp/mpqr.sass | sort -rn | while read a b c d e f name; do printf "%7d %5d %5d %5d %5d %7d  %s\n" $a $b $c $d $e $f "$(echo $name | c++filt | cut -c1-110)"; done | head -60
