#!/bin/bash
# profiles/sass_summary.txt: the tcgen05 / TMA / TMEM instructions of the shipped library (cuobjdump -sass | grep -c)
cd /root/repo
OUT=profiles/sass_summary.txt
SO=conv_tasnet_b200/libctn_b200.so
cuobjdump -sass $SO > /tmp/ctn.sass 2>/dev/null
{
  echo "# cuobjdump -sass $SO | grep -c <mnemonic>   ($(date -u +%Y-%m-%d), sources $(git rev-parse --short HEAD)+)"
  for m in UTCHMMA UTCQMMA "UTCHMMA.*TS\|UTCHMMA.2CTA" UTMALDG UTMALDG.2D.MULTICAST UTMAREDG UTMASTG LDTM STTM UTCBAR UTCBAR.MULTICAST UTCATOMSWS "SYNCS" "ELECT" LDGSTS "ACQBULK\|UCGABAR" UBLKCP; do
    printf "%-28s %s\n" "$m" "$(grep -c "$m" /tmp/ctn.sass)"
  done
  echo "# per kernel (function name : UTCHMMA / UTMALDG / LDTM / STTM / UTMAREDG)"
  awk '/Function :/{name=$3} /UTCHMMA/{a[name]++} /UTMALDG/{b[name]++} /LDTM/{c[name]++} /STTM/{d[name]++} /UTMAREDG/{e[name]++} END{for(n in a) printf "%s : %d / %d / %d / %d / %d\n", n, a[n], b[n], c[n], d[n], e[n]}' /tmp/ctn.sass | c++filt | sed 's/ctn::(anonymous namespace):://' | sort | cut -c1-300
} > $OUT
cat $OUT | head -40
