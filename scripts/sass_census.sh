#!/bin/bash
# SASS instruction census of the in-tree library: which kernels use the Blackwell tensor-core / TMA paths.
# usage: scripts/sass_census.sh [path/to/libax_whisper.so] > profiles/rNN_sass_census.txt
SO=${1:-$(dirname "$0")/../whisper.axera_b200/libax_whisper.so}
echo "# cuobjdump -sass $(basename $SO) ($(date -u +%F)), counts per kernel: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st,"
echo "# UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, UTCBAR = tcgen05.commit, HMMA = mma.sync (comparator kernels only)"
cuobjdump -sass "$SO" | awk '
  /Function :/ { name=$3; sub(/^_ZN5b200w[0-9]*_?/, "", name); names[++n]=name; cur=name }
  /UTCHMMA/ { c[cur,"UTCHMMA"]++; if ($0 ~ /2CTA/) c[cur,"UTCHMMA.2CTA"]++ }
  /LDTM/ { c[cur,"LDTM"]++ }
  /STTM/ { c[cur,"STTM"]++ }
  /UTMALDG/ { c[cur,"UTMALDG"]++ }
  /UTMASTG/ { c[cur,"UTMASTG"]++ }
  /UTMAREDG/ { c[cur,"UTMAREDG"]++ }
  /UTCBAR/ { c[cur,"UTCBAR"]++ }
  /HMMA/ && !/UTCHMMA/ { c[cur,"HMMA"]++ }
  /UCGABAR|CGABAR/ { c[cur,"CLUSTER_BAR"]++ }
  END {
    split("UTCHMMA UTCHMMA.2CTA LDTM STTM UTMALDG UTMASTG UTMAREDG UTCBAR HMMA CLUSTER_BAR", keys, " ")
    printf "%-110s", "kernel"; for (k=1;k<=10;k++) printf " %12s", keys[k]; printf "\n"
    for (i=1;i<=n;i++) { tot=0; for (k=1;k<=10;k++) tot+=c[names[i],keys[k]];
      if (tot>0) { printf "%-110s", substr(names[i],1,110); for (k=1;k<=10;k++) { printf " %12d", c[names[i],keys[k]]; sum[k]+=c[names[i],keys[k]] } printf "\n" } }
    printf "%-110s", "TOTAL"; for (k=1;k<=10;k++) printf " %12d", sum[k]; printf "\n"
  }'
