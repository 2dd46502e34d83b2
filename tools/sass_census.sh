#!/bin/bash
# SASS opcode census of libgwd_b200.so: proves which kernels are tcgen05 / TMEM / TMA (UTCHMMA, LDTM, UTMALDG, UTMASTG, UTCBAR)
# and which run on warp-level MMA (HMMA); ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch).  Usage: tools/sass_census.sh > profiles/rNN_sass_census.txt
set -e
LIB="$(dirname "$0")/../gw-depth_b200/libgwd_b200.so"
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
echo "# $(basename "$LIB") $(stat -c %s "$LIB") bytes, $(grep -c 'Function :' "$TMP") kernels"
echo "# whole library"
for op in UTCHMMA UTCQMMA UTCOMMA LDTM STTM UTMALDG UTMASTG UTMAPF UTCBAR UTCCP SYNCS HMMA ACQBULK PREEXIT; do
  printf "%-8s %d\n" $op "$(grep -c "^ *\/\*[0-9a-f]*\*\/ *\(@!\?U\?P[0-9T]* \)\?$op" "$TMP" || true)"
done
echo "# per kernel (only kernels with tensor-core or TMA instructions)"
awk '/Function :/ {name=$3} 
     / UTCHMMA/ {t[name]++} / LDTM/ {l[name]++} / UTMALDG/ {g[name]++} / UTMASTG/ {s[name]++} / UTCBAR/ {b[name]++} / HMMA/ {h[name]++}
     END {for (n in t) names[n]=1; for (n in l) names[n]=1; for (n in g) names[n]=1; for (n in s) names[n]=1; for (n in h) names[n]=1;
          for (n in names) printf "UTCHMMA=%-4d LDTM=%-3d UTMALDG=%-3d UTMASTG=%-2d UTCBAR=%-3d HMMA=%-4d %s\n", t[n], l[n], g[n], s[n], b[n], h[n], n}' "$TMP" | sort -k7 | c++filt | sed 's/(anonymous namespace):://g' | cut -c1-190
rm -f "$TMP"
