#!/bin/bash
# SASS evidence for the product library: instruction counts that prove TMA / mbarrier / DMMA (and, for the f32 path,
# tcgen05) are in the binary.  Usage: bash tools/sass_summary.sh > profiles/rNN_sass_summary.txt
so=nalgebra_b200/libnalgebra_b200.so
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  cuobjdump -sass $so  ($(/usr/local/cuda/bin/nvcc --version | tail -1))"
cuobjdump -sass $so > /tmp/nab_sass.txt
echo "sm_100a cubins: $(cuobjdump -lelf $so | grep -c sm_100a)"
for m in UTMALDG UTMASTG UBLKCP SYNCS DMMA DFMA DMUL DADD FFMA HMMA UTCHMMA UTCQMMA "UTC.*MMA" LDTM STTM UTCBAR REDUX CREDUX "LDG.E.ENL2.256" "STG.E.ENL2.256" BAR.SYNC; do
  printf "%-16s %6d\n" "$m" "$(grep -cE "$m" /tmp/nab_sass.txt)"
done
echo "# per kernel (Function : name, then DMMA / UTMALDG / UTC*MMA / CREDUX counts)"
awk '/Function :/ {name=$3} /DMMA/ {d[name]++} /UTMALDG/ {t[name]++} /UTC.*MMA/ {u[name]++} /CREDUX/ {c[name]++} END {for (n in d) printf "%s DMMA=%d UTMALDG=%d\n", n, d[n], t[n]; for (n in u) printf "%s UTCMMA=%d UTMALDG=%d\n", n, u[n], t[n]; for (n in c) printf "%s CREDUX=%d\n", n, c[n]}' /tmp/nab_sass.txt | sort
