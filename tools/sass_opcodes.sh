#!/bin/bash
# Per-kernel count of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
# LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads / stores, HMMA = mma.sync side tiles, BRA.U.ANY = R2UR waterfall loops.
#   tools/sass_opcodes.sh > profiles/rNN_sass_opcodes.txt
cd "$(dirname "$0")/.."
SO=maskbit_b200/csrc/libmaskbit_b200.so
echo "# cuobjdump -sass $SO  ($(date -u +%F), $(nvcc --version | grep release | sed 's/.*release //'))"
printf "%-72s %8s %7s %6s %6s %8s %8s %6s %10s\n" kernel UTCHMMA .2CTA LDTM STTM UTMALDG UTMASTG HMMA BRA.U.ANY
cuobjdump -sass $SO | c++filt | awk '
/Function :/ { if (name != "") out(); name=$0; sub(/.*Function : /, "", name); sub(/\(.*/, "", name); u=c2=l=s=tl=ts=h=b=0 }
/UTCHMMA/ {u++} /UTCHMMA.2CTA/ {c2++} /LDTM/ {l++} /STTM/ {s++} /UTMALDG/ {tl++} /UTMASTG/ {ts++} / HMMA/ {h++} /BRA.U.ANY/ {b++}
function out() { if (u+l+s+tl+ts+h > 0) printf "%-72s %8d %7d %6d %6d %8d %8d %6d %10d\n", substr(name,1,72), u, c2, l, s, tl, ts, h, b }
END { out() }' | sort
