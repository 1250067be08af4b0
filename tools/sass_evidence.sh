#!/bin/sh
# Regenerates profiles/r02_sass_evidence.txt: per kernel of liboctreelib_b200.so, how many SASS instructions of the kinds
# DESIGN.md claims are actually in the sm_100a machine code (cuobjdump works without a GPU):
#   UBLKCP        TMA 1-D bulk copy global -> shared (cp.async.bulk..., RANSAC block staging)
#   SYNCS         mbarrier arrive / try_wait (completion of the bulk copies)
#   LDGSTS        cp.async global -> shared (value column of the onesweep tiles)
#   MATCH.ANY     warp-level match (digit / (leaf, digit) peer ranking)
#   DFMA/DADD/DMUL  float64 arithmetic (RANSAC exact path: separate DMUL / DADD, no contraction; keygen: DFMA only where fma() is written)
#   ATOMS/ATOMG/RED  shared / global atomics
# No tensor-core instructions (UTCMMA / HMMA) are expected: no stage of the path is a dense contraction.
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
LIB="$HERE/octreelib_b200/liboctreelib_b200.so"
OUT="$HERE/profiles/r02_sass_evidence.txt"
TMP="$(mktemp)"
cuobjdump -sass "$LIB" > "$TMP"
{
  echo "# cuobjdump -sass octreelib_b200/liboctreelib_b200.so  ($(cuobjdump --version | tail -1))"
  echo "# arch: $(grep -m1 -o 'sm_[0-9a-z]*' "$TMP")"
  echo "# kernel | instructions | UBLKCP | SYNCS | LDGSTS | MATCH.ANY | DFMA | DADD | DMUL | ATOMS | ATOMG+RED | UTCMMA/HMMA"
  awk '
    /Function :/ { if (name != "") print_row(); name=$3; n=0; ub=0; sy=0; ld=0; ma=0; df=0; da=0; dm=0; as=0; ag=0; tc=0; next }
    /^[[:space:]]*\/\*[0-9a-f]+\*\// {
      n++;
      if ($0 ~ /UBLKCP/) ub++; if ($0 ~ /SYNCS/) sy++; if ($0 ~ /LDGSTS/) ld++; if ($0 ~ /MATCH\.ANY/) ma++;
      if ($0 ~ /DFMA/) df++; if ($0 ~ /DADD/) da++; if ($0 ~ /DMUL/) dm++; if ($0 ~ /ATOMS/) as++;
      if ($0 ~ /ATOMG/ || $0 ~ / RED\./) ag++; if ($0 ~ /UTCMMA/ || $0 ~ /HMMA/) tc++;
    }
    function print_row() { printf "%s | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d\n", name, n, ub, sy, ld, ma, df, da, dm, as, ag, tc }
    END { if (name != "") print_row() }
  ' "$TMP" | while IFS='|' read -r name rest; do
      printf "%s |%s\n" "$(echo "$name" | c++filt | cut -c1-110)" "$rest"
  done | sort
  echo
  echo "# excerpt: the TMA bulk copy and its mbarrier wait in ransac_small_kernel<false>"
  awk '/Function : .*ransac_small_kernelILb0/ {on=1} on && /Function :/ && !/ransac_small_kernelILb0/ {on=0} on && (/UBLKCP/ || /SYNCS/)' "$TMP" | head -12
  echo
  echo "# excerpt: cp.async (LDGSTS) and MATCH.ANY in os_pass_kernel<unsigned int, 256, 16, 3>"
  awk '/Function : .*os_pass_kernelIjLi256ELi16ELi3E/ {on=1} on && /Function :/ && !/os_pass_kernelIjLi256ELi16ELi3E/ {on=0} on && (/LDGSTS/ || /MATCH/)' "$TMP" | head -12
} > "$OUT"
rm -f "$TMP"
echo "wrote $OUT ($(wc -l < "$OUT") lines)"
