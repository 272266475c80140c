#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove (or disprove) Blackwell-native code paths in the shipped library:
#   UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, UTMAPF = TMA prefetch, LDTM / STTM = tcgen05.ld / .st (TMEM),
#   UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = mma.sync (Ampere-style tensor path), UBLKCP = cp.async.bulk (non-tensor)
# Usage: tools/sass_opcodes.sh [lib] > profiles/r02_sass_opcodes.txt
LIB=${1:-time_r1_b200/libtimer1.so}
echo "# cuobjdump -sass $LIB ($(date -u +%F), $(cuobjdump --version | tail -1))"
echo "# kernel | UTCHMMA UTMALDG UTMASTG UTMAPF LDTM STTM UTCBAR SYNCS HMMA UBLKCP | total SASS instructions"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
  /Function :/ { if (name != "") emit(); name=$3; delete c; total=0; next }
  /^ +\/\*[0-9a-f]+\*\/ +[A-Z@!]/ {
    total++
    for (k in ops) if (index($0, " " ops[k])) c[ops[k]]++      # leading blank: "HMMA" must not count "UTCHMMA"
  }
  function emit() {
    printf "%s |", name
    for (i = 1; i <= n; i++) printf " %d", c[ops[i]] + 0
    printf " | %d\n", total
  }
  BEGIN { n = split("UTCHMMA UTMALDG UTMASTG UTMAPF LDTM STTM UTCBAR SYNCS HMMA UBLKCP", ops, " ") }
  END { if (name != "") emit() }
' | while IFS='|' read -r name counts total; do
  printf "%s |%s|%s\n" "$(echo "$name" | c++filt | sed -E 's/\(anonymous namespace\):://g; s/\(.*//' | cut -c1-90)" "$counts" "$total"
done | sort
