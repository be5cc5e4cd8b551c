#!/bin/bash
# Per-kernel counts of the SASS mnemonics that tell a Blackwell-native kernel from a recompiled one (B200_PROFILING.md):
#   UTC*MMA = tcgen05.mma   UTMALDG/UTMASTG/UBLKCP = TMA   LDTM/STTM = tcgen05.ld/st   HMMA = mma.sync   LDGSTS = cp.async   SYNCS = mbarrier
# usage: tools/sass_summary.sh [lib.so] > profiles/rNN_sass_summary.txt      (no GPU needed)
LIB=${1:-kaldi-aslp_b200/libaslp_b200.so}
echo "# tensor-core / TMA / TMEM / legacy-MMA / cp.async / mbarrier instruction counts per kernel of $LIB (cuobjdump -sass, kernels with none of them omitted)"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function : /{fn=$3}
/ UTC[A-Z]*MMA/{u[fn]++; next}
/UTMALDG|UTMASTG|UBLKCP/{t[fn]++}
/ LDTM| STTM/{l[fn]++}
/ HMMA/{h[fn]++}
/LDGSTS/{g[fn]++}
/SYNCS/{s[fn]++}
END{printf "%-96s %8s %6s %10s %6s %7s %6s\n","kernel (mangled, truncated)","UTC*MMA","TMA","LDTM/STTM","HMMA","LDGSTS","SYNCS"; for(f in u)k[f]=1; for(f in t)k[f]=1; for(f in l)k[f]=1; for(f in h)k[f]=1; for(f in g)k[f]=1; for (f in k) printf "%-96s %8d %6d %10d %6d %7d %6d\n", substr(f,1,96), u[f], t[f], l[f], h[f], g[f], s[f]}' | sort
