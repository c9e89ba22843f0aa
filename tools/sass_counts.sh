#!/bin/bash
# Per-kernel counts of the SASS instructions that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md): run on the
# build host (no GPU needed).  usage: bash tools/sass_counts.sh > profiles/sass_counts.txt
LIB=${1:-boosted_detr_b200/libbdetr.so}
echo "# cuobjdump -sass $LIB  ($(git rev-parse --short HEAD 2>/dev/null), $(date -u +%FT%TZ))"
echo "# columns: UTCHMMA (tcgen05.mma) | LDTM (tcgen05.ld) | STTM (tcgen05.st) | UTMALDG (TMA load) | UTMASTG (TMA store) | UTMAREDG (TMA reduce) | UBLKCP (bulk copy) | SYNCS (mbarrier) | REDUX | FFMA2+FADD2 | MUFU.EX2"
cuobjdump -sass "$LIB" | awk '
/Function :/ { if (name != "") emit(); name=$3; for (k in c) delete c[k]; next }
{ if ($0 ~ /UTCHMMA/) c["a"]++; if ($0 ~ /LDTM/) c["b"]++; if ($0 ~ /STTM/) c["c"]++; if ($0 ~ /UTMALDG/) c["d"]++; if ($0 ~ /UTMASTG/) c["e"]++;
  if ($0 ~ /UTMAREDG/) c["f"]++; if ($0 ~ /UBLKCP/) c["g"]++; if ($0 ~ /SYNCS/) c["h"]++; if ($0 ~ /REDUX/) c["i"]++; if ($0 ~ /FFMA2|FADD2/) c["j"]++; if ($0 ~ /MUFU.EX2/) c["k"]++ }
function emit() { t=c["a"]+c["b"]+c["c"]+c["d"]+c["e"]+c["f"]+c["g"]+c["i"]+c["j"]; if (t>0) printf "%-110s %5d %5d %5d %5d %5d %5d %5d %5d %5d %5d %5d\n", name, c["a"], c["b"], c["c"], c["d"], c["e"], c["f"], c["g"], c["h"], c["i"], c["j"], c["k"] }
END { if (name != "") emit() }' | c++filt | sort
