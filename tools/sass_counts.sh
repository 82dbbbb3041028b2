#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove which pipes a kernel uses (tcgen05 MMA / TMEM / TMA / packed fp32 FMA).
#   tools/sass_counts.sh [lib] > profiles/sass_r2.txt          (no GPU needed)
LIB=${1:-inbed_pose_estimation_b200/libsmplify_b200.so}
echo "# cuobjdump -sass $LIB : instructions per kernel"
echo "# UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st (TMEM), UTMALDG / UTMASTG = TMA load / store,"
echo "# UTCBAR = tcgen05.commit, FFMA2 = packed fp32 FMA, FFMA = scalar fp32 FMA"
cuobjdump -sass "$LIB" | awk '
/Function :/ { name = $3; order[++n] = name }
/^[ \t]+\/\*[0-9a-f]+\*\// {
    total[name]++
    if ($0 ~ /UTCHMMA|UTCQMMA|UTCMMA/) mma[name]++
    if ($0 ~ /LDTM/) ldtm[name]++
    if ($0 ~ /STTM/) sttm[name]++
    if ($0 ~ /UTMALDG/) tmal[name]++
    if ($0 ~ /UTMASTG/) tmas[name]++
    if ($0 ~ /UTCBAR/) commit[name]++
    if ($0 ~ /FFMA2/) ffma2[name]++
    else if ($0 ~ /FFMA/) ffma[name]++
}
END {
    printf "%-90s %8s %7s %6s %6s %8s %8s %7s %7s %7s\n", "kernel", "instr", "UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "FFMA2", "FFMA"
    for (i = 1; i <= n; i++) { k = order[i]; printf "%-90s %8d %7d %6d %6d %8d %8d %7d %7d %7d\n", k, total[k], mma[k], ldtm[k], sttm[k], tmal[k], tmas[k], commit[k], ffma2[k], ffma[k] }
}' | c++filt | cut -c1-200
