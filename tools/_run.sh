for d in 0 2; do
VEL_MATCH_DEBUG_NOLOAD=$d timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum --clock-control none -k regex:"knn2_hamming_tc" -s 6 -c 1 python tools/match_check.py 2>&1 | grep -E "duration|issue_active|inst_exec" | tail -3
done
