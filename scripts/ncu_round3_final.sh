OUT=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --vidor-videos 0 --no-graph --no-pipeline --modes="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r3_final_launches_vidvrd200.csv $B > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r3_final_launches_vidor50.csv $B --workload vidor --videos 50 > /dev/null 2>&1
python scripts/summarize_launches.py $OUT/r3_final_launches_vidvrd200.csv > $OUT/r3_final_launches_vidvrd200.summary.txt
python scripts/summarize_launches.py $OUT/r3_final_launches_vidor50.csv > $OUT/r3_final_launches_vidor50.summary.txt
head -14 $OUT/r3_final_launches_vidvrd200.summary.txt; head -16 $OUT/r3_final_launches_vidor50.summary.txt
