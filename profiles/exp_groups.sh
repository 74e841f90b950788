# experiment: groups (128-thread MMA tiles) per CTA -- more warps (3) vs more L1 (2, 1); binned queries
cd /root/repo
for G in 1 2 3; do
  NGLOD_EXTRA_NVCC_FLAGS="-DNGLOD_TRACE_GROUPS=3 -DNGLOD_FWD_GROUPS=$G" python nglod_b200/build.py --force > /dev/null
  echo "== fwd groups $G"
  ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/l.csv -k regex:"sdf_forward_tc" -s 5 -c 2 python profiles/perf_fwd.py > /dev/null 2>&1; python profiles/ncu_csv.py gpurun_out/l.csv
done
