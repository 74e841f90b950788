# experiment: groups (128-thread MMA tiles) per CTA -- more warps (3) vs more L1 (2)
cd /root/repo
for G in 2 3; do
  NGLOD_EXTRA_NVCC_FLAGS="-DNGLOD_TRACE_GROUPS=$G -DNGLOD_FWD_GROUPS=$G" python nglod_b200/build.py --force > /dev/null
  echo "== groups $G"
  timeout 200 python profiles/perf_fwd.py 2>&1 | tail -1
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('tracer ms', j['ms_per_step'])"
done
