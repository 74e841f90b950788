cd /root/repo
for f in "-DNGLOD_TRACE_REFILL_MIN=6" "-DNGLOD_TRACE_REFILL_MIN=12"; do
  echo "== flags: $f"
  NGLOD_EXTRA_NVCC_FLAGS="$f" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
  timeout -s KILL 60 python profiles/perf_trace.py 2>&1 | tail -1 | cut -c1-60
done
