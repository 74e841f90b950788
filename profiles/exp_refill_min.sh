cd /root/repo
for f in "" "-DNGLOD_TRACE_REFILL_MIN=4" "-DNGLOD_TRACE_REFILL_MIN=8" "-DNGLOD_TRACE_REFILL_MIN=16"; do
  echo "== flags: $f"
  NGLOD_EXTRA_NVCC_FLAGS="$f" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
  timeout -s KILL 120 python profiles/perf_trace.py 2>&1 | tail -1 | cut -c1-60
  timeout -s KILL 120 python profiles/perf_e2e.py 2>&1 | tail -1 | cut -c1-70
done
python nglod_b200/build.py --force > /dev/null
