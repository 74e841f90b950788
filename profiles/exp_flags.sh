# A/B of build flags on the bench frame / forward / backward (profiles/perf_trace.py)
cd /root/repo
for f in "$@"; do
  echo "== flags: $f"
  NGLOD_EXTRA_NVCC_FLAGS="$f" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
  timeout -s KILL 120 python profiles/perf_trace.py 2>&1 | tail -1
done
python nglod_b200/build.py --force > /dev/null
