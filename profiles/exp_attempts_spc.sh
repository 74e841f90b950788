cd /root/repo
for f in "-DNGLOD_SPC_REFILL_ATTEMPTS=1" "-DNGLOD_SPC_REFILL_ATTEMPTS=2" ""; do
  echo "== flags: $f"
  NGLOD_EXTRA_NVCC_FLAGS="$f" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
  timeout -s KILL 200 python profiles/perf_spc.py 2>&1 | tail -5 | cut -c1-95
done
python nglod_b200/build.py --force > /dev/null
