# phase timing of the tcgen05 backward (-DBT_TIMING): 2^20-query launches per flag set
cd /root/repo
for f in "$@"; do
  echo "== flags: $f"
  NGLOD_EXTRA_NVCC_FLAGS="-DBT_TIMING $f" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
  timeout -s KILL 120 python profiles/prof_bwd.py 2>&1 | tail -3
done
python nglod_b200/build.py --force > /dev/null
