# experiment: groups per CTA / pipeline depth for the single-grid kernels; gather-only and decoder-only timings
cd /root/repo
run() {
  NGLOD_EXTRA_NVCC_FLAGS="$1" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
  echo "== flags: $1"
  python profiles/perf_half.py 2>&1 | grep -E "^tc " | grep "summed=True"
}
for f in "$@"; do run "$f"; done
python nglod_b200/build.py --force > /dev/null
