# sweep: load depth of the warp-specialised forward (NGLOD_FWD_WS=0 = the serial-group kernel)
cd /root/repo
for f in "$@"; do
  echo "== flags: $f"
  NGLOD_EXTRA_NVCC_FLAGS="$f" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
  timeout -s KILL 120 python profiles/exp_ws.py 2>&1 | tail -14
done
python nglod_b200/build.py --force > /dev/null
