# sweep: producer warps of the warp-specialised forward vs the serial-group kernel
cd /root/repo
for f in "-DNGLOD_FWD_WS=0" "-DNGLOD_WS_PRODUCERS=11" "-DNGLOD_WS_PRODUCERS=15" "-DNGLOD_WS_PRODUCERS=19" "-DNGLOD_WS_PRODUCERS=23"; do
  echo "== flags: $f"
  NGLOD_EXTRA_NVCC_FLAGS="$f" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
  timeout -s KILL 120 python profiles/exp_ws.py 2>&1 | tail -14
done
