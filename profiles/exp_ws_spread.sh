cd /root/repo
NGLOD_EXTRA_NVCC_FLAGS="-DNGLOD_WS_TIMING" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
for n in 1048576 8388608; do
python profiles/exp_ws_spread.py $n > /tmp/spread.txt 2>&1
python - <<PY
import re
runs=open('/tmp/spread.txt').read().split('LAUNCH_END')
last=runs[2]
rows=[(int(m.group(1)),int(m.group(2)),int(m.group(3)),int(m.group(4))) for m in re.finditer(r'WSCTA (\d+) sm (\d+) tiles (\d+) end_ns (\d+)', last)]
t0=min(r[3] for r in rows); ends=sorted((r[3]-t0)/1000 for r in rows)
import statistics
print("n=$n: CTAs", len(rows), "finish spread us: min 0, p10 %.1f, median %.1f, p90 %.1f, max %.1f" % (ends[len(ends)//10], statistics.median(ends), ends[9*len(ends)//10], ends[-1]))
by_tiles={}
for r in rows: by_tiles.setdefault(r[2],[]).append((r[3]-t0)/1000)
for k,v in sorted(by_tiles.items()): print("   tiles", k, "n", len(v), "median end %.1f us" % statistics.median(v))
PY
done
python nglod_b200/build.py --force > /dev/null
