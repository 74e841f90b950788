# last GPU call of round 2: the whole -m gpu suite, the frame timing, the default bench line, then the sparse re-fill knob
cd /root/repo
mkdir -p gpurun_out
( time timeout -s KILL 170 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/final_tests.txt 2>&1
timeout -s KILL 60 python profiles/perf_trace.py 2>&1 | tail -1 > gpurun_out/final_perf.txt
( time timeout -s KILL 200 python bench.py > gpurun_out/bench_n1_final.log 2> gpurun_out/bench_n1_final.err ) 2> gpurun_out/bench_n1_final.time
NGLOD_EXTRA_NVCC_FLAGS="-DNGLOD_SPC_REFILL_MIN=8" python nglod_b200/build.py --force > /dev/null || echo BUILD FAILED
timeout -s KILL 120 python profiles/perf_spc.py 2>&1 | tail -5 | cut -c1-95 > gpurun_out/final_spc_min8.txt
cat gpurun_out/final_tests.txt gpurun_out/final_perf.txt gpurun_out/bench_n1_final.time gpurun_out/final_spc_min8.txt
