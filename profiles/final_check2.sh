# after NGLOD_SPC_REFILL_MIN = 8 became the default: the whole -m gpu suite again and the default bench line
cd /root/repo
mkdir -p gpurun_out
( time timeout -s KILL 170 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/final2_tests.txt 2>&1
( time timeout -s KILL 200 python bench.py > gpurun_out/bench_n1_final2.log 2> gpurun_out/bench_n1_final2.err ) 2> gpurun_out/bench_n1_final2.time
cat gpurun_out/final2_tests.txt gpurun_out/bench_n1_final2.time; tail -3 gpurun_out/bench_n1_final2.err
