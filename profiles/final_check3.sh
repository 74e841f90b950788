# final tree: the -m gpu suite and __graft_entry__.smoke()
cd /root/repo
mkdir -p gpurun_out
( timeout -s KILL 80 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/final3_tests.txt 2>&1
( timeout -s KILL 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/final3_smoke.txt 2>&1
cat gpurun_out/final3_tests.txt gpurun_out/final3_smoke.txt
