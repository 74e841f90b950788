# ncu --set full of the tracer kernel of the final build (re-fill batching in), summarised with ncu_keys.py
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 150 ncu --set full --clock-control none --import-source on -k regex:sphere_trace_kernel -s 2 -c 1 -f -o gpurun_out/r2_trace_final python profiles/prof_target.py > gpurun_out/r2_trace_final.log 2>&1
ncu -i gpurun_out/r2_trace_final.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_keys.py > gpurun_out/sphere_trace_final_r2.txt
head -12 gpurun_out/sphere_trace_final_r2.txt
