set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
python profiles/run_shape.py q1 30000000 4 > gpurun_out/plain_q1_r1d.log 2>&1 && \
$NCU -k regex:pipeline_kernel -s 2 -c 1 -f -o gpurun_out/prof_q1_r1d python profiles/run_shape.py q1 30000000 4 > gpurun_out/ncu_q1_r1d.log 2>&1
python profiles/run_shape.py q3 59986052 3 > gpurun_out/plain_q3_r1d.log 2>&1 && \
$NCU -k 'regex:.*pipeline_kernel<1, 0, 1, 1.*' -s 1 -c 1 -f -o gpurun_out/prof_q3_r1d python profiles/run_shape.py q3 59986052 3 > gpurun_out/ncu_q3_r1d.log 2>&1
python profiles/run_shape.py q3bloom 59986052 3 > gpurun_out/plain_q3b_r1d.log 2>&1 && \
$NCU -k 'regex:.*pipeline_kernel<1, 0, 1, 1.*' -s 1 -c 1 -f -o gpurun_out/prof_q3b_r1d python profiles/run_shape.py q3bloom 59986052 3 > gpurun_out/ncu_q3b_r1d.log 2>&1
python profiles/run_shape.py bloom 64000000 4 > gpurun_out/plain_bloom_r1d.log 2>&1 && \
$NCU -k regex:bloom_probe -s 2 -c 1 -f -o gpurun_out/prof_bloom_r1d python profiles/run_shape.py bloom 64000000 4 > gpurun_out/ncu_bloom_r1d.log 2>&1
cat gpurun_out/plain_*_r1d.log
tail -3 gpurun_out/ncu_*_r1d.log
