set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
TAG=${1:-r1e}
$NCU -k regex:pipeline_kernel -s 5 -c 1 -f -o gpurun_out/prof_q3_$TAG python profiles/run_shape.py q3 59986052 3 > gpurun_out/ncu_q3_$TAG.log 2>&1
$NCU -k regex:pipeline_kernel -s 5 -c 1 -f -o gpurun_out/prof_q3b_$TAG python profiles/run_shape.py q3bloom 59986052 3 > gpurun_out/ncu_q3b_$TAG.log 2>&1
$NCU -k regex:pipeline_kernel -s 4 -c 1 -f -o gpurun_out/prof_q3b_orders_$TAG python profiles/run_shape.py q3bloom 59986052 3 > gpurun_out/ncu_q3bo_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_q3*_$TAG.log
