set -x
python -m pytest tests -m gpu -q > gpurun_out/r2b_gputest.log 2>&1
python bench.py > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err
python tools/bench_paths.py > gpurun_out/r2b_paths.jsonl 2> gpurun_out/r2b_paths.err
( cd tools/proto && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/regbank regbank.cu && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ldsprobe ldsprobe.cu && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DKU=8 -DMODE=0 -o /tmp/cl2 ctab_loop2.cu ) > gpurun_out/r2b_probe_build.log 2>&1
( echo "# tools/proto/regbank.cu"; /tmp/regbank; echo; echo "# tools/proto/ldsprobe.cu"; /tmp/ldsprobe; echo; echo "# tools/proto/ctab_loop2.cu (-DKU=8 -DMODE=0)"; /tmp/cl2; echo; echo "# tools/fp64_micro.py"; python tools/fp64_micro.py ) > gpurun_out/r2b_fp64_operand_probe.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_raw.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mc_static_ctab_kernel -c 1 -s 3 -f -o gpurun_out/r2b_mc_static_ctab_kernel python tools/proto/r2_kernels.py mc_static > gpurun_out/r2b_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:mc_static_finish_kernel -c 1 -s 3 -f -o gpurun_out/r2b_mc_static_finish_kernel python tools/proto/r2_kernels.py mc_static >> gpurun_out/r2b_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
REPS=2 ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:"mc_static_ctab|mc_static_finish" -c 2 -s 4 --csv --log-file gpurun_out/r2b_fp64_ops.csv python tools/proto/r2_kernels.py mc_static > gpurun_out/r2b_fp64_ops.log 2>&1
