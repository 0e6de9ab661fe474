# Round-end measurement pass on one B200 (run under gpurun from the repo root).  "bench": the GPU suite and the bench lines;
# "ncu": launch list and full captures on the 10 M-point workload (a launch list of the 200 M tile is 40 k launches).
set -x
mkdir -p gpurun_out
if [ "$1" != "ncu" ]; then
(timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -5) > gpurun_out/r2f_pytest.log
timeout 400 python bench.py > gpurun_out/r2f_bench_c5_200m.json 2> gpurun_out/r2f_bench_c5_200m.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_bench_reference_arm.err
timeout 100 python bench.py --workload C2 --steps 5 --warmup 3 > gpurun_out/r2f_bench_c2_10m.json 2> gpurun_out/r2f_bench_c2_10m.err
timeout 100 python bench.py --workload C1 --steps 5 --warmup 3 > gpurun_out/r2f_bench_c1_1m.json 2> gpurun_out/r2f_bench_c1_1m.err
BSEG_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-io > gpurun_out/r2f_dbg_c5.json 2> gpurun_out/r2f_dbg_c5.err
timeout 200 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu --no-io > gpurun_out/r2f_bench_c3_50m.json 2> gpurun_out/r2f_bench_c3_50m.err
tail -3 gpurun_out/r2f_pytest.log; tail -6 gpurun_out/r2f_dbg_c5.err
fi
if [ "$1" != "bench" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches_c2_10m.csv \
    python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu --no-io > gpurun_out/r2f_ncu_launch.log 2>&1
for k in spec_sweep_kernel:30 spec_grow_kernel:30 knn_groups_kernel:0; do
  name=${k%%:*}; skip=${k##*:}
  timeout 200 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:$name -s $skip -c 2 \
      -o gpurun_out/r2f_$name python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu --no-io > gpurun_out/r2f_ncu_$name.log 2>&1
done
fi
ls -la gpurun_out | tail -14
