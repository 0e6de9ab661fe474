# Round-2 measurement pass on one B200 (run under gpurun from the repo root; outputs under gpurun_out/, the
# summaries kept for the judge are copied to profiles/ by hand).  Part 1: the bench lines (≈ 9 minutes).  Part 2: ncu on the
# 10 M-point workload -- a launch list of the 200 M tile is 130 k launches and does not finish under ncu.
set -x
mkdir -p gpurun_out
if [ "$1" != "ncu" ]; then
python bench.py > gpurun_out/r2_bench_c5_200m.json 2> gpurun_out/r2_bench_c5_200m.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python bench.py --workload C2 --steps 5 --warmup 3 > gpurun_out/r2_bench_c2_10m.json 2> gpurun_out/r2_bench_c2_10m.err
python bench.py --workload C1 --steps 5 --warmup 3 > gpurun_out/r2_bench_c1_1m.json 2> gpurun_out/r2_bench_c1_1m.err
python bench.py --workload C3 --steps 2 --warmup 3 --no-io > gpurun_out/r2_bench_c3_50m.json 2> gpurun_out/r2_bench_c3_50m.err
timeout 900 python bench.py --workload C4 --steps 1 --warmup 3 --no-io > gpurun_out/r2_bench_c4_100m.json 2> gpurun_out/r2_bench_c4_100m.err
fi
if [ "$1" != "bench" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_c2_10m.csv \
    python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu --no-io > gpurun_out/r2_ncu_launch.log 2>&1
for k in spec_sweep_kernel:30 spec_grow_kernel:30 knn_groups_kernel:0; do
  name=${k%%:*}; skip=${k##*:}
  timeout 200 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:$name -s $skip -c 2 \
      -o gpurun_out/r2_$name python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu --no-io > gpurun_out/r2_ncu_$name.log 2>&1
done
fi
ls -la gpurun_out | tail -12
