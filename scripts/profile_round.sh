set -x
python bench.py > gpurun_out/r4_bench_default.json 2> gpurun_out/r4_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4_bench_reference.json 2> gpurun_out/r4_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r4_launches_10m.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r4_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spec_grow_kernel -s 40 -c 1 -o gpurun_out/r4_grow python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r4_ncu_grow.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_groups_kernel -s 1 -c 1 -o gpurun_out/r4_knn_groups python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r4_ncu_knn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spec_sweep_kernel -s 40 -c 1 -o gpurun_out/r4_sweep python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r4_ncu_sweep.log 2>&1
tail -c 600 gpurun_out/r4_bench_default.json; tail -c 400 gpurun_out/r4_bench_reference.json
