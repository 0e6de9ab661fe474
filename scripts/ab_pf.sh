# A/B of the sweeper's L2 warmer (BSEG_PF), with the 32-bit kNN selection and the scout's alive filter on
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_knn.py tests/test_gpu_grow.py tests/test_gpu_scale.py tests/test_gpu_golden.py tests/test_gpu_abi2.py -m gpu -x -q --timeout 120 2>&1 | tail -5) > gpurun_out/pf_pytest.log
tail -3 gpurun_out/pf_pytest.log
for pf in 1 0; do
  BSEG_PF=$pf BSEG_DEBUG=1 timeout 200 python bench.py --workload C5 --steps 2 --warmup 1 --no-cpu --no-io > gpurun_out/pf${pf}_c5.json 2> gpurun_out/pf${pf}_c5.err
done
BSEG_DEBUG=1 timeout 100 python bench.py --workload C2 --steps 3 --warmup 2 --no-cpu --no-io > gpurun_out/pf1_c2.json 2> gpurun_out/pf1_c2.err
BSEG_DEBUG=1 timeout 200 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu --no-io > gpurun_out/pf1_c3.json 2> gpurun_out/pf1_c3.err
for f in gpurun_out/pf1_c5 gpurun_out/pf0_c5 gpurun_out/pf1_c2 gpurun_out/pf1_c3; do
  echo "== $f"; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"knn": {"ms": [0-9.]*' $f.json; grep -o '"knn_fallback": {"ms": [0-9.]*' $f.json; grep -o '"grow": {"steps[^}]*}' $f.json; grep "rounds\|sweeper: front" $f.err | tail -2
done
