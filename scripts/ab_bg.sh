# A/B of the background slices (BSEG_BG) on one B200: parity tests first, then C5 / C2 / C3 passes with the grower's debug lines
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_grow.py tests/test_gpu_scale.py tests/test_gpu_golden.py -m gpu -x -q --timeout 120 2>&1 | tail -5) > gpurun_out/bg_pytest.log
tail -3 gpurun_out/bg_pytest.log
for bg in 1 0; do
  BSEG_BG=$bg BSEG_DEBUG=1 timeout 200 python bench.py --workload C5 --steps 2 --warmup 1 --no-cpu --no-io > gpurun_out/bg${bg}_c5.json 2> gpurun_out/bg${bg}_c5.err
  BSEG_BG=$bg BSEG_DEBUG=1 timeout 100 python bench.py --workload C2 --steps 3 --warmup 2 --no-cpu --no-io > gpurun_out/bg${bg}_c2.json 2> gpurun_out/bg${bg}_c2.err
done
BSEG_BG=1 BSEG_DEBUG=1 timeout 200 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu --no-io > gpurun_out/bg1_c3.json 2> gpurun_out/bg1_c3.err
for f in gpurun_out/bg1_c5 gpurun_out/bg0_c5 gpurun_out/bg1_c2 gpurun_out/bg0_c2 gpurun_out/bg1_c3; do
  echo "== $f"; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"grow": {"steps[^}]*}' $f.json; grep "rounds\|sweeper: front" $f.err | tail -2
done
