# per-round fixed costs (RCH = 4, two launches folded, slot assignment without scan kernels): parity, then C5 / C2 and a smaller scout window
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_grow.py tests/test_gpu_scale.py tests/test_gpu_golden.py tests/test_gpu_abi2.py tests/test_gpu_tile.py -m gpu -x -q --timeout 200 2>&1 | tail -5) > gpurun_out/round_pytest.log
tail -3 gpurun_out/round_pytest.log
BSEG_DEBUG=1 timeout 200 python bench.py --workload C5 --steps 2 --warmup 1 --no-cpu --no-io > gpurun_out/round_c5.json 2> gpurun_out/round_c5.err
BSEG_WINDOW=131072 BSEG_DEBUG=1 timeout 200 python bench.py --workload C5 --steps 2 --warmup 1 --no-cpu --no-io > gpurun_out/round_c5_w128k.json 2> gpurun_out/round_c5_w128k.err
BSEG_DEBUG=1 timeout 100 python bench.py --workload C2 --steps 3 --warmup 2 --no-cpu --no-io > gpurun_out/round_c2.json 2> gpurun_out/round_c2.err
for f in gpurun_out/round_c5 gpurun_out/round_c5_w128k gpurun_out/round_c2; do
  echo "== $f"; grep -o '"ms_per_step": [0-9.]*' $f.json | tr '\n' ' '; grep -o '"grow": {"ms": [0-9.]*' $f.json; grep -o '"grow": {"steps[^}]*}' $f.json; grep "rounds\|sweeper: front" $f.err | tail -2
done
