# the last check of the round: parity of the grower after the "no steals beside the sweeper" rule, and the C5 labels checksum
mkdir -p gpurun_out
(timeout 100 python -m pytest tests/test_gpu_grow.py tests/test_gpu_scale.py tests/test_gpu_golden.py -m gpu -x -q --timeout 90 2>&1 | tail -3) > gpurun_out/last_pytest.log &
BSEG_DEBUG=1 timeout 140 python bench.py --steps 1 --warmup 1 --no-cpu --no-io > gpurun_out/last_c5.json 2> gpurun_out/last_c5.err
wait
tail -2 gpurun_out/last_pytest.log
grep -o '"labels_checksum": [^]]*]' gpurun_out/last_c5.json; grep -o '"ms_per_step": [0-9.]*' gpurun_out/last_c5.json | tr '\n' ' '; grep "rounds\|sweeps ended" gpurun_out/last_c5.err | tail -2
