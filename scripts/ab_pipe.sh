# A/B of the pipelined host loop of the grower (BSEG_PIPE) after the whole GPU suite
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -5) > gpurun_out/pipe_pytest.log
tail -3 gpurun_out/pipe_pytest.log
for pp in 1 0; do
  BSEG_PIPE=$pp BSEG_DEBUG=1 timeout 200 python bench.py --workload C5 --steps 2 --warmup 1 --no-cpu --no-io > gpurun_out/pipe${pp}_c5.json 2> gpurun_out/pipe${pp}_c5.err
  BSEG_PIPE=$pp BSEG_DEBUG=1 timeout 100 python bench.py --workload C2 --steps 3 --warmup 2 --no-cpu --no-io > gpurun_out/pipe${pp}_c2.json 2> gpurun_out/pipe${pp}_c2.err
done
for f in gpurun_out/pipe1_c5 gpurun_out/pipe0_c5 gpurun_out/pipe1_c2 gpurun_out/pipe0_c2; do
  echo "== $f"; grep -o '"ms_per_step": [0-9.]*' $f.json | tr '\n' ' '; grep -o '"grow": {"ms": [0-9.]*' $f.json; grep -o '"grow": {"steps[^}]*}' $f.json; grep "rounds\|sweeper: front" $f.err | tail -2
done
