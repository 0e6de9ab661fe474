# (1) the labels of a C5 pass with and without the concurrent parts of the round loop (config.labels_checksum must agree);
# (2) DRAM bytes of every launch of one C2 pass (roofline.traffic)
mkdir -p gpurun_out
BSEG_BG=0 BSEG_PIPE=0 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-io > gpurun_out/r2f_chk_plain_c5.json 2> gpurun_out/r2f_chk_plain_c5.err
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-io > gpurun_out/r2f_chk_default_c5.json 2> gpurun_out/r2f_chk_default_c5.err
grep -o '"labels_checksum": [^]]*]' gpurun_out/r2f_chk_plain_c5.json gpurun_out/r2f_chk_default_c5.json
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2f_traffic_c2.csv \
    python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu --no-io > gpurun_out/r2f_traffic.log 2>&1
wc -l gpurun_out/r2f_traffic_c2.csv
