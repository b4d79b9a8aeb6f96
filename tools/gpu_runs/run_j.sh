mkdir -p gpurun_out/r2j
timeout 120 ./tools/mma_ts_microbench > gpurun_out/r2j/mma_ts.txt 2>&1; echo "ts rc=$?"

cat gpurun_out/r2j/mma_ts.txt
