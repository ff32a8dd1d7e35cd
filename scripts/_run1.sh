timeout 300 python scripts/full_c2.py 2>&1 | grep "^{" > gpurun_out/full_c2.json
timeout 300 python scripts/scale_bench.py soup 1000000 4 2>&1 | grep "^{" > gpurun_out/soup1m.json
timeout 600 python scripts/scale_bench.py soup 10000000 2 2>&1 | grep "^{" > gpurun_out/soup10m.json
cat gpurun_out/full_c2.json gpurun_out/soup1m.json gpurun_out/soup10m.json
