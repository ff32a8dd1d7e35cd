set -x
timeout 200 python scripts/path_check.py --no-wavefront 2>&1 | grep -v "^Import\|^Done" 
NCU=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_path -o gpurun_out/r01_k_path -f python scripts/prof_run.py 6 2000 2>&1 | tail -5
