timeout 300 python scripts/path_check.py 2>&1 | grep -v "^Import\|^Done"
for wt in 8 12 16 20 28; do echo "wt $wt"; LISA_WAIT_THRESH=$wt timeout 120 python scripts/path_check.py --skip-parity --no-wavefront 2>&1 | tail -1; done
for wt in 12 16 20 24; do echo "variant mb6 wt $wt"; LISA_RT_LIB=lisa_b200/variants/liblisa_rt_mb6.so LISA_WAIT_THRESH=$wt timeout 120 python scripts/path_check.py --skip-parity --no-wavefront 2>&1 | tail -1; done
