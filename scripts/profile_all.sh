#!/bin/sh
# The profile pass behind profiles/<TAG>_*: run on the GPU box from the repo root (gpurun -- 'sh scripts/profile_all.sh r03 [parts]').
# parts (default: all): pool knot soup launches collapse build sanitize
#   pool      ncu --set full of ONE k_pool launch of 50 spp on BASELINE configs[1] (what bench.py's roofline object reads),
#             + per-stage / per-region shares of its instructions (needs -lineinfo, which the Makefile passes)
#   knot      same on the C3 scene (871k-triangle knot, 16 spp);   soup: on the C4 10M-triangle soup (4 spp, deep flavour)
#   launches  the launch list of a short bench.py run (shares per kernel)
#   collapse  ncu --set full of k_collapse8 (a wide level) at 10M triangles;  build: stage times at 1M / 10M / 100M
#   sanitize  compute-sanitizer memcheck + racecheck on scripts/sanitize_run.py
# Numbers printed under a profiler are never bench values.
TAG=${1:-r03}; shift
PARTS=${*:-pool knot soup launches collapse build sanitize}
O=gpurun_out; mkdir -p $O out
NCU="ncu --set full --import-source on --clock-control none"
has() { case " $PARTS " in *" $1 "*) return 0;; esac; return 1; }
if has knot; then python assets/gen_knot.py out/knot.obj > /dev/null 2>&1; fi
if has pool; then
  NCU=1 $NCU -k regex:k_pool -c 1 -f -o $O/${TAG}_k_pool python scripts/prof_run.py 50 > $O/${TAG}_k_pool.log 2>&1
  python scripts/ncu_summary.py $O/${TAG}_k_pool.ncu-rep $O/${TAG}_k_pool.json 50 "ncu --set full --import-source on --clock-control none -k regex:k_pool -c 1 python scripts/prof_run.py 50 (BASELINE configs[1]: Cornell 2000x2000, 7 bounces, one launch of 50 spp)" > /dev/null
  { echo "# k_pool<wide8, flat, shallow>, Cornell 2000x2000, one launch of 50 spp ($O/${TAG}_k_pool.ncu-rep)"; echo "# by stage (scripts/ncu_stages.py)"; python scripts/ncu_stages.py $O/${TAG}_k_pool.ncu-rep; echo; echo "# by 10-line region (scripts/ncu_regions.py)"; python scripts/ncu_regions.py $O/${TAG}_k_pool.ncu-rep | head -60; } > $O/${TAG}_k_pool_regions.txt 2>&1
fi
if has knot; then
  NCU=1 $NCU -k regex:k_pool -c 1 -f -o $O/${TAG}_k_pool_knot python scripts/prof_scene.py knot 16 > $O/${TAG}_k_pool_knot.log 2>&1
  python scripts/ncu_summary.py $O/${TAG}_k_pool_knot.ncu-rep $O/${TAG}_k_pool_knot.json 16 "ncu ... -k regex:k_pool -c 1 python scripts/prof_scene.py knot 16 (C3: 871,218 triangles, 1920x1080, 12 bounces, one launch of 16 spp)" > /dev/null
fi
if has soup; then
  NCU=1 $NCU -k regex:k_pool -c 1 -f -o $O/${TAG}_k_pool_soup10m python scripts/prof_scene.py soup 10000000 4 > $O/${TAG}_k_pool_soup10m.log 2>&1
  python scripts/ncu_summary.py $O/${TAG}_k_pool_soup10m.ncu-rep $O/${TAG}_k_pool_soup10m.json 4 "ncu ... -k regex:k_pool -c 1 python scripts/prof_scene.py soup 10000000 4 (C4: 10M-triangle soup, 1024x1024, 7 bounces, one launch of 4 spp; deep flavour, leaves by the surface-area rule)" > /dev/null
  { echo "# k_pool<wide8, flat, deep>, 10M-triangle soup 1024x1024, one launch of 4 spp"; python scripts/ncu_stages.py $O/${TAG}_k_pool_soup10m.ncu-rep; } > $O/${TAG}_k_pool_soup10m_regions.txt 2>&1
  rm -f $O/${TAG}_k_pool_soup10m.ncu-rep
fi
if has launches; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/${TAG}_launches.log 2>&1
  python scripts/ncu_launch_shares.py $O/${TAG}_launches.csv "# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e" "# Serialised launch times under the profiler: compare SHARES.  (profiles/${TAG}_launches.csv is the raw list.)" > $O/${TAG}_launch_shares.txt 2>&1
fi
if has collapse; then
  $NCU -k regex:k_collapse8 -s 8 -c 1 -f -o $O/${TAG}_k_collapse8 python scripts/prof_build.py 10000000 > $O/${TAG}_k_collapse8.log 2>&1
  python scripts/ncu_summary.py $O/${TAG}_k_collapse8.ncu-rep $O/${TAG}_k_collapse8.json 0 "ncu --set full ... -k regex:k_collapse8 -s 8 -c 1 python scripts/prof_build.py 10000000 (10M-triangle soup, the 9th launch = a wide level; with the surface-area rule for leaves)" > /dev/null
  rm -f $O/${TAG}_k_collapse8.ncu-rep
fi
if has build; then
  python scripts/build_bench.py 1000000 10000000 100000000 > $O/${TAG}_build_bench.jsonl 2> $O/${TAG}_build_bench.err
fi
if has sanitize; then
  { echo "# compute-sanitizer on python scripts/sanitize_run.py (every schedule, both k_pool flavours, flat and cone filters, PLOC look-back rounds, collapse with / without the surface-area rule, serialised BVH)"; echo "## memcheck"; compute-sanitizer --tool memcheck python scripts/sanitize_run.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|error" | head -20; echo "## racecheck"; compute-sanitizer --tool racecheck python scripts/sanitize_run.py 2>&1 | grep -E "RACECHECK SUMMARY|hazard" | head -20; } > $O/${TAG}_sanitizer.txt
fi
ls -la $O | grep ${TAG}_
