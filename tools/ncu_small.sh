#!/bin/bash
# Source-level Nsight Compute capture of the step's mid-sized CUDA-core kernels (one launch each, 6th step):
#     gpurun --timeout 900 -- 'bash tools/ncu_small.sh r02ai'
# exports: <tag>_small_raw.csv (details) and <tag>_small_src.csv (per-SASS stall samples; read with tools/ncu_src.py)
TAG=${1:-r02}
OUT=gpurun_out
KERN=${KERN:-cat_bwd_kernel|cat_fwd_kernel|rowopt_catchup_bm|rowrms_bwd|fm_bwd_kernel|transpose_bf16|headproj_bwd|sc_reduce}
CMD="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline --batches 8"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:"$KERN" --launch-skip ${SKIP:-50} -c ${COUNT:-10} -o $OUT/${TAG}_small $CMD > $OUT/${TAG}_ncu_small.log 2>&1
ncu -i $OUT/${TAG}_small.ncu-rep --page raw --csv > $OUT/${TAG}_small_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_small.ncu-rep --page source --csv > $OUT/${TAG}_small_src.csv 2>/dev/null
rm -f $OUT/${TAG}_small.ncu-rep
ls -la $OUT/${TAG}_*
