#!/bin/bash
# Reproducible Nsight Compute captures of one training step (run on the GPU box, ONE GPU):
#     gpurun --timeout 1500 -- 'bash tools/ncu_capture.sh r02'
# then, here:  python tools/ncu_summarize.py gpurun_out/r02_top_raw.csv profiles/r02_traffic.json ; cp the csv files to profiles/.
# Every ncu run is preceded by the same command without ncu (exit code checked).  Numbers printed under ncu are never bench values.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline --batches 8"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
# (1) every launch with its device time: 15 steps (3 warm-up + 2 timed + 2 profiled + 2 x (2 + 2) host-fed); shares of a step, not absolutes
ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu1.log 2>&1
# (2) full counter set on the 36 fused / retrieval launches of the 4th step (3 x [attn fwd, attn bwd, ffn fwd, ffn bwd1, ffn bwd2,
#     2 lin_fwd, 2 lin_bwd] + 2 fm + 5 gemm_tma + retrieve fwd / bwd)
ncu --set full --clock-control none -k regex:'attn_tc|attn_h16|attn_mma|ffn_|lin_fwd|lin_bwd|fm_fwd|fm_bwd|gemm_tma|retrieve_fwd|retrieve_bwd' \
    --launch-skip 108 -c 36 -o $OUT/${TAG}_top $CMD > $OUT/${TAG}_ncu2.log 2>&1
ncu -i $OUT/${TAG}_top.ncu-rep --page raw --csv > $OUT/${TAG}_top_raw.csv 2>/dev/null
rm -f $OUT/${TAG}_top.ncu-rep
# (3) the fp32 mode's split-precision GEMMs (one step's worth), with source correlation for the dominant shape (SKIP_FP32=1 skips)
[ -n "$SKIP_FP32" ] && { ls -la $OUT/${TAG}_*; exit 0; }
CMD32="python bench.py --amp none --quick --steps 2 --warmup 3 --no-cpu-baseline --batches 8"
$CMD32 > $OUT/${TAG}_plain32.json 2> $OUT/${TAG}_plain32.err || { echo "fp32 plain run failed"; exit 1; }
ncu --set full --clock-control none -k regex:'gemm_tc' --launch-skip 200 -c 16 -o $OUT/${TAG}_gemm32 $CMD32 > $OUT/${TAG}_ncu3.log 2>&1
ncu -i $OUT/${TAG}_gemm32.ncu-rep --page raw --csv > $OUT/${TAG}_gemm32_raw.csv 2>/dev/null
rm -f $OUT/${TAG}_gemm32.ncu-rep          # gpurun merges at most 64 MiB back: only the CSV exports travel
ls -la $OUT/${TAG}_*
