#!/bin/bash
# Round-end evidence capture on one B200 (run under gpurun): every program first runs plain (must exit 0), then under ncu.
# Outputs land in gpurun_out/; scripts/ncu_summary.py / launch_summary.py turn them into profiles/*.md here.
# usage: capture_profiles.sh [1|2|all]   (two parts: gpurun brings back at most 64 MiB of gpurun_out/ per call)
set -u
PART=${1:-all}
O=gpurun_out
T="timeout 170"
NCU="ncu --set full --clock-control none --import-source on"
run() { echo "== $*"; "$@"; echo "   rc=$?"; }

if [ "$PART" = 1 ] || [ "$PART" = all ]; then
run $T python scripts/prof_fused.py 256 1 rows > $O/r2f_rows_plain.log 2>&1 && \
  run $T $NCU -k regex:conv_rows -c 12 -f -o $O/r2f_rows python scripts/prof_fused.py 256 1 rows > $O/r2f_rows_ncu.log 2>&1
run $T python scripts/prof_fused.py 256 1 flat > $O/r2f_flat_plain.log 2>&1 && \
  run $T $NCU -k regex:conv_flat -c 10 -f -o $O/r2f_flat python scripts/prof_fused.py 256 1 flat > $O/r2f_flat_ncu.log 2>&1
fi
if [ "$PART" = 2 ] || [ "$PART" = all ]; then
run $T python scripts/attn_bench.py 256 > $O/r2f_attn_plain.log 2>&1 && \
  run $T $NCU -k regex:attn_kernel -c 4 -f -o $O/r2f_attn python scripts/attn_bench.py 256 x > $O/r2f_attn_ncu.log 2>&1
run $T python scripts/gn_bwd_bench.py 32 > $O/r2f_gnbwd_plain.log 2>&1 && \
  run $T $NCU -k regex:gn_bwd16 -c 4 -f -o $O/r2f_gnbwd python scripts/gn_bwd_bench.py 32 > $O/r2f_gnbwd_ncu.log 2>&1
run $T python scripts/train_breakdown.py 32 fused16 > $O/r2f_tb_fused16.txt 2>&1 && \
  run $T $NCU -k regex:conv_wgrad --launch-skip 198 -c 8 -f -o $O/r2f_wgrad python scripts/train_breakdown.py 32 fused16 > $O/r2f_wgrad_ncu.log 2>&1
fi
if [ "$PART" = 1 ] || [ "$PART" = all ]; then
# launch lists (gpu__time_duration only)
run $T python scripts/one_eval.py 256 4 t > $O/r2f_one_eval_plain.log 2>&1 && \
  run $T ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2f_launches_eval_B256.csv \
      python scripts/one_eval.py 256 4 > $O/r2f_one_eval_ncu.log 2>&1
run $T python scripts/train_bench.py 32 5 > $O/r2f_train_plain.log 2>&1 && \
  run $T ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2f_launches_train_B32.csv \
      python scripts/train_bench.py 32 2 > $O/r2f_train_ncu.log 2>&1
python scripts/eval_breakdown.py 256 > $O/r2f_eval_breakdown.txt 2>&1
fi
if [ "$PART" = 3 ]; then      # training-side refresh only: weight-gradient kernel + the step's launch list
run $T python scripts/train_breakdown.py 32 fused16 > $O/r2f_tb_fused16.txt 2>&1 && \
  run $T $NCU -k regex:conv_wgrad --launch-skip 198 -c 8 -f -o $O/r2f_wgrad python scripts/train_breakdown.py 32 fused16 > $O/r2f_wgrad_ncu.log 2>&1
run $T python scripts/wgrad_bench.py > $O/r2f_wgrad_plain.log 2>&1
run $T python scripts/train_bench.py 32 5 > $O/r2f_train_plain.log 2>&1 && \
  run $T ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2f_launches_train_B32.csv \
      python scripts/train_bench.py 32 2 > $O/r2f_train_ncu.log 2>&1
fi
ls -la $O | grep r2f
