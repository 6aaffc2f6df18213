#!/usr/bin/env bash
# development helper (GPU box): launch list + one `ncu --set full` capture of every big kernel of one step, for the bench line (n1)
# and for one GPU's chunk of configs[2] (multi).  Each only after the same command ran clean without ncu.
set -u
BIG='tile_scatter|reduce_moments|onesweep|voxel_stats|tile_digits|tile_offsets|p2v_lookup|digit_hist|count_heads|p2v_fill|minmax_partial|scatter_rows'
for W in configs1 configs2; do
  TAG=r2_$W
  python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_plain.json 2> gpurun_out/bench_${TAG}_plain.err || { echo "plain bench failed ($W)"; tail -5 gpurun_out/bench_${TAG}_plain.err; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
  # one step's worth of the big kernels after the 3 warm-up steps (10-14 per step)
  ncu --set full --clock-control none --import-source on -k regex:"$BIG" -s 45 -c 16 -f -o gpurun_out/prof_${TAG} \
      python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_full_${TAG}.log
  # the reports are too big to travel back (64 MiB limit): summarise them here, keep the CSVs
  KEY=n1; [ "$W" = configs2 ] && KEY=multi
  python tools/ncu_summary.py gpurun_out/prof_${TAG}.ncu-rep --csv gpurun_out/ncu_full_${TAG}.csv > gpurun_out/ncu_full_${TAG}.txt
  VAPC_NCU_TRAFFIC_OUT=gpurun_out/r2_ncu_traffic.json python tools/ncu_traffic.py $KEY gpurun_out/prof_${TAG}.ncu-rep > /dev/null
  for K in tile_scatter reduce_moments onesweep voxel_stats; do
    python tools/ncu_lines.py gpurun_out/prof_${TAG}.ncu-rep $K 0 25 > gpurun_out/ncu_lines_${TAG}_$K.txt 2>&1
  done
  rm -f gpurun_out/prof_${TAG}.ncu-rep
done
ls -la gpurun_out/
