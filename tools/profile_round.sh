#!/bin/bash
# dev tool (runs ON the GPU box via gpurun): the ncu captures profiles/ is built from.
# Every capture follows a plain run of the same command that exited 0.
tag=${1:-r02}
out=gpurun_out
ncu_full() {   # cfg skip count
  python tools/stage_time_cfg.py $1 1 2 > $out/${tag}_plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  cat $out/${tag}_plain_$1.log | tail -1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_ --launch-skip $2 --launch-count $3 \
      -f -o $out/${tag}_$1 python tools/stage_time_cfg.py $1 1 2 > $out/${tag}_ncu_$1.log 2>&1
  ncu -i $out/${tag}_$1.ncu-rep --page raw --csv > $out/${tag}_raw_$1.csv 2>/dev/null
}
ncu_full C2 10 5
ncu_full C3 12 6
ncu_full C4 12 6
rm -f $out/${tag}_C3.ncu-rep $out/${tag}_C4.ncu-rep       # the CSV exports travel back, the C2 report too (source page)
export FLAKE_BENCH_SKIP_OTHERS=1 FLAKE_BENCH_E2E_TRACKS=16
python bench.py --steps 2 --warmup 1 > $out/${tag}_bench_for_launches.json 2> $out/${tag}_bench_for_launches.err || echo "bench failed"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 450 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 > $out/${tag}_ncu_launches.log 2>&1
# (compute-sanitizer is closed on this GPU pool: racecheck / memcheck runs are refused; gpurun_out/r02_racecheck.log)
