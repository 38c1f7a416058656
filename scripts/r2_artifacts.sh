#!/bin/bash
# Round-2 measurement batch on ONE B200 (run under gpurun): bench lines, launch list, ncu summaries, sanitizer runs.
# Everything lands in gpurun_out/r2_*; the summaries worth keeping are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --nw 4096 --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_4096.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --beads 32 --nw 512 --steps 20 --warmup 3 > gpurun_out/r2_bench_beads32x512.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --ions 64 --nw 4096 --steps 20 --warmup 3 > gpurun_out/r2_bench_nacl.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --flags 0 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_flags0.json 2>> gpurun_out/r2_bench_n1.err
# launch list of one step (cold-cache, serialised: the SHARES must agree with the bench phases, not the absolute times)
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 45 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
# full captures of the four cluster kernels
for k in k_trimers k_trimers_gather k_dimers k_dimers_gather; do
  tag=r2_$k;
  ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 2 -c 1 -f -o gpurun_out/$tag python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/$tag.log 2>&1
  ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv 2>/dev/null
  ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/$tag.src.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/$tag.raw.csv gpurun_out/$tag.src.csv > gpurun_out/${tag}_ncu.txt 2>&1
  python scripts/ncu_compact.py gpurun_out/$tag.src.csv gpurun_out/$tag.pc.csv.gz
  rm -f gpurun_out/$tag.ncu-rep gpurun_out/$tag.src.csv
done
# sanitizer: memcheck + racecheck of the whole path on a small box (the scatter-free gather kernels and the staged record stores)
compute-sanitizer --tool memcheck python scripts/gpu_check.py 64 > gpurun_out/r2_memcheck.txt 2>&1
compute-sanitizer --tool racecheck python scripts/gpu_check.py 64 > gpurun_out/r2_racecheck.txt 2>&1
tail -3 gpurun_out/r2_memcheck.txt gpurun_out/r2_racecheck.txt
ls -la gpurun_out | head -40
