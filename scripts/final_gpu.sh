set -x
mkdir -p gpurun_out/final
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err; tail -c 600 gpurun_out/final/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2>&1; tail -c 400 gpurun_out/final/bench_reference.json
python bench.py --nw 4096 > gpurun_out/final/bench_n1_4096.json 2>&1
python bench.py --nw 512 --beads 32 --no-cpu-baseline > gpurun_out/final/bench_beads.json 2>&1
python bench.py --nw 4096 --ions 64 > gpurun_out/final/bench_nacl.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/final/launches_nacl.csv python bench.py --nw 4096 --ions 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_nacl.log 2>&1
ls -la gpurun_out/final
