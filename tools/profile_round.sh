set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --steps 2 --warmup 3 --batches-per-step 1 --no-full --no-cpu-baseline > gpurun_out/r02k_launches_bench.log 2>&1
MB_BATCHES=3 ncu --set full --clock-control none -s 24 -c 12 --csv --page raw --log-file gpurun_out/r02k_kernels_full.csv python tools/mini_bench.py > gpurun_out/r02k_full.log 2>&1
wc -l gpurun_out/r02k_launches.csv gpurun_out/r02k_kernels_full.csv
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err; tail -c 600 gpurun_out/bench_r02k.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02k_reference.json 2> gpurun_out/bench_r02k_reference.err; tail -c 400 gpurun_out/bench_r02k_reference.json
