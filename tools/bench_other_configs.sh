python bench.py --config a --steps 10 --warmup 3 > gpurun_out/bench_r02k_a.json 2> gpurun_out/bench_r02k_a.err; tail -c 300 gpurun_out/bench_r02k_a.json
python bench.py --config e --steps 10 --warmup 3 > gpurun_out/bench_r02k_e.json 2> gpurun_out/bench_r02k_e.err; tail -c 300 gpurun_out/bench_r02k_e.json
