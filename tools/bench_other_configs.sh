# usage: bash tools/bench_other_configs.sh TAG   -- one GPU call: the multi-context probe and the bench lines of configs[0] and configs[4]
TAG=${1:-r02n}
python tools/multi_ctx_probe.py > gpurun_out/${TAG}_multi_ctx.json 2>&1; cat gpurun_out/${TAG}_multi_ctx.json
python bench.py --config a --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_a.json 2> gpurun_out/bench_${TAG}_a.err; tail -c 300 gpurun_out/bench_${TAG}_a.json
python bench.py --config e --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_e.json 2> gpurun_out/bench_${TAG}_e.err; tail -c 300 gpurun_out/bench_${TAG}_e.json
