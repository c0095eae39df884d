# usage: bash tools/profile_round.sh TAG   -- one GPU call: launch list, full capture, both bench arms of the current build
TAG=${1:-r02m}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_gputests.log; cat gpurun_out/${TAG}_gputests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --batches-per-step 1 --no-full --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
# direct launches (tools/mini_bench.py, use_graph=0) on ONE stream (SFE_NO_PYR_FORK=1): 12 kernels per batch, the third batch is captured
SFE_NO_PYR_FORK=1 MB_BATCHES=3 ncu --set full --clock-control none -s 24 -c 12 --csv --page raw --log-file gpurun_out/${TAG}_kernels_full.csv python tools/mini_bench.py > gpurun_out/${TAG}_full.log 2>&1
wc -l gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_kernels_full.csv
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 300 gpurun_out/bench_${TAG}.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -c 300 gpurun_out/bench_${TAG}_reference.json
python -c "import __graft_entry__ as g; g.smoke()"
