#!/bin/bash
# One GPU-box pass producing everything profiles/ needs for a round tag: tools/gpu_round.sh r02a
# smoke, GPU parity tests, bench (ours + reference arm), launch list, ncu --set full of the dominant kernels, sweep.
tag=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; cut -c1-300 gpurun_out/${tag}_bench_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_n1.err; cut -c1-300 gpurun_out/${tag}_bench_reference.json
for w in wideband multiradio refexact sc16; do
  python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 3 > gpurun_out/${tag}_bench_$w.json 2>> gpurun_out/${tag}_bench_n1.err; cut -c1-200 gpurun_out/${tag}_bench_$w.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu --no-others --e2e-steps 1 > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sense_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_sense_n1024 \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-others --e2e-steps 1 > gpurun_out/${tag}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sense_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_sense_n8192 \
  python bench.py --workload wideband --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu8192.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sense_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_sense_n2048 \
  python bench.py --workload multiradio --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu2048.log 2>&1
python tools/sweep.py --samples 1e9 > gpurun_out/${tag}_sweep.json 2> gpurun_out/${tag}_sweep.err; tail -c 1500 gpurun_out/${tag}_sweep.json
python tools/latency.py > gpurun_out/${tag}_latency.txt 2>&1; tail -4 gpurun_out/${tag}_latency.txt | cut -c1-200
python tools/many_radios.py --radios 256 --mode ref > gpurun_out/${tag}_many_radios_ref.json 2>/dev/null; cut -c1-300 gpurun_out/${tag}_many_radios_ref.json
