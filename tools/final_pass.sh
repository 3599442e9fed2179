#!/bin/bash
# developer tool (GPU box): the measurements that go into profiles/ - every step under `timeout`
mkdir -p gpurun_out
T=${1:-s2}
tools/gpu_check.sh || exit 1
timeout 400 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"; cat gpurun_out/bench_$T.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "ref rc=$?"
timeout 300 python tools/size_sweep.py > gpurun_out/size_sweep_$T.log 2>&1; echo "sweep rc=$?"
timeout 200 python tools/aux_bench.py > gpurun_out/aux_$T.log 2>&1; echo "aux rc=$?"
timeout 200 python tools/blot_bench.py > gpurun_out/blot_$T.log 2>&1; echo "blot rc=$?"
timeout 200 python tools/pipeline_bench.py > gpurun_out/pipe_$T.log 2>&1; echo "pipe rc=$?"
timeout 200 python tools/ingest_bench.py > gpurun_out/ingest_$T.log 2>&1; echo "ingest rc=$?"
timeout 200 python tools/single_call_bench.py > gpurun_out/single_$T.log 2>&1; echo "single call rc=$?"
timeout 200 python tools/dropin_bench.py > gpurun_out/dropin_$T.log 2>&1; echo "drop-in rc=$?"
timeout 200 python tools/files_to_fit_bench.py > gpurun_out/files_$T.log 2>&1; echo "files to fit rc=$?"
timeout 200 python tools/stress_large.py > gpurun_out/stress_$T.log 2>&1; echo "stress rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cc_fit_kernel|FillFunctor" -c 40 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$T.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cc_fit_kernel -s 3 -c 1 -o gpurun_out/prof_${T}_64 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu64_$T.log 2>&1; echo "ncu64 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cc_fit_kernel -s 2 -c 1 -o gpurun_out/prof_${T}_128 python tools/prof_stack.py 128 3000 > gpurun_out/ncu128_$T.log 2>&1; echo "ncu128 rc=$?"
