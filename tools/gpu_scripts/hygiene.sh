#!/bin/bash
# end-of-round evidence: smoke, GPU tests, default bench, ncu launch list of one step, ncu --set full of the dominant launch
V=${1:-v10}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke $?" > gpurun_out/summary.txt
timeout 600 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err; echo "bench $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$V.json 2> gpurun_out/bench_ref_$V.err; echo "bench-ref $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_$V.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "ncu $?" >> gpurun_out/summary.txt
python tools/summarize_launches.py gpurun_out/launches_step_$V.csv > gpurun_out/step_breakdown_$V.md 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 3 -c 1 -o gpurun_out/conv_full_$V -f python tools/profile_conv.py > gpurun_out/ncu_full.log 2>&1
echo "ncufull $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -2 gpurun_out/smoke.log
tail -3 gpurun_out/t_all.log
cat gpurun_out/step_breakdown_$V.md
cut -c1-2800 gpurun_out/bench_$V.json
cut -c1-800 gpurun_out/bench_ref_$V.json
tail -3 gpurun_out/ncu_full.log
