#!/bin/bash
# round-2 evidence pass (tag $1): smoke, GPU tests, default bench (both arms), ncu launch list of one eager step, in-situ
# ablation, per-call timing, ncu --set full of the dominant conv, the GroupNorm-apply variants and the render kernel
V=${1:-v3}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_$V.log 2>&1; echo "smoke $?"
timeout 900 python -m pytest -q --timeout=200 -p no:cacheprovider tests -m gpu > gpurun_out/r2_tests_$V.log 2>&1; echo "tests $?"
tail -2 gpurun_out/r2_tests_$V.log
timeout 1500 python bench.py > gpurun_out/r2_bench_n1_$V.json 2> gpurun_out/r2_bench_n1_$V.err; echo "bench $?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_ref_n1_$V.json 2> gpurun_out/r2_bench_ref_n1_$V.err; echo "bench-ref $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_step_$V.csv python tools/profile_step.py > gpurun_out/r2_prof_step.log 2>&1; echo "launchlist $?"
python tools/summarize_launches.py gpurun_out/r2_launches_step_$V.csv > gpurun_out/r2_launches_step_$V.md 2>&1
timeout 300 python tools/ablate_step.py > gpurun_out/r2_ablate_step_$V.log 2>&1
timeout 300 python tools/time_calls.py > gpurun_out/r2_time_calls_$V.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 3 -c 1 -f -o gpurun_out/r2_conv_full_$V python tools/profile_conv.py > gpurun_out/r2_ncu_conv.log 2>&1; echo "ncu conv $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gn_apply -s 3 -c 1 -f -o gpurun_out/r2_gn_apply_f32in_$V python tools/profile_gn.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gn_apply -s 26 -c 1 -f -o gpurun_out/r2_gn_apply_f16in_$V python tools/profile_gn.py > /dev/null 2>&1; echo "ncu gn $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render_tc5 -c 1 -s 1 -f -o gpurun_out/r2_render_tc5_$V python tools/render_ncu.py fp16 > gpurun_out/r2_ncu_render.log 2>&1; echo "ncu render $?"
timeout 300 python tools/render_probe.py fp16 > gpurun_out/r2_render_probe_$V.log 2>&1
timeout 100 python tools/profile_gn.py > gpurun_out/r2_gn_time_$V.log 2>&1
cat gpurun_out/r2_launches_step_$V.md
cat gpurun_out/r2_ablate_step_$V.log
cut -c1-600 gpurun_out/r2_bench_n1_$V.json
cut -c1-400 gpurun_out/r2_bench_ref_n1_$V.json
