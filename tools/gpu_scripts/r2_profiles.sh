set -x
python tools/profile_attention.py > gpurun_out/r2_attn_time.log 2>&1
python bench.py --steps 20 --warmup 3 --no-layered --no-cpu-baseline --no-b64 > gpurun_out/r2_b4_bench.json 2> gpurun_out/r2_b4_bench.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_step_v2.csv python tools/profile_step.py > gpurun_out/r2_prof_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_step_v2.csv > gpurun_out/r2_launches_step_v2.md 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_attention_tc5 -s 3 -c 1 -o gpurun_out/r2_attention_tc5_full python tools/profile_attention.py > gpurun_out/r2_ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 3 -c 1 -o gpurun_out/r2_conv_full python tools/profile_conv.py > gpurun_out/r2_ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_tc5 -c 1 -s 1 -o gpurun_out/r2_render_tc5_v4 python tools/render_ncu.py fp16 > gpurun_out/r2_ncu_render.log 2>&1
python tools/ablate_step.py > gpurun_out/r2_ablate_v2.log 2>&1
cat gpurun_out/r2_attn_time.log
