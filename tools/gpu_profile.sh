#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, full captures of the dominant kernels (each only after
# the same command has exited 0 without ncu).
#   gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh'      then: python tools/ncu_summary.py gpurun_out/<x>.ncu-rep profiles/<y>.txt <traffic key>
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pcm16_decode_batch -s 3 -c 1 -f -o gpurun_out/r02_prof_decode $CMD > gpurun_out/ncu_decode.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:voice_render_mix_tma -s 3 -c 1 -f -o gpurun_out/r02_prof_render_c2 $CMD > gpurun_out/ncu_render.log 2>&1
