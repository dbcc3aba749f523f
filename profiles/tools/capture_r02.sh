#!/bin/bash
# Round-2 evidence capture (run on the GPU box through gpurun; results land in gpurun_out/, summaries are then made
# here with profiles/tools/ncu_summary.py / decoder_layer_table.py and committed under profiles/r02_*).
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum"
# 1. launch list of the bench command (eager enqueue, so every kernel is a separate launch); one step is cut out of it
#    afterwards (profiles/tools/step_from_launches.py)
timeout 600 ncu --metrics $M --clock-control none --kernel-name-base demangled -c 3000 --csv \
    --log-file gpurun_out/r02_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph \
    > gpurun_out/r02_launches_step.out 2>&1
# 2. --set full of the dominant kernels at the bench shapes (depth-2 model, same 8 x 1025 rows)
for spec in "attn:regex:attn_tc_kernel:4" "ffin:regex:EpiSwiglu:2" "qkv:regex:EpiQkvRope:2" "ln:regex:layernorm_kernel:4"; do
  name=${spec%%:*}; rest=${spec#*:}; kern=${rest%:*}; skip=${rest##*:}
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k $kern -s $skip -c 1 \
      -f -o gpurun_out/r02_$name python tests/prof_step.py dit > gpurun_out/r02_$name.log 2>&1
done
# 3. decoder: launch list with the metrics of the per-layer table, and --set full of the fused 128-channel unit
timeout 600 ncu --metrics $M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --kernel-name-base demangled -k regex:satb -s 139 -c 29 --csv --log-file gpurun_out/r02_launches_decoder.csv \
    python tests/prof_step.py oobleck 1024 > gpurun_out/r02_launches_decoder.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:resunit_tcgen05_2cta -s 5 -c 1 \
    -f -o gpurun_out/r02_resunit python tests/prof_step.py oobleck 1024 > gpurun_out/r02_resunit.log 2>&1
ls -la gpurun_out/r02_* | head -30
