#!/bin/bash
# Round profiling recipe (run under gpurun, one GPU): launch list of one eager training step + ncu --set full of the
# hot kernels.  Each program is first run WITHOUT ncu and must exit 0.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1b}
python tools/prof_step.py tf32 > gpurun_out/${TAG}_step_plain.log 2>&1 || { echo "prof_step failed"; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_tf32.csv python tools/prof_step.py tf32 > gpurun_out/${TAG}_step_ncu.log 2>&1
python tools/prof_kernels.py tf32 > gpurun_out/${TAG}_kernels_plain.log 2>&1 || { echo "prof_kernels failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"gemm_umma|attention_fwd_umma|attention_bwd|cost_matrix|cost_targets|lsap_kernel|sgd_" \
    --launch-skip 20 --launch-count 22 -o gpurun_out/${TAG}_kernels -f python tools/prof_kernels.py tf32 > gpurun_out/${TAG}_kernels_ncu.log 2>&1
python tools/prof_attention.py > gpurun_out/${TAG}_attn_plain.log 2>&1 || { echo "prof_attention failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:attention_fwd_umma_ms --launch-skip 2 --launch-count 1 \
    -o gpurun_out/${TAG}_attn_ms -f python tools/prof_attention.py > gpurun_out/${TAG}_attn_ncu.log 2>&1
ls -la gpurun_out/${TAG}_*
