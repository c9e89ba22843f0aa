#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU).  Each program is first run WITHOUT ncu and must exit 0.
#   1. launch list of one eager training step (gpu__time_duration per launch, cold cache, serialised)
#   2. ncu --set full of every hot kernel at BASELINE sizes (one launch each)
#   3. ncu --set full of the long-sequence attention kernel at L = 20 020
# Reduce with tools/summarise_ncu.py / tools/launch_summary.py into profiles/.
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
python tools/prof_step.py tf32 > gpurun_out/${TAG}_step_plain.log 2>&1 || { echo "prof_step failed"; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_tf32.csv python tools/prof_step.py tf32 > gpurun_out/${TAG}_step_ncu.log 2>&1
python tools/prof_kernels_r2.py > gpurun_out/${TAG}_kernels_plain.log 2>&1 || { echo "prof_kernels_r2 failed"; cat gpurun_out/${TAG}_kernels_plain.log | tail -5; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_umma|gemm_ln|attention_fwd_umma|attention_bwd|heads_|batch_reduce|res_ln_bwd|cost_matrix|cost_targets|lsap_reg|matched_loss" \
    -o gpurun_out/${TAG}_kernels -f python tools/prof_kernels_r2.py > gpurun_out/${TAG}_kernels_ncu.log 2>&1
python tools/summarise_ncu.py gpurun_out/${TAG}_kernels.ncu-rep > gpurun_out/${TAG}_ncu_full_kernels.csv
rm -f gpurun_out/${TAG}_kernels.ncu-rep      # (60 MB with sources: gpurun copies at most 64 MiB back; the csv summary is what profiles/ keeps)
python tools/prof_attention.py > gpurun_out/${TAG}_attn_plain.log 2>&1 || { echo "prof_attention failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:attention_fwd_umma_ms_kernel --launch-skip 2 --launch-count 1 \
    -o gpurun_out/${TAG}_attn_ms -f python tools/prof_attention.py > gpurun_out/${TAG}_attn_ncu.log 2>&1
python tools/summarise_ncu.py gpurun_out/${TAG}_attn_ms.ncu-rep > gpurun_out/${TAG}_ncu_full_attention_ms.csv
python tools/prof_attention.py f16 > gpurun_out/${TAG}_attn_f16_plain.log 2>&1 || { echo "prof_attention f16 failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:attention_fwd_umma_ms_f16 --launch-skip 2 --launch-count 1 \
    -o gpurun_out/${TAG}_attn_ms_f16 -f python tools/prof_attention.py f16 > gpurun_out/${TAG}_attn_f16_ncu.log 2>&1
python tools/summarise_ncu.py gpurun_out/${TAG}_attn_ms_f16.ncu-rep > gpurun_out/${TAG}_ncu_full_attention_ms_f16.csv
ls -la gpurun_out/${TAG}_*
