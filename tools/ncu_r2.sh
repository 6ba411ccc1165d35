#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU).  Numbers printed by runs under ncu are never bench values.
# Every .ncu-rep is summarised ON THE BOX (tools/ncu_summary.py -> text) and deleted: gpurun merges at most 64 MiB back.
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 )) s"; }
B="python bench.py --scaling weak --no-cuda-graph --no-cpu-baseline --no-e2e --no-rollout --no-roofline"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
summarise() {   # <name>: gpurun_out/<name>.ncu-rep -> gpurun_out/<name>.txt (+ raw csv when small), rep removed
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1.txt 2>&1
  rm -f gpurun_out/$1.ncu-rep
}
# (0) launch list of the bench command itself (default arm, batch 64, eager and graph), per-launch time only
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-e2e --no-rollout > gpurun_out/r2_ncu_bench.log 2>&1
lap bench-launch-list
# (1) launch lists with DRAM bytes of one denoising step at batch 16: whole batch vs micro-batches of 2 (the L2-resident schedule)
ncu --metrics $M --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_b16.csv $B --batch 16 --steps 1 --warmup 1 > gpurun_out/r2_ncu_b16.log 2>&1
lap b16
ncu --metrics $M --clock-control none -c 12000 --csv --log-file gpurun_out/r2_launches_b16_mb2.csv $B --batch 16 --steps 1 --warmup 1 --micro-batch 2 > gpurun_out/r2_ncu_b16_mb2.log 2>&1
lap b16-mb2
python tools/launch_list_summary.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench.md
python tools/launch_list_summary.py gpurun_out/r2_launches_b16.csv > gpurun_out/r2_launches_b16.md
python tools/launch_list_summary.py gpurun_out/r2_launches_b16_mb2.csv > gpurun_out/r2_launches_b16_mb2.md
gzip -f gpurun_out/r2_launches_b16.csv gpurun_out/r2_launches_b16_mb2.csv gpurun_out/r2_launches_bench.csv
# (2) --set full: one capture per kernel family of the smoke step (batch 8)
ncu --set full --clock-control none -k regex:"groupnorm_silu|guided_step|final_proj|layernorm_channels|pack_input|gn_fold" -c 8 -o gpurun_out/r2_full_stream -f $B --batch 8 --steps 1 --warmup 1 > gpurun_out/r2_ncu_full_stream.log 2>&1
summarise r2_full_stream; lap full-stream
ncu --set full --clock-control none -k regex:"conv3d_tc_kernel" -s 2 -c 14 -o gpurun_out/r2_full_conv -f $B --batch 8 --steps 1 --warmup 0 > gpurun_out/r2_ncu_full_conv.log 2>&1
summarise r2_full_conv; lap full-conv
ncu --set full --clock-control none -k regex:"temporal_attention_mma|linattn_context_kernel|linattn_apply_kernel|spatial_attention_mma|temporal_block|linattn_.*_tc|stem_conv" -c 10 -o gpurun_out/r2_full_attn -f $B --batch 8 --steps 1 --warmup 0 > gpurun_out/r2_ncu_full_attn.log 2>&1
summarise r2_full_attn; lap full-attn
# (3) the dominant kernel alone at batch 16 (roofline.traffic of bench.py): the .ncu-rep is small and kept
ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 2 -c 1 -o gpurun_out/r2_full_dominant -f python tools/run_kernels_once.py conv 16 > gpurun_out/r2_ncu_dominant.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_full_dominant.ncu-rep > gpurun_out/r2_full_dominant.txt 2>&1
ncu -i gpurun_out/r2_full_dominant.ncu-rep --page raw --csv > gpurun_out/r2_full_dominant_raw.csv 2>/dev/null
lap dominant
# (4) jellyfish surrogate-net backward kernels (2 trajectories x 20 frames at 128x128)
ncu --set full --clock-control none -k regex:"linattn.*bwd|gn_silu_bwd|attention.*bwd|layernorm.*bwd|mean_head|time_mlp_bwd|sumpool" -c 12 -o gpurun_out/r2_full_jelly -f python tools/time_jelly_nets.py 2 128 > gpurun_out/r2_ncu_full_jelly.log 2>&1
summarise r2_full_jelly; lap full-jelly
# (5) rollout CG kernel, 4 trajectories x 16 frames
ncu --set full --clock-control none -k regex:smoke_rollout -c 1 -o gpurun_out/r2_full_rollout -f python tools/time_rollout.py 4 16 > gpurun_out/r2_ncu_full_rollout.log 2>&1
summarise r2_full_rollout; lap full-rollout
ls -la gpurun_out/; du -sh gpurun_out
