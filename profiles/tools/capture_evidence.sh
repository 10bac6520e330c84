#!/bin/bash
# Evidence capture of one round on one B200 (run under gpurun from the repo root):
#   bash profiles/tools/capture_evidence.sh
# writes into gpurun_out/; profiles/tools/export_evidence.py copies the summaries into profiles/.
set -u
out=gpurun_out
mkdir -p $out
python bench.py --steps 20 --warmup 5 > $out/bench_final.json 2> $out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_final.csv \
    python bench.py --steps 3 --warmup 3 > $out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nms_candidates_kernel" -s 3 -c 1 \
    -f -o $out/prof_final_k1 python bench.py --steps 3 --warmup 3 --hot-only > $out/ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"amax_scan_kernel|fused_block_kernel|block_list_kernel" \
    -s 9 -c 3 -f -o $out/prof_final_k1f python profiles/tools/k1f_bench.py > $out/ncu_k1f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"group_kernel|limb_score_kernel|select_topk" \
    -s 12 -c 4 -f -o $out/prof_final_k23 python profiles/tools/k1f_bench.py > $out/ncu_k23.log 2>&1
python profiles/tools/dev_pipeline_profile.py > $out/dev_pipeline_profile.txt 2>&1
python profiles/tools/e2e_host_profile.py > $out/e2e_host_profile.txt 2>&1
python profiles/tools/k3_occupancy_sweep.py > $out/k3_occupancy_sweep.txt 2>&1
python -c "
import json
d = json.load(open('$out/bench_final.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'features_dev', d['features_dev']['value'], 'frac', d['roofline']['frac'])
"
