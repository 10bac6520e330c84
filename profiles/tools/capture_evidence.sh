#!/bin/bash
# Evidence capture of one round on one B200 (run under gpurun from the repo root):
#   bash profiles/tools/capture_evidence.sh [tag]        (tag defaults to r2)
# writes gpurun_out/<tag>_*; profiles/tools/export_evidence.py <tag> copies the summaries into
# profiles/ and exports the ncu reports as CSV.
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_final.json 2> $out/${tag}_bench_final.err
python bench.py --impl reference --steps 3 --warmup 3 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --lean > $out/${tag}_ncu_list.log 2>&1
# K1 (the SURVEY 8d roofline kernel) on the materialised maps
ncu --set full --clock-control none --import-source on -k regex:"nms_candidates_kernel" -s 3 -c 1 \
    -f -o $out/${tag}_prof_k1 python bench.py --steps 3 --warmup 3 --hot-only > $out/${tag}_ncu_k1.log 2>&1
# the product chain: K1f (scan, list, blocks), select, K2, K3
ncu --set full --clock-control none --import-source on \
    -k regex:"amax_scan_kernel|fused_block_kernel|block_list_kernel|select_topk_kernel|limb_score_kernel|group_warp_kernel" \
    -s 12 -c 6 -f -o $out/${tag}_prof_chain python profiles/tools/chain_once.py 64 > $out/${tag}_ncu_chain.log 2>&1
python profiles/tools/k3_time.py > $out/${tag}_k3_images_per_sm.txt 2>&1
python profiles/tools/k1_cases.py > $out/${tag}_k1_cases.txt 2>&1
python profiles/tools/pipeline_probe.py 640 8,64 1,4,8,16 > $out/${tag}_pipeline_probe.txt 2>&1
python -c "
import json
d = json.load(open('$out/${tag}_bench_final.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'fused frac', d['roofline_fused']['frac'])
"
