"""Copies the summaries of gpurun_out/ (written by capture_evidence.sh [tag]) into profiles/ and
exports the ncu reports as raw-page CSV (run where ncu is installed; no GPU needed):
    python profiles/tools/export_evidence.py [tag]"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, 'gpurun_out')
DST = os.path.join(ROOT, 'profiles')
COPIES = ['bench_final.json', 'bench_reference.json', 'launches.csv', 'k3_images_per_sm.txt', 'k1_cases.txt',
          'pipeline_probe.txt']


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
    for name in COPIES:
        p = os.path.join(SRC, '%s_%s' % (tag, name))
        if os.path.exists(p) and os.path.getsize(p):
            shutil.copyfile(p, os.path.join(DST, '%s_%s' % (tag, name)))
    for part in ('k1', 'chain'):
        rep = os.path.join(SRC, '%s_prof_%s.ncu-rep' % (tag, part))
        if os.path.exists(rep):
            with open(os.path.join(DST, '%s_ncu_%s_raw.csv' % (tag, part)), 'w') as f:
                subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=f, check=False)


if __name__ == '__main__':
    main()
