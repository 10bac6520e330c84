"""Copies the summaries of gpurun_out/ (written by capture_evidence.sh) into profiles/ and exports
the ncu reports as raw-page CSV (run where ncu is installed; no GPU needed)."""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, 'gpurun_out')
DST = os.path.join(ROOT, 'profiles')
COPIES = {
    'bench_final.json': 'r1_s2_bench_final.json',
    'bench_reference.json': 'r1_s2_bench_reference_arm.json',
    'launches_final.csv': 'r1_s2_launches.csv',
    'dev_pipeline_profile.txt': 'r1_s2_dev_pipeline_profile.txt',
    'e2e_host_profile.txt': 'r1_e2e_chunk_sweep.txt',
    'k3_occupancy_sweep.txt': 'r1_k3_occupancy_sweep.txt',
}


def main():
    for src, dst in COPIES.items():
        p = os.path.join(SRC, src)
        if os.path.exists(p):
            shutil.copyfile(p, os.path.join(DST, dst))
    for tag in ('k1', 'k1f', 'k23'):
        rep = os.path.join(SRC, 'prof_final_%s.ncu-rep' % tag)
        if os.path.exists(rep):
            with open(os.path.join(DST, 'r1_s2_ncu_%s_raw.csv' % tag), 'w') as f:
                subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=f, check=False)


if __name__ == '__main__':
    main()
