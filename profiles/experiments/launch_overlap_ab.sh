set -x
DDZ_LAUNCH_OVERLAP=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
for ov in 0 1; do
DDZ_LAUNCH_OVERLAP=$ov python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ov=$ov 20', d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_single_chain'), d['e2e']['ms_per_step'], d['verify'])"
DDZ_LAUNCH_OVERLAP=$ov python bench.py --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ov=$ov 1000', d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_single_chain'), d['e2e']['ms_per_step'], d['verify'])"
done; done
