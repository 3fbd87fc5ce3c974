for i in 1 2; do
for v in 0 1; do
  if [ $v = 1 ]; then export MVLPT_NO_FUSED_LN=1; else unset MVLPT_NO_FUSED_LN; fi
  python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-eager-baseline --no-roofline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('coop nofuse=$v', round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
  python bench.py --config 3 --steps 15 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-roofline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('vpt  nofuse=$v', round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
done; done
