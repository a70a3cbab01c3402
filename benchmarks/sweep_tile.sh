run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],1), round(d['roofline']['frac'],3), round(d['roofline']['kernel_share_of_step'],3))"; }
run BR_TILE_TPB=2
run BR_TILE_TPB=8
run BR_TILE_TPB=16
run BR_TILE_G=2
run BR_ROW_FRAC=0.1
run BR_ROW_FRAC=0.35
run BR_HOT_FRAC=0.125
run BR_HOT_FRAC=0.5
run BR_HOT_FRAC=1.0
