# developer A/B (GPU box): step time in us of the default build against environment switches, alternating
run() { python bench.py --quick --steps 300 --warmup 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1000,1))"; }
for rep in 1 2 3; do
echo -n "default: "; run
for v in "$@"; do echo -n "$v: "; env $v bash -c "$(declare -f run); run"; done
done
