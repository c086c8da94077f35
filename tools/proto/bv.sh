for v in 161 171 181; do for r in 1 2; do
  echo -n "variant $v: "; RBQ_MC_VARIANT=$v python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
done; done
