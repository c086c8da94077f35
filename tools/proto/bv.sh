for v in 161 181 141 142 132 281 261; do
  echo -n "variant $v: "; RBQ_MC_VARIANT=$v python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
echo -n "no-table 161: "; RBQ_NO_TABLE_CLASS=1 RBQ_MC_VARIANT=161 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo -n "no-table 181: "; RBQ_NO_TABLE_CLASS=1 RBQ_MC_VARIANT=181 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
