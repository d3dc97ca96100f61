for cpw in 32 16 8; do
  LQB_CPW=$cpw timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cpw $cpw', round(d['value']), d['ms_per_step'], d['roofline']['frac'])"
done
