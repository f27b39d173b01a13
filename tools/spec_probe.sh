for R in 512 256 128; do
  for SP in 1 0; do
    PTFNN_SPEC=$SP timeout 200 python bench.py --no-cpu-baseline --replicas-per-gpu $R --steps 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('R',$R,'PTFNN_SPEC',$SP, round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))"
  done
done
PTFNN_SPEC=1 timeout 100 python bench.py --no-cpu-baseline --workload sunspot 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sunspot spec off', round(d['value']))"
timeout 100 python bench.py --no-cpu-baseline --workload sunspot 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sunspot auto', round(d['value']), round(d['e2e']['value']))"
