for pair in 0 1; do for dbg in 1 3; do
PB200_TC_PAIR=$pair PB200_TC_DEBUG=$dbg timeout 100 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('pair $pair dbg $dbg kernel_ms', d['roofline']['kernel_ms'])"
done; done
