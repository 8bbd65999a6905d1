for v in 0 3 5 7; do
  NM_NVCC_EXTRA="-DNM_SIM_COOP_MAX_ROUNDS=$v" python -c "from navigation_model_b200 import _build; _build.build_native(force=True)" 
  echo "== coop max rounds $v"
  python tools/bench_sim.py --mice 1048576 --steps 1000 --reps 2 --cpu-mice 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kernel_ms_best'], d['value'])"
done
