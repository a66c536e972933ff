# developer loop for the fp32-accurate tensor-core prefill: parity cases, then the shapes profiles/r02_prefill.md quotes
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc3 or 3xtf32 or prefill" 2>&1 | tail -3
for T in 1024 2048 4096; do timeout 120 python tools/prefill_bench.py --shape 124m --B 16 --T $T --path 4 --check 2>&1 | tail -1 | cut -c150-260; done
timeout 120 python tools/prefill_bench.py --shape long --B 2 --T 2048 --before 30000 --path 4 --check 2>&1 | tail -1| cut -c150-260
timeout 120 python tools/prefill_bench.py --shape xl --B 16 --T 2048 --path 4 --check 2>&1 | tail -1| cut -c150-260
PA_PREFILL_TC3_TIMELINE=1 timeout 120 python tools/prefill_bench.py --shape 124m --B 16 --T 2048 --path 4 --iters 1 --warmup 2 2>&1 | grep -A2 "tc3 timeline" | tail -3 | cut -c1-900
