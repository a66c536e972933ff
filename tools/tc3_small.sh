# developer loop: tiled fp32 SIMT prefill (path 1) against the 3xTF32 tensor-core kernel (path 4) at small steps
for shp in "64 4 1024" "64 16 1024" "16 64 1024" "8 128 4000" "64 2 1024" "256 2 512" "1 8 30000" "4 32 8000"; do set -- $shp
  for path in 1 4; do
    timeout 60 python tools/prefill_bench.py --shape 124m --B $1 --T $2 --before $3 --path $path --iters 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B',d['B'],'T',d['T'],'before',d['before'],'path',$path, 'ms %.4f'%d['ms'], 'TF %.1f'%d['tflops'])"
  done
done
for shp in "8 4 8000" "2 64 30000"; do set -- $shp
  for path in 1 4; do
    timeout 60 python tools/prefill_bench.py --shape long --B $1 --T $2 --before $3 --path $path --iters 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('long B',d['B'],'T',d['T'],'before',d['before'],'path',$path, 'ms %.4f'%d['ms'], 'TF %.1f'%d['tflops'])"
  done
done
