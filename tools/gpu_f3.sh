mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_preprocess.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout=200 > gpurun_out/preprocess_z.log 2>&1
echo "preprocess rc=$?"; tail -15 gpurun_out/preprocess_z.log | cut -c1-300
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout=200 -k "topk" > gpurun_out/topk_z.log 2>&1
echo "topk rc=$?"; tail -8 gpurun_out/topk_z.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_z.json')); print(d['value'], d['e2e'], d.get('e2e_u8_input'))"; tail -5 gpurun_out/bench_z.err
