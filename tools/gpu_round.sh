python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python tools/run_configs.py --out gpurun_out/configs_r1_final.jsonl 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(d['config'][:44], '| enc', round(d['encode_gbs'],1), 'GB/s | dec', round(d.get('decode_gbs',0),1), 'GB/s | exact', d['roundtrip_exact'], '| analyze', d['stage_ms']['analyze_ms'], 'stereo', d['stage_ms']['stereo_ms'], 'parse', d['stage_ms'].get('parse_ms'))
    else: print(ln.strip())
"
