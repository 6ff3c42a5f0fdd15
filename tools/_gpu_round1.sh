set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_exact.py tests/test_trapezoid.py tests/test_gpu_nlp.py -m gpu -x -q 2>&1 | tail -n 3
cd tools && timeout 600 python ab_small.py > ../gpurun_out/ab_small4.jsonl 2> ../gpurun_out/ab_small4.err; cd ..; cat gpurun_out/ab_small4.jsonl; tail -n 5 gpurun_out/ab_small4.err
