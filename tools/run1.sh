set -x
python -c "import torch;print(torch.cuda.device_count())"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_pytest.txt
