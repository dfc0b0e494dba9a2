set -x
TAG=${TAG:-r2r}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.txt
tail -4 gpurun_out/${TAG}_pytest.txt
timeout 800 ncu --set full --import-source on --clock-control none -k regex:solve_track2_kernel -c 1 -o gpurun_out/prof_${TAG} python tools/prof_run.py --calls 1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_run.py --calls 2 > /dev/null 2>&1
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json
