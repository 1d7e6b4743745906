TAG=${1:-r02g}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
(time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_reference_n1.json 2> gpurun_out/${TAG}_reference_n1.err); cut -c1-300 gpurun_out/${TAG}_reference_n1.json
(time timeout 600 python bench.py --impl reference --path raster --steps 20 --warmup 5 > gpurun_out/${TAG}_reference_raster_n1.json 2>> gpurun_out/${TAG}_reference_n1.err); cut -c1-300 gpurun_out/${TAG}_reference_raster_n1.json
(time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err); tail -3 gpurun_out/${TAG}_bench_n1.err
python tools/bench_summary.py gpurun_out/${TAG}_bench_n1.json
python -c "import __graft_entry__ as g; g.smoke()"
