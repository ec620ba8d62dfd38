mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
