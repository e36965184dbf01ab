set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | cut -c1-330
