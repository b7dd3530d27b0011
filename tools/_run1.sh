python -m pytest tests -x -q -m gpu -k "multiclass or final or post or Post or config or graph" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline | tail -1 | cut -c1-160
