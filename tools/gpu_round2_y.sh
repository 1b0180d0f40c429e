#!/bin/bash
# HEAD validation: whole GPU parity suite (with the slowest tests listed) + smoke
s=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
echo "pytest seconds: $(( $(date +%s) - s ))"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
