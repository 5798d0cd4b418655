#!/bin/bash
# compute-sanitizer memcheck over one small sampler run (smoke: B=2, N=24, 3 steps) and a few multi-tile GPU tests
mkdir -p gpurun_out
O=gpurun_out/r02_compute_sanitizer_memcheck.txt
echo "# compute-sanitizer --tool memcheck on B200 (round-2 final build)" > $O
echo "## python -c 'import __graft_entry__ as g; g.smoke()'" >> $O
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -8 >> $O
echo "## pytest -m gpu -k 'multi_tile or embed_ipa_edge or tmem or seq_tfmr'" >> $O
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k "multi_tile or embed_ipa_edge or tmem or seq_tfmr" 2>&1 | grep -v "^$" | tail -8 >> $O
cat $O
