python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
echo "=== baseline (no L2 prefetch)"; python scripts/microbench.py engine 2>&1 | grep -E "prefill step|decode step|replay|attention|all "
echo "=== B200_L2_PREFETCH=16"; B200_L2_PREFETCH=16 python scripts/microbench.py engine 2>&1 | grep -E "prefill step|decode step|replay|attention|all "
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention or silu" -p no:cacheprovider 2>&1 | tail -2
