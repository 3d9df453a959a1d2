#!/bin/bash
# A/B of an environment knob on the engine microbenchmark (usage: KNOB="B200_FUSE_ROPE=0" bash scripts/ab.sh)
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
F='prefill step|decode step|replay|attention|all |rope'
echo "=== default"; python scripts/microbench.py engine 2>&1 | grep -E "$F"
if [ -n "$KNOB" ]; then echo "=== $KNOB"; env $KNOB python scripts/microbench.py engine 2>&1 | grep -E "$F"; fi
