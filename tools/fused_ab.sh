#!/bin/bash
# A/B of the fused factorisation + inverse launch (csrc/fused2.cuh): every rank of an N-way shard emulated on ONE GPU
# (tools/shard_probe.py), fused off / on; the 1-GPU bench with the fused launch forced; parity tests with it forced.
for W in 8 4; do
  for F in 0 1; do echo "== world $W fused $F"; DSMGP_FUSED_EVAL=$F python tools/shard_probe.py $W cfg3 2>&1 | tail -3; done
done
for F in 0 1; do DSMGP_FUSED_EVAL=$F python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('1 GPU fused=$F ms', round(d['ms_per_step'],3), d['phases_ms_per_step'], d['roofline']['kernel'], round(d['roofline']['achieved'],2))"; done
DSMGP_FUSED_EVAL=1 python -m pytest tests -m gpu -q -p no:cacheprovider -k "dsmgp or golden or full_size_cfg3_against or finetune or mpmath or single_gp or medium or streaming or multirank" 2>&1 | tail -4
