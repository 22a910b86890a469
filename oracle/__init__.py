"""TEST INFRASTRUCTURE ONLY — CPU restatement of the RealPDEBench FNO hot path.

Nothing in the product package (`realpdebench_b200/`) may import this package.
Allowed importers: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs, where it is the checker or the timed
CPU baseline and never the thing shipped.
"""
