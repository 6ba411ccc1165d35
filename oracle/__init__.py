"""oracle/ — TEST INFRASTRUCTURE, never imported by the product path.

CPU restatements of the DiffPhyCon hot path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package.
"""
