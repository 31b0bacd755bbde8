"""ORACLE package: CPU restatements of the DROP-CLIP fusion/grounding hot path.

Test infrastructure only. Importable from `tests/`, `__graft_entry__.smoke()` and the
CPU-baseline / `--impl reference` legs of `bench.py`; the product package never imports it.
"""
