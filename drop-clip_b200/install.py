"""Makes the reference's own drivers pick up this implementation without editing them.

    import dropclip_b200.install; dropclip_b200.install.install()

aliases `utils.feature_fusion`, `utils.projections` and `models.similarity` in `sys.modules` to
the modules of this package, so `tools/preprocess_data.py`, `scripts/run_eval.py`,
`tools/validate_upper_bound.py` and `engine/distil.py` (which import those names) run unchanged
(SURVEY.md §8b layer 1). With `voxelizer=True` a stand-in `MinkowskiEngine.utils` exposing
sparse_quantize / sparse_collate is registered as well, when ME itself is not importable; with
`metrics=True` the reference's `utils.misc.trainMetricPC` / `intersectionAndUnionGPU` are replaced by the
CUDA versions of `dropclip_b200.metrics`.
"""
from __future__ import annotations

import importlib
import sys
import types

ALIASES = {
    "utils.feature_fusion": "dropclip_b200.feature_fusion",
    "utils.projections": "dropclip_b200.projections",
    "models.similarity": "dropclip_b200.similarity",
}


def install(voxelizer: bool = False, metrics: bool = False) -> None:
    for ref_name, ours in ALIASES.items():
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        parent, _, child = ref_name.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], child, mod)
    if voxelizer and "MinkowskiEngine" not in sys.modules:
        from . import voxelize
        me = types.ModuleType("MinkowskiEngine")
        me.utils = types.ModuleType("MinkowskiEngine.utils")
        me.utils.sparse_quantize = voxelize.sparse_quantize
        me.utils.sparse_collate = voxelize.sparse_collate
        sys.modules["MinkowskiEngine"] = me
        sys.modules["MinkowskiEngine.utils"] = me.utils


    if metrics:
        # utils/misc.py holds unrelated helpers too (meters, logger, seeds), so only the two metric functions
        # are replaced, inside the reference's own module, before the drivers do `from utils.misc import ...`
        from . import metrics as ours
        misc = importlib.import_module("utils.misc")  # the reference tree must be on sys.path
        misc.trainMetricPC = ours.trainMetricPC
        misc.intersectionAndUnionGPU = ours.intersectionAndUnionGPU


def uninstall() -> None:
    for ref_name in ALIASES:
        sys.modules.pop(ref_name, None)
