"""Writes tests/golden/class_sim.npz: outputs of the UNMODIFIED reference closure `_get_similarity`
(engine/distil.py:244-246; its twin tools/validate_upper_bound.py:59-61) followed by the reference's own arg max line
(`pred = torch.max(sims, 1)[1]`, :290 / :102) on seeded inputs. The closure is nested inside validate_segmentation, so
it is lifted out of the reference's source text with `ast` and compiled as it stands - no line of it is restated here.
Run in the build container only (`python tests/make_golden_class_sim.py`)."""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "class_sim.npz")
CASES = (("m10k_k44", 21, 10_000, 44, 768), ("m3k_k300", 22, 3_000, 300, 768), ("m1_k44", 23, 1, 44, 768), ("m500_k5_c512", 24, 500, 5, 512))


def reference_closure(rel_path):
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, rel_path)).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "_get_similarity":
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"torch": torch}
            exec(compile(mod, rel_path, "exec"), ns)
            return ns["_get_similarity"]
    raise RuntimeError(f"_get_similarity not found in {rel_path}")


def class_inputs(seed, m, k, c):
    """(features (M,C) fp32 - network outputs, not normalised -, class table (K,C) fp32 - not normalised)."""
    rng = np.random.default_rng(seed)
    table = rng.standard_normal((k, c)).astype(np.float32) * rng.uniform(0.5, 3.0, size=(k, 1)).astype(np.float32)
    owner = rng.integers(0, k, size=m)
    x = (table[owner] / np.linalg.norm(table[owner], axis=1, keepdims=True) * 2.0 + 0.35 * rng.standard_normal((m, c))).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(table)


def main():
    f_distil = reference_closure("engine/distil.py")
    f_upper = reference_closure("tools/validate_upper_bound.py")
    g = {}
    for name, seed, m, k, c in CASES:
        x, table = class_inputs(seed, m, k, c)
        query = table.clone()
        sims = f_distil(x, query.float())   # engine/distil.py:289 passes query.float(): the same tensor, normalised in place
        pred = torch.max(sims, 1)[1]        # :290
        sims2 = f_upper(x, table.clone().float())
        assert torch.equal(sims, sims2)
        g[f"{name}_pred"] = pred.numpy().astype(np.int16)
        g[f"{name}_sims_rows"] = sims[:: max(1, m // 512)].numpy()
        g[f"{name}_table_after"] = query.numpy()  # the in-place normalisation the caller observes
        top2 = torch.topk(sims, min(2, k), dim=1).values
        g[f"{name}_margin"] = (top2[:, 0] - top2[:, -1]).numpy().astype(np.float32)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()
