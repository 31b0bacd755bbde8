"""shard.run_scene_driver at the configs[1] scene size: scenes as .npy directories on tmpfs -> loader threads (readinto the
pinned slots) -> FusionPipeline -> writer threads (reference file layout). UNIQUE scenes are generated and linked under
SCENES distinct ids (page-cache resident inputs: this measures the loop, not the disk)."""
import json, os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200 import shard
from dropclip_b200.scenes import make_scene

UNIQUE = int(os.environ.get("UNIQUE", "8")); SCENES = int(os.environ.get("SCENES", "96"))
LOADERS = int(os.environ.get("LOADERS", "8")); WRITERS = int(os.environ.get("WRITERS", "4")); B = int(os.environ.get("B", "4"))
base = os.environ.get("DC_TMP", "/dev/shm/dc_driver")
shutil.rmtree(base, ignore_errors=True)
os.makedirs(base + "/in")
for i in range(UNIQUE):
    sc = make_scene(1234 + i, n_views=73, n_points=100_000, n_objects=21, device="cuda")
    shard.SceneDirSource.save(base + "/in", i, sc, seg_dtype=np.uint8, objects_info="{}")
for i in range(UNIQUE, SCENES):
    os.symlink("{:0>6}".format(i % UNIQUE), base + "/in/{:0>6}".format(i))
src = shard.SceneDirSource(base + "/in")
for write in (False, True):
    for rep in range(2):
        shutil.rmtree(base + "/out", ignore_errors=True)
        st = shard.run_scene_driver(src, base + "/out", sc.intrinsic, device="cuda:0", batch_scenes=B, loader_threads=LOADERS,
                                    writer_threads=WRITERS, write=write, fmt="npz", n_slots=3 * B)
        print(json.dumps({"write": write, "rep": rep, "scenes_per_s": st["fused"] / st["seconds"], "fused": st["fused"],
                          "h2d_gbs": st["h2d_bytes"] / st["seconds"] / 1e9, "written_gbs": st["written_bytes"] / st["seconds"] / 1e9,
                          "loader_thread_seconds_per_scene": st["loader_seconds"] / max(1, st["fused"]), "errors": len(st["errors"])}))
shutil.rmtree(base, ignore_errors=True)
