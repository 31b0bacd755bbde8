"""ORACLE (test infrastructure): numpy restatement of the deterministic part of MVDistilDataset.__getitem__
(data/dataset_blender.py:330-362,400-414), with the reference's random draws passed in. PARITY UNPINNED for
the voxelisation it ends with (MinkowskiEngine, see projections_ref.sparse_quantize_ref); everything before
it is plain numpy indexing copied line by line."""
import numpy as np
import torch

from . import projections_ref as pr


def build_sample_ref(xyz, rgb, label, per_obj, vis_mask, view_ids, indices, voxel_size, use_color=True):
    feat = per_obj[label]                                   # reconstruct_per_obj_feat :128-130
    if view_ids is not None and len(view_ids):
        visibility_mask = vis_mask[np.asarray(view_ids, dtype=int), :].sum(0).astype(bool)   # :343-346
        xyz, rgb = xyz[visibility_mask, :], rgb[visibility_mask, :]
        label = label[visibility_mask].astype(np.uint8)
        feat = feat[visibility_mask, :]
    xyz, rgb, label, feat = xyz[indices, :].copy(), rgb[indices, :], label[indices], feat[indices, :]   # :358-361
    xyz -= xyz.mean(0)                                      # :364
    t_xyz, t_rgb = torch.from_numpy(xyz).float(), torch.from_numpy(np.ascontiguousarray(rgb)).float()
    t_feat, t_lab = torch.from_numpy(np.ascontiguousarray(feat)).float(), torch.from_numpy(np.ascontiguousarray(label)).int()
    cat = torch.cat([t_feat, t_xyz] + ([t_rgb] if use_color else []), dim=-1)
    coords, vfeat, vlab, umap, imap = pr.sparse_quantize_ref(t_xyz.numpy(), cat.numpy(), t_lab.numpy(), ignore_label=0,
                                                            quantization_size=voxel_size)
    return {"xyz": t_xyz.numpy(), "rgb": t_rgb.numpy(), "feat": t_feat.numpy(), "raw_label": t_lab.numpy(), "coords": coords,
            "vfeat": vfeat, "vlabels": vlab, "unique_map": umap, "inverse_map": imap}


def view_clip_ref(pc, world_matrix, K, clip_feature, h=480, w=640):
    """generate_view_clip (data/dataset_blender.py:132-171) without the file reads and the CLIP tower:
    `clip_feature` is the (patch_h, patch_w, C) tensor of :151. Pinned by tests/golden/view_clip.npz (outputs of the
    unmodified reference method run with a stub extractor, tests/make_golden_view_clip.py)."""
    pc = np.asarray(pc)
    projected = np.zeros((pc.shape[0], 2), dtype=int)                                   # :139
    inv = np.linalg.inv(np.asarray(world_matrix))                                       # utils/transforms.py:54
    cam = np.dot(inv, np.vstack([pc.T, np.ones((1, pc.shape[0]))]))[:3, :].T            # :57-59
    cam[:, 1] = -cam[:, 1]                                                              # :134-137
    cam[:, 2] = -cam[:, 2]
    q = (np.asarray(K) @ cam.T).T                                                       # :148
    mask = q[:, 2] != 0                                                                 # :149
    with np.errstate(all="ignore"):
        projected[mask] = np.column_stack([[q[:, 0][mask] / q[:, 2][mask], q[:, 1][mask] / q[:, 2][mask]]]).T   # :150-151
    up = torch.nn.functional.interpolate(clip_feature.permute(2, 0, 1).unsqueeze(0), size=(h, w), mode="bicubic",
                                         align_corners=False).squeeze().permute(1, 2, 0)   # :154-159
    projected[:, 1] = np.clip(projected[:, 1], 0, h - 1)                                # :161-162
    projected[:, 0] = np.clip(projected[:, 0], 0, w - 1)
    return up[projected[:, 1], projected[:, 0]].cpu(), projected                        # :168-170
