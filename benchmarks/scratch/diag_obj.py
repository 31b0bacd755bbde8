import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from dropclip_b200.scenes import make_scene
from dropclip_b200.feature_fusion import MultiviewFeatureFusion
from oracle import fusion_ref
import unittest.mock as m
sc = make_scene(1234, n_views=73, n_points=100000, n_objects=21, device="cpu")
M = MultiviewFeatureFusion(sc.intrinsic, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False, device="cuda")
(feat, w, vis), _ = M.fuse(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features, sc.query_embeddings, return_obj=True, device="cuda")
K = fusion_ref.intrinsic_matrix(sc.intrinsic)
v1 = torch.ones((73, sc.n_points), dtype=torch.int64)
with m.patch.object(fusion_ref, "visibility_mask", lambda *a, **k: v1):
    (f32,w32,_),_ = fusion_ref.fuse_object_level(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features, sc.query_embeddings, K, 480, 640, return_obj=True)
    (f64,w64,_),_ = fusion_ref.fuse_object_level(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features, sc.query_embeddings, K, 480, 640, return_obj=True, work=torch.float64)
w=w.cpu().numpy(); feat=feat.cpu().numpy(); w32=w32.numpy(); w64=w64.numpy(); f32=f32.numpy(); f64=f64.numpy()
rel=np.abs(w-w64)/np.maximum(np.abs(w64),1e-30); rel[w64==0]=0
idx=np.argsort(rel.ravel())[::-1][:10]
for i in idx:
    o,v=np.unravel_index(i,rel.shape); print("obj",o,"view",v,"ours",w[o,v],"ref32",w32[o,v],"exact",w64[o,v],"rel",rel[o,v])
print("median rel", np.median(rel[w64>0]))
ok=~np.isnan(f64); scale=np.maximum(np.abs(f64),np.abs(f64[ok]).max()*1e-3)
e=np.where(ok,np.abs(feat-f64)/scale,0); print("feat err per object", np.round(e.max(1),5))
