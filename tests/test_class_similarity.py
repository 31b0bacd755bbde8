"""a18: `_get_similarity` + arg max of the reference's evaluation loops (engine/distil.py:244-246,289-290,
tools/validate_upper_bound.py:59-61,101-102) against outputs of the unmodified reference closure
(tests/golden/class_sim.npz, generator tests/make_golden_class_sim.py), and the prompt axis beyond one 256-column
GEMM block (models/similarity.py:47-61 has no prompt limit)."""
import numpy as np
import pytest
import torch

from tests import golden_io as gio
from tests.make_golden_class_sim import CASES, class_inputs


def _rows(m):
    return slice(None, None, max(1, m // 512))


@pytest.mark.parametrize("case", CASES)
def test_oracle_class_similarity_vs_reference_golden(case):
    from oracle import similarity_ref
    name, seed, m, k, c = case
    g = gio.load("class_sim.npz")
    x, table = class_inputs(seed, m, k, c)
    sims, pred = similarity_ref.class_similarity(x, table)
    np.testing.assert_allclose(sims[_rows(m)].numpy(), g[f"{name}_sims_rows"], rtol=1e-5, atol=1e-5)
    sure = g[f"{name}_margin"] > 1e-4
    assert np.array_equal(pred.numpy()[sure], g[f"{name}_pred"].astype(np.int64)[sure])
    np.testing.assert_allclose(table.numpy(), g[f"{name}_table_after"], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_class_similarity_vs_reference_golden(case):
    from dropclip_b200.similarity import _get_similarity, class_similarity
    name, seed, m, k, c = case
    g = gio.load("class_sim.npz")
    x, table = class_inputs(seed, m, k, c)
    xd, td = x.cuda(), table.cuda()
    sims, pred = class_similarity(xd, td)
    assert sims.shape == (m, k) and sims.dtype == torch.float32 and pred.shape == (m,) and pred.dtype == torch.int64
    want = g[f"{name}_sims_rows"]
    got = sims[_rows(m)].cpu().numpy()
    assert np.abs(got - want).max() <= 1e-3 * np.abs(want).max(), np.abs(got - want).max()
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=2e-5)
    sure = g[f"{name}_margin"] > 1e-4  # rows whose two best classes are closer than fp32 summation noise may flip
    assert sure.mean() > 0.99
    assert np.array_equal(pred.cpu().numpy()[sure], g[f"{name}_pred"].astype(np.int64)[sure])
    # arg max of OUR matrix is exactly what the epilogue kept (first index on ties)
    assert torch.equal(pred, torch.max(sims, 1)[1])
    np.testing.assert_allclose(td.cpu().numpy(), g[f"{name}_table_after"], rtol=2e-6)  # class table normalised in place
    # without the matrix: same arg max, nothing of size M x K written
    none, pred2 = class_similarity(xd, table.cuda(), return_sims=False)
    assert none is None and torch.equal(pred2, pred)
    assert torch.equal(_get_similarity(xd, table.cuda()), sims)


@pytest.mark.gpu
def test_class_similarity_empty_and_fp16_and_errors():
    from dropclip_b200.similarity import class_similarity
    t = torch.randn((44, 768), device="cuda")
    s, p = class_similarity(torch.empty((0, 768), device="cuda"), t.clone())
    assert s.shape == (0, 44) and p.shape == (0,) and p.dtype == torch.int64
    x = torch.randn((300, 768), device="cuda").half()
    s, p = class_similarity(x, t.clone().half())
    ref = x.float() @ (t / t.norm(dim=-1, keepdim=True)).half().float().T
    assert s.dtype == torch.float16 and (s.float() - ref).abs().max().item() < 3e-2
    with pytest.raises(RuntimeError):
        class_similarity(torch.randn((4, 512), device="cuda"), t)
    with pytest.raises(RuntimeError):
        class_similarity(torch.randn((4, 768)), t)


@pytest.mark.gpu
@pytest.mark.parametrize("n_prompts", [257, 1000])
def test_more_than_256_prompts(n_prompts):
    """use_kernel_neg == 'all' concatenates the whole class table (tools/preprocess_data.py:253-257); the reference has
    no prompt limit. The 256-column GEMM blocks combine their partial sums / maxima / arg max through the carry rows."""
    from oracle import ref_shim, similarity_ref
    from tests.test_gpu_parity import make_cs
    dim, n = 768, 3000
    rng = np.random.default_rng(31)
    prompts = [f"class {i}" for i in range(n_prompts)]
    tower = ref_shim.FakeTextTower(dim, torch.float32)
    emb = tower.encode_text(ref_shim.fake_tokenize(prompts))
    x = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32))
    x[: n // 3] += 4.0 * emb[0] / emb[0].norm()
    x[n // 3: n // 2] += 4.0 * emb[n_prompts - 3] / emb[n_prompts - 3].norm()  # a strong negative in the LAST block
    cs = make_cs(dim, torch.float32)
    for method in ("paired", "argmax"):
        want_pred, want = similarity_ref.predict_from_embeds(x.clone(), emb[:1].clone(), emb[1:].clone(), method=method)
        xd = x.clone().cuda()
        pred, sims = cs.predict(xd, prompts[0], qneg=prompts[1:], method=method, threshold=0.7)
        np.testing.assert_allclose(sims.cpu().numpy(), want.numpy(), rtol=1e-3, atol=1e-5)
        if method == "paired":
            sure = (want - 0.7).abs() > 1e-3
            assert torch.equal(pred.cpu()[sure], want_pred[sure])
        else:
            assert (pred.cpu() != want_pred).float().mean().item() < 2e-3
    xn = x.clone().cuda()
    xn /= xn.norm(dim=-1, keepdim=True)
    raw = cs.compute_similarity(xn, prompts[0], prompts[1:], method="argmax")
    t = torch.cat([emb[:1], emb[1:]]).clone()
    t /= t.norm(dim=-1, keepdim=True)
    ref = xn.cpu() @ t.T
    assert raw.shape == (n, n_prompts) and (raw.cpu() - ref).abs().max().item() < 2e-5
    from dropclip_b200.similarity import class_similarity
    sims, idx = class_similarity(x.clone().cuda(), emb.clone().cuda())
    assert torch.equal(idx, torch.max(sims, 1)[1])
    rs = x @ t.T
    top2 = torch.topk(rs, 2, dim=1).values
    sure = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(idx.cpu()[sure], rs.argmax(1)[sure])
