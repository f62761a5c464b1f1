"""Concat-fusion ablation (07_multimodal_average_fusion.py: seven-table BEHRTModel, 2 x 256 concat head, FocalLoss gamma 1,
plain Adam without clipping) on the B200 kernels against golden vectors of the unmodified reference."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
WSEED = 29
SIZES = dict(num_diseases=10, num_ages=5, num_segments=2, num_adm=4, num_disch=6, num_genders=2, num_eth=5, num_ins=5)


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


def _model():
    from fairmultimodal_b200 import average_fusion as AF, modules, synth
    behrt = AF.BEHRTModel(SIZES["num_diseases"], SIZES["num_ages"], SIZES["num_segments"], SIZES["num_adm"], SIZES["num_disch"],
                          SIZES["num_genders"], SIZES["num_eth"], SIZES["num_ins"])
    m = AF.MultimodalTransformer(768, behrt, "cuda")
    shapes = synth.average_fusion_shapes(**SIZES)
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())                      # same keys, same order as the reference
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, WSEED).items()}, strict=True)
    modules.set_dropout(m, 0.0)
    return m.cuda()


def _inputs(g):
    B = g["labels"].shape[0]
    ids = torch.zeros((B, 1), dtype=torch.long, device="cuda")
    mask = torch.ones((B, 1), dtype=torch.long, device="cuda")
    return [ids, mask, *[torch.from_numpy(c).cuda() for c in g["codes"]], torch.from_numpy(g["text"]).cuda()], \
        torch.from_numpy(g["labels"]).cuda()


def test_embed_mean_kernels_bit_exact_vs_torch():
    """n-table code embedding mean (incl. the clamp of out-of-range codes) and its backward scatter."""
    from fairmultimodal_b200 import ops
    torch.manual_seed(0)
    B, H = 37, 768
    rows = (5, 2, 4, 6, 2, 5, 5)
    tabs = [torch.randn(r, H, device="cuda") for r in rows]
    ids = [torch.randint(-1, r + 2, (B,), device="cuda") for r in rows]
    cls = torch.randn(B, H, device="cuda")
    out = ops.embed_mean_add(cls, H, ids, tabs)
    # reference arithmetic on the CPU, where the golden vectors were made: torch's CPU kernel divides (IEEE), its CUDA
    # kernel multiplies by the rounded reciprocal of a scalar divisor -- 1/7 is not exact, so the two differ in the last bit
    extra = 0
    for t, i in zip(tabs, ids):
        extra = extra + t.cpu()[i.cpu().clamp(0, t.shape[0] - 1)]
    assert torch.equal(out.cpu(), cls.cpu() + extra / 7.0)
    dout = torch.randn(B, H, device="cuda")
    dt = [torch.zeros_like(t) for t in tabs]
    ops.embed_mean_add_bwd(dout, ids, dt)
    for t, i, d in zip(tabs, ids, dt):
        ref = torch.zeros_like(t).index_add_(0, i.clamp(0, t.shape[0] - 1), dout / 7.0)
        assert torch.allclose(d, ref, rtol=1e-5, atol=1e-6)            # float atomics: summation order only


def test_forward_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "average_fusion.npz"))
    model = _model().eval()
    batch10, _ = _inputs(g)
    a, b, c, pre = model(*batch10)
    assert a.shape == (g["labels"].shape[0], 1) and pre.shape == (g["labels"].shape[0], 512)
    got = torch.cat([a, b, c], dim=1).cpu().numpy()
    assert np.abs(got - g["logits_eval"]).max() <= 1e-2 * np.abs(g["logits_eval"]).max()     # bf16 tower: rel 1e-2
    assert np.abs(pre.cpu().numpy() - g["pre_relu_eval"]).max() <= 1e-2 * np.abs(g["pre_relu_eval"]).max()


def test_training_step_matches_reference_golden(golden_dir):
    from fairmultimodal_b200 import average_fusion as AF
    g = np.load(os.path.join(golden_dir, "average_fusion.npz"))
    model = _model().train()
    batch10, labels = _inputs(g)
    pw = torch.from_numpy(g["pos_weight"]).cuda()
    loss, _ = AF.forward_backward(model, batch10, labels, pw, gamma=1.0)
    assert abs(loss.item() - float(g["loss"])) < 5e-3
    st = AF.get_state(model)
    names = [str(n) for n in g["gnorm_names"]]
    got = np.array([st.gr(n).norm().item() for n in names])
    ref = g["gnorm"]
    live = ref > 1e-5 * ref.max()          # query / key projections and unused word-embedding rows: exactly zero here
    rel = np.abs(got[live] - ref[live]) / ref[live]
    assert np.median(rel) < 0.03 and rel.max() < 0.3, sorted(zip(rel, np.array(names)[live]))[-5:]
    assert np.all(got[~live] <= 1e-5 * ref.max())
    assert not any(n.startswith("BEHRT.bert.pooler.") for n in st.offsets)          # grad None in the reference
    errs = {}
    for k in g.files:
        if k.startswith("grad."):
            mine = st.gr(k[5:]).cpu().numpy()
            errs[k] = float(np.linalg.norm(mine - g[k]) / (np.linalg.norm(g[k]) + 1e-12))
    print("07 per-tensor gradient errors:", errs)
    assert max(errs.values()) < 0.25 and np.median(list(errs.values())) < 0.12, errs
    # the drop-in epoch: plain Adam, no clipping; returns the SUM of batch losses (one batch here)
    model2 = _model()
    ds = torch.utils.data.TensorDataset(*[b.cpu() for b in batch10], labels[:, 0].cpu(), labels[:, 1].cpu(), labels[:, 2].cpu())
    loader = torch.utils.data.DataLoader(ds, batch_size=labels.shape[0], shuffle=False)
    opt = torch.optim.Adam(model2.parameters(), lr=1e-4)
    crit = [AF.FocalLoss(gamma=1, pos_weight=torch.tensor(float(p)), reduction="mean") for p in g["pos_weight"]]
    before = {k: v.clone() for k, v in model2.state_dict().items()}
    ep = AF.train_step(model2, loader, opt, "cuda", *crit)
    assert abs(ep - float(g["loss"])) < 5e-3
    after = model2.state_dict()
    # Adam's first step moves every element with a non-zero gradient by lr (no weight decay, no clipping)
    d = (after["classifier.3.weight"] - before["classifier.3.weight"]).abs()
    # (hidden units whose ReLU is off for every patient of this small batch have gradient exactly 0 and stay put)
    assert d.max().item() <= 1.01e-4 and bool(((d > 0.9e-4) | (d == 0)).all()) and (d > 0.9e-4).float().mean().item() > 0.5
    assert torch.equal(after["BEHRT.bert.pooler.dense.weight"], before["BEHRT.bert.pooler.dense.weight"])
