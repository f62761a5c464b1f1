"""Text-only baseline (02_BioClinicalBERT.py: UnstructuredClassifier, FocalLoss, train_model) on the B200 kernels against
golden vectors of the unmodified reference.  The head is fp32 end to end, so tolerances are tight."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SHAPES = {"classifier.0.weight": (256, 768), "classifier.0.bias": (256,), "classifier.3.weight": (3, 256),
          "classifier.3.bias": (3,)}


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


def _model():
    from fairmultimodal_b200 import modules, synth, unstructured as U
    m = U.UnstructuredClassifier(768, 256)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == SHAPES
    m.load_state_dict({k: torch.from_numpy(synth.synth_tensor(k, shp, 13) * (3.0 if "weight" in k else 1.0))
                       for k, shp in SHAPES.items()})
    modules.set_dropout(m, 0.0)
    return m.cuda()


def test_classifier_and_focal_loss_match_reference(golden_dir):
    from fairmultimodal_b200 import unstructured as U
    g = np.load(os.path.join(golden_dir, "text_classifier.npz"))
    model = _model().eval()
    logits = model(torch.from_numpy(g["emb"]).cuda())
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits_eval"], atol=2e-5, rtol=1e-4)
    crit = U.FocalLoss(gamma=2, pos_weight=torch.tensor(float(g["pos_weight"][0])), reduction="mean")
    fl = crit(torch.from_numpy(g["fl_z"]).cuda(), torch.from_numpy(g["fl_y"]).cuda())
    assert abs(fl.item() - float(g["fl_value"])) < 1e-5
    # gradient of the first batch
    model.train()
    pw = torch.from_numpy(g["pos_weight"]).cuda()
    loss, _ = U.forward_backward(model, torch.from_numpy(g["emb"][:8]).cuda(), torch.from_numpy(g["labels"][:8]).cuda(), pw)
    assert abs(loss.item() - float(g["loss_batch0"])) < 1e-5
    st = U.get_state(model)
    for k in SHAPES:
        ref = g["grad." + k]
        np.testing.assert_allclose(st.gr(k).cpu().numpy()[:ref.shape[0]], ref, rtol=1e-3, atol=1e-6)


def test_train_model_epoch_matches_reference(golden_dir):
    """Drop-in train_model: three batches, summed focal losses, AdamW without clipping -> same mean loss and weights."""
    from fairmultimodal_b200 import modules, unstructured as U
    g = np.load(os.path.join(golden_dir, "text_classifier.npz"))
    model = _model()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    ds = torch.utils.data.TensorDataset(t(g["emb"]), t(g["labels"][:, 0:1]), t(g["labels"][:, 1:2]), t(g["labels"][:, 2:3]))
    loader = torch.utils.data.DataLoader(ds, batch_size=8, shuffle=False)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
    crit = [U.FocalLoss(gamma=2, pos_weight=torch.tensor(float(p)), reduction="mean") for p in g["pos_weight"]]
    loss = U.train_model(model, loader, opt, "cuda", *crit)
    assert abs(loss - float(g["epoch_loss"])) < 2e-5
    sd = model.state_dict()
    for k in SHAPES:
        ref = g["after." + k]
        np.testing.assert_allclose(sd[k].cpu().numpy()[:ref.shape[0]], ref, rtol=1e-3, atol=1e-5)
    # with dropout the epoch still runs and the loss stays finite and of the same order
    modules.set_dropout(model, 0.1)
    l2 = U.train_model(model, loader, opt, "cuda", *crit)
    assert np.isfinite(l2) and 0.1 < l2 < 5.0
